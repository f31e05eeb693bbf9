#!/usr/bin/env python3
"""Summarise an ncu report (raw + source pages) into the text kept under profiles/.

    python tools/summarize_profile.py <report.ncu-rep> <members> <steps_per_member>
"""
import csv, collections, sys, subprocess
rep=sys.argv[1]; nmembers=float(sys.argv[2]); nsteps=float(sys.argv[3])
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__waves_per_multiprocessor','launch__block_size','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__issue_active.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','smsp__average_warp_latency_per_inst_issued.ratio','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__maximum_warps_per_active_cycle_pct','smsp__cycles_active.avg']
for h,u,v in zip(hdr,units,vals):
    if h in want or h.startswith('smsp__average_warps_issue_stalled') and float(v or 0)>0.05:
        print('%-75s %-12s %s'%(h,u,v))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; iS=hdr.index('Source'); iE=hdr.index('Instructions Executed')
data=rows[2:]
thr=max(int(r[iE]) for r in data)*0.02
op=collections.Counter()
for r in data:
    if int(r[iE])>thr:
        m=r[iS].split(); name=m[0] if not m[0].startswith('@') else m[1]
        op[name.split('.')[0]]+=int(r[iE])
ws=nmembers*nsteps/32
tot=0
for n,c in op.most_common(25):
    print('%-10s %7.2f per warp-step'%(n,c/ws)); tot+=c/ws
print('hot total',tot, ' fp64:',sum(op[k] for k in ('DADD','DMUL','DFMA','DSETP'))/ws)
