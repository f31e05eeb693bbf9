#!/usr/bin/env python3
"""Summarise an `ncu --set full --import-source on` report of the step kernel into the text kept
under profiles/ and into the entry of profiles/ncu_traffic.json that bench.py reads.

    python tools/summarize_profile.py <report.ncu-rep> <members> <steps_per_member> <out.txt> \
        [--key c2:100000:f64:0 --json profiles/ncu_traffic.json]

The JSON entry carries the hash of the CUDA sources (bench.csrc_hash) the capture was made from:
bench.py only uses it for `roofline.frac_pipe` when it runs the same sources.
"""
import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WANT = [
    'Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
    'launch__block_size', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'smsp__average_warp_latency_per_inst_issued.ratio', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'sm__cycles_active.avg', 'sm__cycles_active.min', 'sm__cycles_active.max',
]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("members", type=float)
    ap.add_argument("steps", type=float)
    ap.add_argument("out")
    ap.add_argument("--key")
    ap.add_argument("--json")
    args = ap.parse_args()

    raw = subprocess.run(['ncu', '-i', args.report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    metrics = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    lines = []
    for h in hdr:
        u, v = metrics[h]
        if h in WANT or (h.startswith('smsp__average_warps_issue_stalled') and float(v or 0) > 0.05):
            lines.append('%-78s %-12s %s' % (h, u, v))

    src = subprocess.run(['ncu', '-i', args.report, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    shdr = rows[1]
    i_src, i_exe = shdr.index('Source'), shdr.index('Instructions Executed')
    ops = collections.Counter()
    for r in rows[2:]:
        words = r[i_src].split()
        name = words[0] if not words[0].startswith('@') else words[1]
        ops[name.split('.')[0]] += int(r[i_exe])
    warp_steps = args.members * args.steps / 32
    total = sum(ops.values())
    fp64 = sum(c for n, c in ops.items() if n in ('DADD', 'DMUL', 'DFMA', 'DSETP'))
    fp32 = sum(c for n, c in ops.items() if n in ('FADD', 'FMUL', 'FFMA'))
    lines.append('')
    lines.append('executed warp instructions per member-step-warp (all SASS lines of the kernel):')
    for name, count in ops.most_common(24):
        lines.append('  %-10s %7.2f' % (name, count / warp_steps))
    lines.append('  total %.2f   FP64 (DADD+DMUL+DFMA+DSETP) %.2f   FP32 (FADD+FMUL+FFMA) %.2f' % (
        total / warp_steps, fp64 / warp_steps, fp32 / warp_steps))
    import bench
    sha = bench.csrc_hash()
    lines.append('')
    lines.append('csrc_sha %s   report %s' % (sha, os.path.basename(args.report)))
    text = "\n".join(lines) + "\n"
    with open(args.out, "w") as f:
        f.write(text)
    sys.stdout.write(text)

    if args.key and args.json:
        def nbytes(name):
            u, v = metrics[name]
            return float(v) * UNIT.get(u, 1.0)
        is32 = args.key.split(":")[2] == "f32"
        entry = {
            "dram_bytes_read": nbytes('dram__bytes_read.sum'),
            "dram_bytes_write": nbytes('dram__bytes_write.sum'),
            "pipe_active_pct": float(metrics['sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'][1]) if is32 else
            float(metrics['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'][1]),
            "pipe_active_pct_of_elapsed": None if is32 else
            float(metrics['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed'][1]),
            "issue_active_pct": float(metrics['smsp__issue_active.avg.pct_of_peak_sustained_active'][1]),
            "inst_per_step": total / warp_steps,
            "kernel": "%s, %s %s under ncu, %s registers" % (
                metrics['Kernel Name'][1].replace('void <unnamed>::', '').replace('(<unnamed>::KArgs)', ''),
                metrics['gpu__time_duration.sum'][1], metrics['gpu__time_duration.sum'][0],
                metrics['launch__registers_per_thread'][1]),
            "source": os.path.relpath(args.out, ROOT),
            "csrc_sha": sha,
        }
        entry["fp32_inst_per_step" if is32 else "fp64_inst_per_step"] = (fp32 if is32 else fp64) / warp_steps
        if is32:
            entry["fp64_inst_per_step_on_the_side"] = fp64 / warp_steps
        try:
            with open(args.json) as f:
                table = json.load(f)
        except (OSError, ValueError):
            table = {}
        table["_comment"] = ("per launch of the dominant kernel, from ncu --set full captures (tools/measure_round.sh); key = "
                             "workload:members_per_gpu:precision:flags; an entry counts only for the sources it names (csrc_sha)")
        table[args.key] = entry
        with open(args.json, "w") as f:
            json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()
