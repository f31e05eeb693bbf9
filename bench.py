#!/usr/bin/env python3
"""bench.py -- the SMART hot path on B200: member-timesteps/s and FP64-pipe roofline fraction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c3|c4a|c4b|c5] [--precision f64|f32] [--members M] [--no-also]

One "step" = one pass of the hot path over one batch of members (warm-up run + main run,
objective functions fused) on synthetic/fixture forcing:

  c2 (default, BASELINE.json configs[1]): LHS sample of 1e5 parameter sets on the reference's
      test catchment (2007-2016 hourly, 87,672 steps + 8,760 warm-up), NSE/KGE/... scored.
  c3: 30-year hourly synthetic forcing (262,992 + 8,760 steps), 1.25e6 members per GPU, scores only
      (8 GPUs = BASELINE's target run, 1e7 parameter sets).
  c4a: 1e4 synthetic catchments x 100 members, daily [t][catchment] totals split on the device,
      8,760 hourly steps, HOURLY discharge written (70 GB per GPU).
  c4b: same catchments, 10 years, daily discharge written.
  c5: c2 shape in FP32 state (--precision f32).

The headline line is --workload (c2 unless asked otherwise); unless --no-also is given, a few timed
steps of each of the OTHER configurations are measured in the same run and reported under `also`
(each with its own ms, throughput, e2e, pipe and HBM fractions), so that the driver sees every
BASELINE configuration and not only C2.

With N > 1 (torchrun, one rank per GPU) every rank runs the same per-GPU batch on its own LHS
rows / its own catchments (weak scaling, members are independent) and a scored step ends with one
all-gather of the [members, 9] block (8 scores + gw) over NCCL -- the block the kernel wrote.

`value` has inputs resident in HBM.  `e2e` goes through the public host-buffer API
(`BatchEngine.run_host`: parameter rows in the engine's page-locked input buffer in, numpy scores
out of its pinned staging, H2D and D2H inside the timed region; a plain numpy array is accepted too
and costs one staging copy more).  `roofline.frac_pipe` = FP64 (FP32) instructions the
kernel EXECUTES per member-step (ncu capture of this very source, checked by hash) x throughput /
FMA-pipe peak MEASURED in this run; `roofline.frac_alg` is the same with SURVEY.md 8(d)'s
algorithmic instruction count (the contract figure; > 1 because the kernel undercuts the per-step
restatement).  `roofline.frac` = frac_pipe when a capture of this source exists, else frac_alg.
`cpu_baseline` is the C oracle on the host cores.

--impl reference: the reference's CPU algorithm for this path on all host cores.  The
reference itself is pure Python and cannot travel to the GPU box, so this arm runs the
oracle's C restatement of it (bit-identical discharge, ~80x faster than the Python original:
3.7e6 vs 4.6e4 member-timesteps/s/core measured in the build container).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

EXTRA = {'aar': 1200, 'r-o_ratio': 0.45, 'r-o_split': (0.10, 0.15, 0.15, 0.30, 0.30)}
METRIC = "member-timesteps/s"

# algorithmic FP64-pipe instructions per member-timestep of the minimal
# faithful restatement (FMA = one instruction): 65.5 on a dry step, 135.5 on a wet step.
I_DRY, I_WET = 65.5, 135.5

# members each host worker simulates per CPU-arm step (C oracle: ~26 ms per 96k-step member)
CPU_MEMBERS_PER_WORKER = 96      # a bounded sample: a few seconds of wall time per CPU-arm step on 16 cores

ALSO = ("c3", "c4a", "c4b", "c5")            # configurations reported beside the headline one


# ---------------------------------------------------------------------------------------- workloads
def synthetic_forcing(n_days, catchment_id=0):
    """daily values, each / 24 and repeated 24x (as timeframe.py:181-183 does)."""
    rain_d, peva_d = synthetic_daily(n_days, catchment_id)
    return np.repeat(rain_d / 24, 24), np.repeat(peva_d / 24, 24)


def synthetic_daily(n_days, catchment_id=0):
    rng = np.random.Generator(np.random.PCG64(20260101 + catchment_id))
    wet = rng.random(n_days) < 0.80
    rain_d = np.where(wet, rng.gamma(0.7, 4.57, n_days), 0.0)
    doy = np.arange(n_days) % 365.25
    peva_d = np.maximum(0.1, 1.47 + 1.2 * np.sin(2 * np.pi * (doy - 100) / 365.25))
    return rain_d, peva_d


def lhs_rows(n, seed):
    from smartpy_b200.montecarlo.lhs import latin_hypercube
    from smartpy_b200.parameters import Parameters
    p = Parameters()
    rng = np.random.RandomState(seed)
    return latin_hypercube(n, [p.ranges[name] for name in p.names], rng=rng)


def simulate_on_device(w, params):
    """Daily discharge of ONE parameter set on workload w's forcing, through the CUDA path (used to
    build synthetic observations; the CPU arm passes the oracle's run instead)."""
    from smartpy_b200.engine import BatchEngine
    eng = BatchEngine(w["rain"], w["peva"], w["area"], w["dt"], w["gap"], extra=w["extra"],
                      warm_up_steps=w["warm_steps"], report=w["report"])
    return eng.run(np.asarray(params)[None, :], discharge=True, scores=False, gw=False)["discharge"][:, 0].cpu().numpy()


def simulate_on_host(w, params):
    """Same with the CPU oracle (bench.py's reference arm / cpu_baseline only)."""
    import oracle
    return oracle.run(w["area"], w["dt"], w["rain"], w["peva"], params, w["extra"], w["n_steps"], w["gap"],
                      report=w["report"], warm_up=w["warm_steps"] * w["dt"] / 86400.0)[0]


def make_workload(name, rank, members=None, simulate=simulate_on_device):
    """-> dict(rain, peva, area, obs, dt, gap, warm_steps, params, label, discharge, mpc)."""
    golden = os.path.join(ROOT, "tests", "golden")
    w = dict(dt=3600.0, gap=24, warm_steps=8760, extra=EXTRA, gwc=0.12667, discharge=False, mpc=None,
             report='summary')
    if name in ("c2", "c5"):
        g = np.load(os.path.join(golden, "catchment_processed.npz"))
        w.update(rain=np.repeat(g["rain_hourly_per_day"], 24), peva=np.repeat(g["peva_hourly_per_day"], 24),
                 obs=g["nd_flow"], area=float(g["area_m2"]))
        n = members or 100000
        w["label"] = "LHS {} parameter sets x test catchment (87672 hourly steps + 8760 warm-up), scores only".format(n)
    elif name == "c3":
        n_days = 10958
        rain, peva = synthetic_forcing(n_days)
        w.update(rain=rain, peva=peva, area=175.46e6)
        # observations = a run of the reference's test parameter file on this forcing
        # x exp(N(0, 0.1^2)) daily noise, 12 % of the days missing (seed + 1)
        p_test = np.load(os.path.join(golden, "runs_single.npz"))["p_test"]
        q = simulate(dict(w, n_steps=rain.shape[0]), p_test)
        rng = np.random.Generator(np.random.PCG64(20260102))
        obs = q * np.exp(rng.normal(0.0, 0.1, n_days))
        obs[rng.random(n_days) < 0.12] = np.nan
        w.update(obs=obs)
        n = members or 1250000
        w["label"] = "LHS {} parameter sets per GPU x 30 yr hourly synthetic forcing (262992 + 8760 steps), scores only".format(n)
    elif name in ("c4a", "c4b"):
        # daily totals in [t][catchment] layout, split on the device (forcing_repeat = 24, the
        # equal split of timeframe.py:167-186).  c4a: one year, HOURLY discharge written (gap 1,
        # 'raw' == hourly values); c4b: ten years, daily means written.
        n_catch, mpc = (members // 100 if members else 10000), 100
        n_days = 365 if name == "c4a" else 3653
        rain = np.empty((n_days, n_catch))
        peva = np.empty((n_days, n_catch))
        for c in range(n_catch):
            rain[:, c], peva[:, c] = synthetic_daily(n_days, rank * n_catch + c)
        rng = np.random.Generator(np.random.PCG64(7))
        area = np.exp(rng.uniform(np.log(10e6), np.log(2000e6), n_catch))
        w.update(rain=rain, peva=peva, obs=None, area=area, warm_steps=0, discharge=True, mpc=mpc, gwc=None,
                 forcing_repeat=24)
        if name == "c4a":
            w.update(gap=1, report='raw')
        n = n_catch * mpc
        w["label"] = ("{} synthetic catchments x {} members, daily [t][catchment] forcing split on the device, {} hourly "
                      "steps, {} discharge written").format(n_catch, mpc, n_days * 24, "hourly" if name == "c4a" else "daily")
    else:
        raise SystemExit("unknown workload " + name)
    w["params"] = lhs_rows(n, 42 + rank)
    w["n_members"] = n
    w["n_steps"] = w["rain"].shape[0] * w.get("forcing_repeat", 1)
    w["name"] = name
    return w


def wet_fraction(w):
    """Fraction of member-steps taking the wet branch (T in 0.9..1.1, midpoint 1.0 used)."""
    rain = w["rain"] if w["rain"].ndim == 1 else w["rain"][:, 0]
    peva = w["peva"] if w["peva"].ndim == 1 else w["peva"][:, 0]
    k = w.get("forcing_repeat", 1)
    seq = (np.concatenate([rain[:w["warm_steps"] // k], rain]), np.concatenate([peva[:w["warm_steps"] // k], peva]))
    return float(np.mean(seq[0] * 1.0 - seq[1] >= 0.0))


def csrc_hash():
    """sha256 (first 16 hex digits) of the CUDA sources + the ABI header: a capture under profiles/
    only counts as evidence for the code that produced it."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "smartpy_b200", "csrc")
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh")))
    for path in files + [os.path.join(ROOT, "include", "smart_b200.h")]:
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def capture_for(captures, key, sha):
    """The ncu capture (profiles/ncu_traffic.json) of workload `key`, and whether it counts: only a
    capture of the very sources that are running (csrc_sha) may feed roofline.frac_pipe / traffic."""
    cap = captures.get(key) or {}
    return cap, bool(cap) and cap.get("csrc_sha") == sha


# ---------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [p.strip() for p in line.split(",")]
                    if len(parts) < 8:
                        continue
                    try:
                        sm.append(float(parts[1]))
                        out["sm_max_mhz"] = float(parts[2])
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         parts[4:8]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            # under load = the upper half of the samples (idle samples bracket the timed region)
            sm.sort()
            out["sm_mhz"] = float(np.median(sm[len(sm) // 2:]))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------- CPU arm (oracle)
_CPU = {}


def _cpu_init(w):
    import oracle
    oracle.build()
    _CPU["w"] = w
    _CPU["oracle"] = oracle


def _cpu_task(rows):
    w, oracle = _CPU["w"], _CPU["oracle"]
    from oracle import scores as oscores
    k = w.get("forcing_repeat", 1)
    rain = w["rain"] if w["rain"].ndim == 1 else np.ascontiguousarray(w["rain"][:, 0])
    peva = w["peva"] if w["peva"].ndim == 1 else np.ascontiguousarray(w["peva"][:, 0])
    if k > 1:
        rain, peva = np.repeat(rain / k, k), np.repeat(peva / k, k)
    area = w["area"] if np.isscalar(w["area"]) else float(np.asarray(w["area"]).ravel()[0])
    warm_days = w["warm_steps"] * w["dt"] / 86400.0
    for p in rows:
        q, gw = oracle.run(area, w["dt"], rain, peva, p, w["extra"], w["n_steps"], w["gap"],
                           report=w["report"], warm_up=warm_days)
        if w.get("obs") is not None:
            oscores.objectivefunction((q, [gw]), (w["obs"], [w["gwc"]]), w["gwc"])
    return len(rows)


def make_pool(w):
    """One worker per host core, each holding the workload (forcing, obs) and the loaded oracle."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")     # the parent may hold a CUDA context: never fork it
    slim = {k: v for k, v in w.items() if k != "params"}
    pool = ctx.Pool(cores, initializer=_cpu_init, initargs=(slim,))
    pool.map(_cpu_task, [w["params"][:1]] * cores)   # import + load once per worker, untimed
    return pool, cores


def cpu_throughput(w, pool, cores, per_worker):
    """Oracle on all host cores: `cores` workers x per_worker members of the workload's shape.
    Returns (member_steps_per_s, seconds, sample_text)."""
    chunks = [w["params"][(i * per_worker) % max(w["n_members"] - per_worker, 1):][:per_worker] for i in range(cores)]
    t0 = time.perf_counter()
    done = sum(pool.map(_cpu_task, chunks))
    secs = time.perf_counter() - t0
    steps = done * (w["n_steps"] + w["warm_steps"])
    sample = "{} members ({} per worker x {} workers) x {} steps incl. warm-up".format(
        done, per_worker, cores, w["n_steps"] + w["warm_steps"])
    return steps / secs, secs, sample


# ---------------------------------------------------------------------------------------- one workload on the GPU(s)
class Ctx(object):
    """What every measurement of this process shares: device, ranks, L2 flush buffer, measured peaks."""

    def __init__(self, rank, local_rank, world):
        import torch
        self.torch = torch
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.dev = torch.device("cuda", local_rank)
        self.stream = torch.cuda.current_stream(self.dev)
        # L2 flush between timed iterations: overwrite a buffer twice the size of the 126 MB L2
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)
        self.peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                self.peaks = json.load(f)
        except (OSError, ValueError):
            pass
        self.captures, self.hash = {}, csrc_hash()
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                self.captures = json.load(f)
        except (OSError, ValueError):
            pass
        self._fma_peak = {}

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, k):
        """k steps, each bracketed by CUDA events on the launching stream, L2 flushed between."""
        torch = self.torch
        total_ms = 0.0
        for _ in range(k):
            self.flush.fill_(1.0)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            fn()
            e1.record(self.stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        return total_ms

    def max_over_ranks(self, *values):
        if self.world == 1:
            return values
        import torch.distributed as dist
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(v) for v in t)

    def fma_peak(self, bits):
        """FMA-pipe peak measured live with the library's probe kernel (thread-instructions/s)."""
        if bits not in self._fma_peak:
            from smartpy_b200.engine import fma_peak
            self._fma_peak[bits] = fma_peak(bits, threads=256, iters=1 << 15)[0]
        return self._fma_peak[bits]


def measure(ctx, name, precision, steps, warmup, members=None, flags=0, do_e2e=True):
    """Time `steps` steps of workload `name` on every rank.  Returns (result dict, workload)."""
    import torch.distributed as dist
    from smartpy_b200.engine import BatchEngine
    torch, dev, world = ctx.torch, ctx.dev, ctx.world
    w = make_workload(name, ctx.rank, members)
    n = w["n_members"]
    eng = BatchEngine(w["rain"], w["peva"], w["area"], w["dt"], w["gap"], obs=w.get("obs"), extra=w["extra"],
                      warm_up_steps=w["warm_steps"], report=w["report"], gw_constraint=w["gwc"],
                      members_per_catchment=w["mpc"], precision=precision, flags=flags,
                      forcing_repeat=w.get("forcing_repeat", 1))
    scored = w.get("obs") is not None
    member_steps = eng.member_steps(n)             # per rank per step

    # device-resident inputs/outputs for `value`: the kernel writes scores + gw into the [n, 9] block
    # that is all-gathered as it is
    p_dev = torch.from_numpy(w["params"]).to(dev)
    out = {"block": torch.empty((n, 9), dtype=torch.float64, device=dev)}
    if w["discharge"]:
        out["discharge"] = torch.empty((eng.n_report, n), dtype=torch.float64 if precision == 'f64' else torch.float32,
                                       device=dev)
    gathered = torch.empty((world * n, 9), dtype=torch.float64, device=dev) if world > 1 and scored else None

    def step_resident():
        eng.run(p_dev, discharge=w["discharge"], scores=scored, gw=True, out=out)
        if gathered is not None:
            dist.all_gather_into_tensor(gathered, out["block"])

    # host input of the e2e steps: the engine's own page-locked row buffer, filled once (a sampler
    # writes its rows straight into it); every step uploads it again
    p_in = eng.pinned_rows(n)
    p_in.copy_(torch.from_numpy(w["params"]))

    def host_step_with_discharge():
        # run_host with the [t][member] discharge kept on the device (70 GB at C4a: it stays sharded
        # where it was written); parameters in and the [n, 9] block out as run_host does
        p_stage = eng.stage_params(p_in)
        eng.run(p_stage, discharge=True, scores=scored, gw=True, out=out)
        blk_pin = eng._pinned('block', out["block"].shape, torch.float64)
        blk_pin.copy_(out["block"], non_blocking=True)
        ctx.stream.synchronize()

    def step_e2e():
        # the public host-buffer call: numpy rows in, numpy scores + gw out
        if w["discharge"]:
            host_step_with_discharge()
        else:
            eng.run_host(p_in)
        if gathered is not None:
            dist.all_gather_into_tensor(gathered, eng._staging['dev_block'])

    for _ in range(warmup):
        step_resident()
    ctx.barrier()
    launches_before = eng.kernel_launches
    ms_value = ctx.timed(step_resident, steps)
    n_launches = eng.kernel_launches - launches_before      # the library's own counter
    ctx.barrier()
    ms_e2e = 0.0
    if do_e2e:
        step_e2e()
        ctx.barrier()
        ms_e2e = ctx.timed(step_e2e, steps)
        ctx.barrier()
    ms_value, ms_e2e = ctx.max_over_ranks(ms_value, ms_e2e)

    total_steps = member_steps * world * steps
    value = total_steps / (ms_value * 1e-3)
    ms_per_step = ms_value / steps
    per_gpu = value / world

    # ---- roofline
    bits = 64 if precision == "f64" else 32
    peak_fma = ctx.fma_peak(bits)
    wfrac = wet_fraction(w)
    i_alg = I_DRY + (I_WET - I_DRY) * wfrac
    q_bytes = eng.n_report * n * (8 if precision == "f64" else 4) if w["discharge"] else 0
    hbm_bytes_per_step = n * (80 + 64 + 8) + 2 * 8 * w["rain"].size + q_bytes
    hbm_peak = ctx.peaks.get("hbm_gbs", 6650.0)
    key = "{}:{}:{}:{}".format(name, n, precision, flags)
    cap, cap_ok = capture_for(ctx.captures, key, ctx.hash)
    inst = cap.get("fp64_inst_per_step" if bits == 64 else "fp32_inst_per_step") if cap_ok else None
    frac_alg = per_gpu * i_alg / peak_fma
    frac_pipe = per_gpu * inst / peak_fma if inst else None
    roofline = {
        "bound": "fp64-pipe" if bits == 64 else "fp32-pipe",
        "unit": "T-instr/s (FMA-pipe thread instructions; FMA = 1)",
        "peak": peak_fma / 1e12,
        "achieved": (per_gpu * inst if inst else per_gpu * i_alg) / 1e12,
        "frac": frac_pipe if frac_pipe is not None else frac_alg,
        "frac_is": "frac_pipe" if frac_pipe is not None else "frac_alg (no ncu capture of this source: see capture)",
        "frac_pipe": frac_pipe,
        "frac_alg": frac_alg,
        "executed_inst_per_member_step": inst,
        "i_alg_per_member_step": i_alg, "wet_fraction": wfrac,
        "ncu_pipe_active_pct": cap.get("pipe_active_pct") if cap_ok else None,
        "ncu_issue_active_pct": cap.get("issue_active_pct") if cap_ok else None,
        "capture": ({"source": cap.get("source"), "csrc_sha": cap.get("csrc_sha"), "kernel": cap.get("kernel")} if cap_ok else
                    {"missing": "no capture for {} at csrc_sha {}".format(key, ctx.hash),
                     "stale": cap.get("csrc_sha") if cap else None}),
        "peak_source": "measured in this run: smart_fma_peak_probe (8 dependent FMA chains/thread, 256 thr x 8 CTA/SM)",
        "peak_nominal": 148 * (64 if bits == 64 else 128) * 1.965e9 / 1e12,
        "traffic": (cap["dram_bytes_read"] + cap["dram_bytes_write"]) if cap_ok and "dram_bytes_read" in cap else None,
        "algorithmic_bytes": hbm_bytes_per_step,
        "hbm": {"achieved_gbs": hbm_bytes_per_step / (ms_per_step * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "frac": hbm_bytes_per_step / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in ctx.peaks else "fallback"},
    }
    res = {"value": value, "ms_per_step": ms_per_step, "roofline": roofline, "gpu_launches": n_launches,
           "dtype": precision, "members_per_gpu": n, "steps_per_member": w["n_steps"] + w["warm_steps"],
           "label": w["label"], "report_gap": w["gap"], "timed_steps": steps}
    if do_e2e:
        h2d = n * 10 * 8
        d2h = n * 9 * 8
        res["e2e"] = {"value": total_steps / (ms_e2e * 1e-3), "unit": METRIC,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / steps,
                      "api": ("BatchEngine.run_host: parameter rows in the engine's page-locked input buffer "
                              "(pinned_rows) -> H2D -> smart_batch_run_{} (C ABI) -> [N, 9] block -> D2H into pinned "
                              "staging -> numpy scores + gw{}").format(
                                  precision, "; discharge stays on the device" if w["discharge"] else "")}
    del eng, out, p_dev, gathered, p_in
    torch.cuda.empty_cache()
    return res, w


# ---------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4a", "c4b", "c5"])
    ap.add_argument("--precision", default=None, choices=["f64", "f32"])
    ap.add_argument("--members", type=int, default=None, help="members per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="measure only --workload")
    ap.add_argument("--also-steps", type=int, default=3)
    ap.add_argument("--flags", type=int, default=0, help="SMART_FLAG_* bits passed to the library")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    precision = args.precision or ("f32" if args.workload == "c5" else "f64")

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: smartpy_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created; stdout must
        # carry exactly one JSON line, so the banner is sent to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    ctx = Ctx(rank, local_rank, world)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    head, w = measure(ctx, args.workload, precision, args.steps, args.warmup, members=args.members, flags=args.flags,
                      do_e2e=not args.no_e2e)
    clocks = sampler.stop() if rank == 0 else None

    also = {}
    if not args.no_also and args.members is None and args.flags == 0:
        for name in ALSO + ("c2",):
            if name == args.workload:
                continue
            prec = "f32" if name == "c5" else "f64"
            try:
                r, _ = measure(ctx, name, prec, args.also_steps, 3, do_e2e=not args.no_e2e)
            except Exception as exc:      # noqa: BLE001 -- a failed side measurement must not lose the headline
                r = {"error": "{}: {}".format(type(exc).__name__, exc)}
            rf = r.pop("roofline", None)
            if rf:
                r.update({"frac_pipe": rf["frac_pipe"], "frac_alg": rf["frac_alg"], "hbm_frac": rf["hbm"]["frac"],
                          "hbm_gbs": rf["hbm"]["achieved_gbs"], "pipe": rf["bound"],
                          "executed_inst_per_member_step": rf["executed_inst_per_member_step"]})
            also[name] = r

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": head["value"], "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": precision, "data": "synthetic" if args.workload in ("c3", "c4a", "c4b") else
        "reference test catchment forcing (tests/golden fixture) + LHS parameter sets (seed 42 + rank)",
        "config": {"workload": "{}: {}".format(args.workload, head["label"]), "members_per_gpu": head["members_per_gpu"],
                   "steps_per_member": head["steps_per_member"], "report_gap": head["report_gap"],
                   "l2": "flushed between timed steps (256 MiB fill)", "sharding": "members over ranks, "
                   "one NCCL all-gather of the [members, 9] block the kernel wrote per step" if world > 1 else "single GPU",
                   "csrc_sha": ctx.hash},
        "roofline": head["roofline"],
        "clocks": clocks,
        "gpu_launches": head["gpu_launches"],
    }
    if "e2e" in head:
        line["e2e"] = head["e2e"]
    if also:
        line["also"] = also
    if world == 1 and not args.no_cpu_baseline:
        pool, cores = make_pool(w)
        v, secs, sample = cpu_throughput(w, pool, cores, per_worker=CPU_MEMBERS_PER_WORKER)
        pool.close()
        pool.join()
        line["cpu_baseline"] = {"value": v, "unit": METRIC, "cores": cores, "kind": "port", "sample": sample,
                                "seconds": secs,
                                "note": "C oracle (bit-identical restatement of the pure-Python reference, which "
                                        "itself measured 4.6e4 member-timesteps/s/core in the build container)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def reference_arm(args, rank, world):
    """The reference's CPU algorithm on all host cores (rank 0 only)."""
    if rank != 0:
        return
    w = make_workload(args.workload, 0, args.members, simulate=simulate_on_host)
    pool, cores = make_pool(w)
    times, sample = [], ""
    for i in range(args.warmup + args.steps):
        v, secs, sample = cpu_throughput(w, pool, cores, per_worker=CPU_MEMBERS_PER_WORKER)
        if i >= args.warmup:
            times.append((v, secs))
    pool.close()
    pool.join()
    value = float(np.mean([t[0] for t in times]))
    secs = float(np.mean([t[1] for t in times]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "same workload as the GPU arm",
        "config": {"workload": "{}: {}".format(args.workload, w["label"]), "step": "bounded sample: " + sample},
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
