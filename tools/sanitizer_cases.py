#!/usr/bin/env python3
"""Small runs of every kernel path, meant to be executed under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitizer_cases.py
    compute-sanitizer --tool racecheck python tools/sanitizer_cases.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import Catchment, load_golden, EXTRA  # noqa: E402
from smartpy_b200.engine import BatchEngine  # noqa: E402


def main():
    import torch
    c = Catchment()
    g = load_golden("runs_members")
    days = 40
    n = days * 24
    obs = c.obs[:days]
    params = np.resize(g["params"], (150, 10))
    cases = []
    for flags in (0, 0x10000, 0x10002, 1):
        for precision in ('f64', 'f32'):
            eng = BatchEngine(c.rain[:n], c.peva[:n], c.area, 3600.0, 24, obs=obs, extra=EXTRA, warm_up_steps=24 * 5,
                              gw_constraint=0.12667, precision=precision, flags=flags)
            r = eng.run(params, discharge=True, scores=True, gw=True, best=('NSE', 1))
            cases.append(("flags=%#x %s" % (flags, precision), float(r["scores"][:, 0].max())))
    # odd lengths, raw reporting, last_state / initial_state (fluxes kernel)
    eng = BatchEngine(c.rain[7:7 + 777], c.peva[7:7 + 777], c.area, 3600.0, 7, extra=EXTRA, warm_up_steps=49, report='raw')
    r = eng.run(params[:5], discharge=True, scores=False, last_state=True)
    eng.run(params[:5], discharge=True, scores=False, initial_state=r["last_state"].cpu().numpy())
    # multi-catchment tiles (cp.async path), per-step and block mode
    rain = np.stack([c.rain[:n] * k for k in (1.0, 0.7, 1.3)], 1)
    peva = np.stack([c.peva[:n]] * 3, 1)
    for mpc in (50, 7):
        eng = BatchEngine(rain, peva, [1e8, 2e8, 3e8], 3600.0, 24, extra=EXTRA, members_per_catchment=mpc)
        eng.run(params[:3 * mpc], discharge=True, scores=False)
        eng = BatchEngine(rain[5:5 + 24 * 30], peva[5:5 + 24 * 30], [1e8, 2e8, 3e8], 3600.0, 1, report='raw',
                          members_per_catchment=mpc)
        eng.run(params[:3 * mpc], discharge=True, scores=False)
    # block-sub mode (reports inside a constant-forcing block), single and multi catchment, both precisions
    for gap, report in ((1, 'raw'), (6, 'summary'), (24, 'raw')):
        for precision in ('f64', 'f32'):
            eng = BatchEngine(c.rain[:n:24] * 24.0, c.peva[:n:24] * 24.0, c.area, 3600.0, gap, extra=EXTRA,
                              warm_up_steps=24 * 5, report=report, precision=precision, forcing_repeat=24)
            r = eng.run(params, discharge=True, scores=False, gw=True)
            cases.append(("block-sub gap=%d %s %s" % (gap, report, precision), float(r["gw"].sum())))
    eng = BatchEngine(rain[::24] * 24.0, peva[::24] * 24.0, [1e8, 2e8, 3e8], 3600.0, 1, extra=EXTRA, report='raw',
                      members_per_catchment=50, forcing_repeat=24)
    eng.run(params[:150], discharge=True, scores=False)
    # member grouping by the library (range/keys/sort/emit kernels), members outside the fast form in
    # CTAs of their own on the side stream, scores + gw into a [N, 9] block, best member over idle CTAs
    big = np.resize(g["params"], (4500, 10)).copy()
    big[::97, 4] = 1.5                         # S > 0.5: needs the branch-faithful form
    eng = BatchEngine(c.rain[:n], c.peva[:n], c.area, 3600.0, 24, obs=obs, extra=EXTRA, warm_up_steps=24 * 5,
                      gw_constraint=0.12667)
    blk = torch.empty((4500, 9), dtype=torch.float64, device='cuda')
    r = eng.run(big, scores=True, gw=True, best=('KGE', 1), out={'block': blk})
    cases.append(("grouped batch with wild members", float(r["best"][0])))
    cases.append(("run_host", float(eng.run_host(big[:300])["gw"].sum())))
    # relay: the timeline in segments, state parked in the launch workspace between two CTAs (tickets,
    # progress counters, acquire/release hand-over); block and per-step mode, both precisions, the two
    # step forms side by side, best member
    from smartpy_b200 import _native
    for flags in (_native.flag_relay_segs(3), _native.flag_relay_segs(5) | 0x10000):
        for precision in ('f64', 'f32'):
            eng = BatchEngine(c.rain[:n], c.peva[:n], c.area, 3600.0, 24, obs=obs, extra=EXTRA, warm_up_steps=24 * 5,
                              gw_constraint=0.12667, precision=precision, flags=flags)
            r = eng.run(big, scores=True, gw=True, best=('NSE', 1))
            cases.append(("relay flags=%#x %s" % (flags, precision), float(r["best"][0])))
            r = eng.run(params, discharge=True, scores=True, gw=True)
            cases.append(("relay + discharge flags=%#x %s" % (flags, precision), float(r["gw"].sum())))
    # conditioning of a score table (mask + ordered compaction, radix select + bitonic sort of the
    # winners: shared-memory histograms, atomics, multi-chunk sort) and the device sampler
    from smartpy_b200.montecarlo import conditioning
    names = ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW']
    rng = np.random.RandomState(1)
    table = rng.randn(9001, 8)
    table[:, 1] = np.round(table[:, 1], 1)
    table[:, 7] = rng.rand(9001) > 0.5
    t = torch.from_numpy(table).cuda()
    for k in (1, 100, 2049, 4100):
        rows = conditioning.best_rows(t, names, 'KGE', k, {'GW': ('equal', (1.0,))})
        cases.append(("best_rows k=%d" % k, float(rows[-1])))
    rows = conditioning.behavioural_rows(t, names, {'NSE': ('min', (0.0,)), 'PBias': ('inside', (-1.0, 1.0))})
    cases.append(("behavioural_rows", float(rows.numel())))
    sample = conditioning.latin_hypercube_device(5003, [[0.0, 1.0]] * 10, seed=3, row_first=11, n_rows=4000)
    cases.append(("latin_hypercube_device", float(sample.sum())))
    torch.cuda.synchronize()
    for name, v in cases:
        print(name, v)
    print("sanitizer cases done")


if __name__ == "__main__":
    main()
