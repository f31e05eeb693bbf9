#!/usr/bin/env python3
"""Measured parity of the CUDA path against the reference-generated goldens (run on the GPU box):

    python tests/parity_report.py > profiles/rNN_parity_report.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import Catchment, load_golden, EXTRA  # noqa: E402
from smartpy_b200.engine import BatchEngine, warm_up_length  # noqa: E402
from oracle import scores as oscores  # noqa: E402


def main():
    c = Catchment()
    g = load_golden("runs_members")
    ref_sc = oscores.score_members(g["q"], g["gw"], c.obs, 0.12667)
    print("40 members (LHS seed 42, range corners, .lhs rows) x 87,672 + 8,760 hourly steps, test catchment")
    print("%-44s %12s %12s %12s %12s" % ("path", "max rel Q", "max rel gw", "max |dNSE|", "max |dKGE|"))
    for label, kw in (("FP64 block mode (default)", dict()),
                      ("FP64 per-step fast path", dict(flags=0x10000)),
                      ("FP64 per-step, no TMA", dict(flags=0x10002)),
                      ("FP64 general (branch-faithful) kernel", dict(flags=0x10001)),
                      ("FP32 state, block mode", dict(precision='f32')),
                      ("FP32 state, per-step", dict(precision='f32', flags=0x10000))):
        eng = BatchEngine(c.rain, c.peva, c.area, c.dt, c.gap, obs=c.obs, extra=EXTRA,
                          warm_up_steps=warm_up_length(365, c.dt), gw_constraint=0.12667, **kw)
        res = eng.run(g["params"], discharge=True, scores=True, gw=True)
        q = res["discharge"].double().cpu().numpy().T
        sc = res["scores"].cpu().numpy()
        gw = res["gw"].cpu().numpy()
        print("%-44s %12.3e %12.3e %12.3e %12.3e" % (
            label, np.max(np.abs(q - g["q"]) / g["q"]), np.max(np.abs(gw - g["gw"]) / g["gw"]),
            np.max(np.abs(sc[:, 0] - ref_sc[:, 0])), np.max(np.abs(sc[:, 1] - ref_sc[:, 1]))))
    # reports inside the constant-forcing day (block-sub mode): hourly values against the oracle
    import oracle
    days = 400
    n = days * 24
    for label, gap, report in (("FP64 block-sub, hourly output (gap 1)", 1, 'raw'), ("FP64 block-sub, 6-hourly means", 6, 'summary')):
        eng = BatchEngine(c.rain[:n:24] * 24.0, c.peva[:n:24] * 24.0, c.area, c.dt, gap, extra=EXTRA,
                          warm_up_steps=warm_up_length(30, c.dt), report=report, forcing_repeat=24)
        res = eng.run(g["params"], discharge=True, scores=False, gw=True)
        q = res["discharge"].cpu().numpy().T
        gw = res["gw"].cpu().numpy()
        q_ref, gw_ref = oracle.run_members(c.area, c.dt, c.rain[:n], c.peva[:n], g["params"], EXTRA, n, gap, report=report,
                                           warm_up=30)
        print("%-44s %12.3e %12.3e %12s %12s" % (label, np.max(np.abs(q - q_ref) / q_ref),
                                                 np.max(np.abs(gw - gw_ref) / gw_ref), "-", "-"))
    print("bars (BASELINE.json north_star): FP64 discharge 1e-10 relative; FP32 NSE/KGE 1e-5 absolute")


if __name__ == "__main__":
    main()
