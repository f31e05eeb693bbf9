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
    print("bars (BASELINE.json north_star): FP64 discharge 1e-10 relative; FP32 NSE/KGE 1e-5 absolute")


if __name__ == "__main__":
    main()
