#!/usr/bin/env python3
"""Where does a C2-sized launch (1e5 members, one wave) lose time?  Times the C2 step with
(a) the LHS sample, (b) 1e5 copies of ONE member (every warp costs the same: no load imbalance
between SMs / sub-partitions left, only the 5.28-warps-per-sub-partition quantisation),
(c) exactly 94,720 = 148 x 4 x 5 x 32 members of each kind.  One JSON object."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from smartpy_b200.engine import BatchEngine
    w = bench.make_workload("c2", 0)
    eng = BatchEngine(w["rain"], w["peva"], w["area"], w["dt"], w["gap"], obs=w["obs"], extra=w["extra"],
                      warm_up_steps=w["warm_steps"], gw_constraint=w["gwc"])
    out = {}
    med = w["params"][np.argsort(w["params"][:, 4] * w["params"][:, 5])[50000]]
    for label, params in (("lhs_100000", w["params"]), ("identical_100000", np.tile(med, (100000, 1))),
                          ("lhs_94720", w["params"][:94720]), ("identical_94720", np.tile(med, (94720, 1))),
                          ("lhs_113664", bench.lhs_rows(113664, 42)), ("identical_113664", np.tile(med, (113664, 1)))):
        p_dev = torch.from_numpy(np.ascontiguousarray(params)).cuda()
        for _ in range(3):
            eng.run(p_dev)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(p_dev)
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        n = len(params)
        out[label] = {"ms": float(np.median(ms)), "ns_per_member_step": float(np.median(ms)) * 1e6 / (n * 96432)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
