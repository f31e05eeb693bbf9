#!/usr/bin/env python3
"""Cost of members outside the fast form's domain inside a C2-sized batch (1e5 members): pure LHS
batch against the same batch with 1 % / 5 % of the rows replaced by out-of-domain parameter sets.
Prints one JSON object.  """
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

WILD = np.array([
    [1.0, 0.5, 0.2, 0.3, 1.5, 60.0, 0.5, 2.0, 30.0, 0.2],
    [1.05, 0.9, 0.6, 0.5, 3.0, 20.0, 5.0, 10.0, 12.0, 3.0],
    [0.95, 0.1, 0.1, 0.9, 0.9, 100.0, 30.0, 30.0, 30.0, 30.0],
    [1.0, 0.5, 0.995, 1.0, 0.01, 50.0, 1.0, 48.0, 1200.0, 1.0],
])


def main():
    import torch
    from smartpy_b200.engine import BatchEngine
    w = bench.make_workload("c2", 0)
    eng = BatchEngine(w["rain"], w["peva"], w["area"], w["dt"], w["gap"], obs=w["obs"], extra=w["extra"],
                      warm_up_steps=w["warm_steps"], gw_constraint=w["gwc"])
    n = w["n_members"]
    out = {}
    rng = np.random.RandomState(1)
    for share in (0.0, 0.01, 0.05):
        params = w["params"].copy()
        k = int(n * share)
        if k:
            at = rng.choice(n, k, replace=False)
            params[at] = WILD[rng.randint(0, len(WILD), k)]
        p_dev = torch.from_numpy(params).cuda()
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
        out["wild_share_%g" % share] = {"ms": float(np.median(ms)), "members": n, "wild": k}
    base = out["wild_share_0"]["ms"]
    for v in out.values():
        v["vs_pure_fast"] = v["ms"] / base
    out["side_stream"] = os.environ.get("SMART_B200_NO_SIDE_STREAM") is None
    print(json.dumps(out))


if __name__ == "__main__":
    main()
