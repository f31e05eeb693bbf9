"""Times the conditioning kernels (smart_best_rows / smart_condition_rows) on a synthetic score
table, next to the torch operations they replaced (nonzero + topk) and numpy's argsort, which is
what the reference does on the host (best.py:287).  Usage: python tools/bench_conditioning.py [N]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from smartpy_b200.montecarlo import conditioning  # noqa: E402

NAMES = ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW']


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = float('inf')
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    g = torch.Generator(device='cuda').manual_seed(1)
    scores = torch.randn((n, 8), dtype=torch.float64, device='cuda', generator=g)
    scores[:, 7] = (scores[:, 7] > 0).to(torch.float64)
    out = {"n_rows": n}
    cons = {'GW': ('equal', (1.0,))}
    for k in (100, 10_000, 1_000_000):
        if k > n:
            continue
        out["best_rows_k%d_ms" % k] = timed(lambda: conditioning.best_rows(scores, NAMES, 'NSE', k, cons))

        def torch_way():
            kept = torch.nonzero(scores[:, 7] == 1.0)[:, 0]
            top = torch.topk(scores[kept, 0], k, largest=True, sorted=True)
            return kept[torch.flip(top.indices, dims=[0])]
        out["torch_nonzero_topk_k%d_ms" % k] = timed(torch_way)
    cond = {'NSE': ('min', (0.5,)), 'PBias': ('inside', (-1.0, 1.0))}
    out["behavioural_rows_ms"] = timed(lambda: conditioning.behavioural_rows(scores, NAMES, cond))
    out["torch_mask_nonzero_ms"] = timed(lambda: torch.nonzero(
        (scores[:, 0] >= 0.5) & (scores[:, 5] >= -1.0) & (scores[:, 5] <= 1.0))[:, 0])
    host = scores[:, 0].cpu().numpy()
    t0 = time.perf_counter()
    np.argsort(host)
    out["numpy_argsort_host_ms"] = (time.perf_counter() - t0) * 1e3
    print(json.dumps(out))


if __name__ == '__main__':
    main()
