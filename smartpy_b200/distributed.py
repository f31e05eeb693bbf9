"""Member sharding over the GPUs of one box.

Members (parameter set x catchment) never interact, so the path shards trivially: rank r
owns the contiguous rows [lo, hi) of the sample, forcing and observations are replicated
(megabytes), and the only exchange is one all-gather of the [N/G, 8] score block (+ gw) after
the kernel -- NCCL over NVLink on the GPU box, gloo in the CPU tests.  This replaces the
reference's spotpy/mpi4py master-worker farming of single simulations
(smartpy/montecarlo/montecarlo.py:62-63, :153).
"""


def _dist():
    import torch.distributed as dist
    return dist


def rank_world():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def barrier():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def shard_bounds(n_rows, rank, world):
    """Contiguous, balanced row block of `rank`: sizes differ by at most one, earlier ranks larger."""
    base, extra = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_rows(scores, gw, n_rows):
    """All-gather the per-rank [n_r, 8] scores and [n_r] gw blocks into [n_rows, 8] / [n_rows]
    on every rank (ragged shards are padded to the largest block)."""
    import torch
    dist = _dist()
    rank, world = rank_world()
    if world == 1:
        return scores, gw
    biggest = shard_bounds(n_rows, 0, world)[1]
    width = scores.shape[1] + 1
    block = torch.full((biggest, width), float('nan'), dtype=torch.float64, device=scores.device)
    block[:scores.shape[0], :-1] = scores
    block[:gw.shape[0], -1] = gw
    gathered = torch.empty((world * biggest, width), dtype=torch.float64, device=scores.device)
    dist.all_gather_into_tensor(gathered, block)
    gathered = gathered.view(world, biggest, width)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_rows, r, world)
        parts.append(gathered[r, :hi - lo])
    full = torch.cat(parts)
    return full[:, :-1].contiguous(), full[:, -1].contiguous()


def all_gather_best(best_score, best_index, first_row, sign=1):
    """Best-member-only runs (no score table wanted): every rank hands in the (score, local row)
    pair its kernel selected (``BatchEngine.run(best=...)``); ONE all-gather of world pairs, then
    the arg-max (sign > 0) or arg-min over them on every rank.  Returns tensors (score [1] float64,
    global row [1] int64) on the input's device.  A NaN score never wins against a number; ties
    go to the lowest global row."""
    import torch
    dist = _dist()
    rank, world = rank_world()
    score = best_score.reshape(1).to(torch.float64)
    row = best_index.reshape(1).to(torch.int64) + int(first_row)
    if world == 1:
        return score, row
    pair = torch.stack([score[0], row[0].to(torch.float64)])       # rows < 2^53: exact in float64
    flat = torch.empty((world * 2,), dtype=torch.float64, device=pair.device)
    dist.all_gather_into_tensor(flat, pair.contiguous())
    pairs = flat.view(world, 2)
    key = pairs[:, 0] if sign > 0 else -pairs[:, 0]
    key = torch.where(torch.isnan(key), torch.full_like(key, float('-inf')), key)
    winners = key == key.max()
    rows = torch.where(winners, pairs[:, 1], torch.full_like(pairs[:, 1], float('inf')))
    who = torch.argmin(rows)
    return pairs[who, 0].reshape(1), pairs[who, 1].to(torch.int64).reshape(1)
