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


def _backend_device(like=None):
    """The device collectives of the current backend work on: the caller's CUDA device under
    NCCL, the host under gloo."""
    import torch
    dist = _dist()
    if dist.get_backend() == 'nccl':
        if like is not None and like.is_cuda:
            return like.device
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def broadcast_rows(rows, src=0):
    """The [N, k] float64 array of rank `src`, on every rank (a numpy array in, a numpy array out).
    Used to give a sharded Monte Carlo run ONE sample: each rank draws its own from numpy's global
    generator at construction, rank 0's is the one that counts."""
    import numpy as np
    import torch
    dist = _dist()
    rank, world = rank_world()
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    if world == 1:
        return rows
    dev = _backend_device()
    shape = torch.tensor(list(rows.shape), dtype=torch.int64, device=dev)
    dist.broadcast(shape, src=src)
    buf = torch.from_numpy(rows).to(dev) if rank == src else torch.empty(tuple(shape.tolist()), dtype=torch.float64, device=dev)
    dist.broadcast(buf, src=src)
    return buf.cpu().numpy()


def all_gather_rows(block, n_rows):
    """All-gather the per-rank row blocks [n_r, w] (rank r holds the rows shard_bounds(n_rows, r,
    world)) into the full [n_rows, w] table on every rank; ragged shards are padded to the largest
    block.  One collective (NCCL all-gather over NVLink on the GPU box)."""
    import torch
    dist = _dist()
    rank, world = rank_world()
    if world == 1:
        return block
    dev = _backend_device(block)
    block = block.to(dev)
    biggest = shard_bounds(n_rows, 0, world)[1]
    width = block.shape[1]
    if block.shape[0] == biggest and block.is_contiguous():
        padded = block
    else:
        padded = torch.full((biggest, width), float('nan'), dtype=block.dtype, device=dev)
        padded[:block.shape[0]] = block
    gathered = torch.empty((world * biggest, width), dtype=block.dtype, device=dev)
    dist.all_gather_into_tensor(gathered, padded)
    if n_rows == world * biggest:
        return gathered
    gathered = gathered.view(world, biggest, width)
    return torch.cat([gathered[r, :shard_bounds(n_rows, r, world)[1] - shard_bounds(n_rows, r, world)[0]]
                      for r in range(world)])


def send_rows_to_first(rows, owner, n_rows):
    """Collective: the [n_rows, w] block held by rank `owner` arrives on rank 0 (returned there;
    other ranks get None).  Every rank calls it with its own block; only `owner`'s is used."""
    import torch
    dist = _dist()
    rank, world = rank_world()
    if world == 1 or owner == 0:
        return rows if rank == 0 else None
    dev = _backend_device(rows)
    width = int(rows.shape[1])
    if rank == owner:
        dist.send(rows.to(dev).contiguous(), dst=0)
        return None
    if rank == 0:
        buf = torch.empty((n_rows, width), dtype=rows.dtype, device=dev)
        dist.recv(buf, src=owner)
        return buf
    return None


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
