"""CPU, world_size 2, gloo: the N > 1 host path -- row sharding, the one all-gather of the
[rows, 9] block (8 scores + gw) that follows the kernel, the broadcast that gives every rank of a
sharded Monte Carlo run the SAME sample, and the hand-over of simulated series to the rank that
writes the database (smartpy_b200/distributed.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smartpy_b200 import distributed as du
    assert du.rank_world() == (rank, world)
    lo, hi = du.shard_bounds(n_rows, rank, world)
    # each rank "computes" its shard: score[row, k] = row + k / 10, gw[row] = -row
    rows = torch.arange(lo, hi, dtype=torch.float64)
    block = torch.empty((hi - lo, 9), dtype=torch.float64)
    block[:, :8] = rows[:, None] + torch.arange(8, dtype=torch.float64)[None, :] / 10
    block[:, 8] = -rows
    full = du.all_gather_rows(block, n_rows)
    assert full.shape == (n_rows, 9)
    np.save(os.path.join(out_dir, "scores_%d.npy" % rank), full[:, :8].numpy())
    np.save(os.path.join(out_dir, "gw_%d.npy" % rank), full[:, 8].numpy())
    # every rank draws its OWN sample (unseeded generator); after the broadcast all hold rank 0's
    mine = np.random.RandomState(1000 + rank).rand(n_rows, 10)
    same = du.broadcast_rows(mine)
    np.save(os.path.join(out_dir, "sample_%d.npy" % rank), same)
    assert (rank == 0) == np.array_equal(same, mine)
    # simulated series travel to the writer rank shard by shard, in row order
    sims = rows[:, None] * torch.ones((1, 5), dtype=torch.float64)
    got = []
    for r in range(world):
        r_lo, r_hi = du.shard_bounds(n_rows, r, world)
        part = du.send_rows_to_first(sims, r, r_hi - r_lo)
        assert (part is not None) == (rank == 0)
        if rank == 0:
            got.append(part)
    if rank == 0:
        np.save(os.path.join(out_dir, "sims.npy"), torch.cat(got).numpy())
    # an empty shard (more ranks than rows) must not break the collective
    lo1, hi1 = du.shard_bounds(1, rank, world)
    tiny = du.all_gather_rows(torch.full((hi1 - lo1, 9), 7.0, dtype=torch.float64), 1)
    assert tiny.shape == (1, 9) and bool((tiny == 7.0).all())
    dist.destroy_process_group()


def test_shard_and_all_gather_world2(tmp_path):
    for n_rows in (7, 10):      # ragged and even shards
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_rows, str(tmp_path)), nprocs=2, join=True)
        expect = np.arange(n_rows, dtype=np.float64)[:, None] + np.arange(8)[None, :] / 10
        for rank in range(2):
            assert np.array_equal(np.load(tmp_path / ("scores_%d.npy" % rank)), expect)
            assert np.array_equal(np.load(tmp_path / ("gw_%d.npy" % rank)), -np.arange(n_rows, dtype=np.float64))
            assert np.array_equal(np.load(tmp_path / ("sample_%d.npy" % rank)),
                                  np.random.RandomState(1000).rand(n_rows, 10))
        assert np.array_equal(np.load(tmp_path / "sims.npy"),
                              np.arange(n_rows, dtype=np.float64)[:, None] * np.ones((1, 5)))


def _best_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smartpy_b200 import distributed as du
    # (local best score, local row) per rank and case; first_row = 100 * rank
    cases = [((0.3, 4), (0.7, 2)), ((0.7, 9), (0.7, 1)), ((float('nan'), 0), (-5.0, 3)), ((0.1, 1), (0.2, 2))]
    got = []
    for k, case in enumerate(cases):
        score, row = case[rank]
        sign = -1 if k == 3 else 1
        s, r = du.all_gather_best(torch.tensor([score], dtype=torch.float64), torch.tensor([row]), 100 * rank, sign)
        got.append((float(s[0]), int(r[0])))
    np.save(os.path.join(out_dir, "best_%d.npy" % rank), np.array(got))
    dist.destroy_process_group()


def test_best_member_pairs_world2(tmp_path):
    mp.spawn(_best_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    expect = np.array([(0.7, 102), (0.7, 9), (-5.0, 103), (0.1, 1)])
    for rank in range(2):
        assert np.array_equal(np.load(tmp_path / ("best_%d.npy" % rank)), expect)
