"""CPU, world_size 2, gloo: the N > 1 host path -- row sharding and the one all-gather of the
[rows, 8] score block + gw that follows the kernel (smartpy_b200/distributed.py)."""
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
    scores = rows[:, None] + torch.arange(8, dtype=torch.float64)[None, :] / 10
    gw = -rows
    full_scores, full_gw = du.all_gather_rows(scores, gw, n_rows)
    np.save(os.path.join(out_dir, "scores_%d.npy" % rank), full_scores.numpy())
    np.save(os.path.join(out_dir, "gw_%d.npy" % rank), full_gw.numpy())
    dist.destroy_process_group()


def test_shard_and_all_gather_world2(tmp_path):
    for n_rows in (7, 10):      # ragged and even shards
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_rows, str(tmp_path)), nprocs=2, join=True)
        expect = np.arange(n_rows, dtype=np.float64)[:, None] + np.arange(8)[None, :] / 10
        for rank in range(2):
            assert np.array_equal(np.load(tmp_path / ("scores_%d.npy" % rank)), expect)
            assert np.array_equal(np.load(tmp_path / ("gw_%d.npy" % rank)), -np.arange(n_rows, dtype=np.float64))


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
