"""GPU, needs >= 2 devices (skipped otherwise): the Monte-Carlo layer sharded over two ranks with
NCCL -- rows split with distributed.shard_bounds, one all-gather of the score block -- must give
exactly the single-GPU result, and rank 0 writes the same database.  The ranks do NOT seed their
generators alike: the run broadcasts rank 0's sample (each rank draws its own at construction)."""
import os
import socket

import numpy as np
import pytest

from conftest import EXTRA
from test_host_logic import catchment_dir  # noqa: F401

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, root, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from smartpy_b200 import montecarlo
    np.random.seed(11 + 100 * rank)                      # rank 0 as the single-GPU run, rank 1 something else
    setup = montecarlo.LHS('Catchment', root, 'csv', 'csv', sample_size=301, parallel='mpi', save_sim=True)
    setup.model.extra = dict(EXTRA)
    setup.db_file = os.path.join(out_dir, "sharded.lhs")
    setup.run()
    np.save(os.path.join(out_dir, "sample_%d.npy" % rank), setup.sample_params)
    np.save(os.path.join(out_dir, "scores_%d.npy" % rank), setup.results['scores'].cpu().numpy())
    # best-member-only run: no score table, one all-gather of `world` (score, row) pairs
    from smartpy_b200 import distributed as du
    lo, hi = du.shard_bounds(setup.sample_params.shape[0], rank, world)
    engine = setup.model.get_engine(report='summary', gw_constraint=setup.constraints['gw'])
    res = engine.run(setup.sample_params[lo:hi], discharge=False, scores=False, gw=False, best=('KGE', 1))
    score, row = du.all_gather_best(res['best'][0], res['best'][1], lo, 1)
    np.save(os.path.join(out_dir, "best_%d.npy" % rank), np.array([float(score[0]), float(row[0])]))
    dist.destroy_process_group()


def test_lhs_run_sharded_over_two_gpus(catchment_dir, tmp_path):  # noqa: F811
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from smartpy_b200 import montecarlo
    mp.spawn(_worker, args=(2, _free_port(), catchment_dir, str(tmp_path)), nprocs=2, join=True)
    np.random.seed(11)
    single = montecarlo.LHS('Catchment', catchment_dir, 'csv', 'csv', sample_size=301, save_sim=True)
    single.model.extra = dict(EXTRA)
    single.run()
    ref = single.results['scores'].cpu().numpy()
    for rank in range(2):
        got = np.load(tmp_path / ("scores_%d.npy" % rank))
        assert np.array_equal(got, ref, equal_nan=True)
        assert np.array_equal(np.load(tmp_path / ("sample_%d.npy" % rank)), single.sample_params)
    with open(single.db_file) as a, open(tmp_path / "sharded.lhs") as b:
        assert a.read() == b.read()
    kge = ref[:, 1]
    for rank in range(2):
        score, row = np.load(tmp_path / ("best_%d.npy" % rank))
        assert int(row) == int(np.nanargmax(kge)) and score == kge[int(row)]
