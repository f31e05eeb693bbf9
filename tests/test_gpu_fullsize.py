"""GPU, BASELINE.json's full sizes: the batches bench.py times (C2: 1e5 members x 96,432 steps;
C3: 1.25e6 members x 271,752 steps per GPU; C5: C2 in FP32 state) are too large for the oracle,
so parity is carried by size-independent properties:

  * members with reference-generated answers (tests/golden) planted anywhere in the batch --
    first, last, either side of warp and CTA boundaries, the ragged tail -- come out with the SAME
    BITS as in a 40-member run (which tests/test_gpu_parity.py pins to the reference within
    1e-10), although the two launches use different instantiations of the kernel (register
    budget chosen from the batch size): binary64 arithmetic is compiled without implicit FMA
    contraction for exactly this reason (smart_kernels.cu);
  * any row range computed on its own (what a rank of a sharded run does) is the same bits;
  * the same launch twice is the same bits;
  * C3: planted members agree with the oracle on the 30-year forcing.
"""
import numpy as np
import pytest

from conftest import load_golden
from test_gpu_parity import make_engine, relmax, ATOL_F32

pytestmark = pytest.mark.gpu


def _merged_form_ok(p, dt=3600.0):
    """The kernel's own rule (fast_form_ok, smart_kernels.cu): a member qualifies for the merged
    ("fast") step when no clamp, cap or leak predicate of the reference can fire."""
    T, C, H, D, S, Z, SK, FK, GK, RK = p
    return bool(min(SK, FK, GK, RK) * 3600.0 >= dt and 0.0 <= S <= 0.5 and Z > 0.0 and 0.0 <= D <= 1.0
                and 0.0 <= H <= 0.99 and T > 0.0)


def _planted_batch(n, planted, seed):
    import bench
    params = bench.lhs_rows(n, seed)
    k = len(planted)
    # positions: the ends, around warp (32) and CTA (64 / 128) boundaries, the middle, the ragged tail
    spots = [0, 1, 31, 32, 63, 64, 127, 128, n // 2, n // 2 + 1, n - 129, n - 65, n - 33, n - 2, n - 1]
    rng = np.random.RandomState(seed)
    spots = spots + sorted(rng.choice(np.arange(200, n - 200), max(k - len(spots), 0), replace=False).tolist())
    spots = np.array(spots[:k])
    params[spots] = planted
    return params, spots


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_c2_full_batch_planted_members_and_row_ranges(catchment, precision):
    import torch
    from oracle import scores as oscores
    g = load_golden("runs_members")
    # (every golden member lies inside the merged form's domain, like the LHS rows around them, so
    # all of these CTAs run the same step variant; members OUTSIDE it take their CTA to the
    # branch-faithful kernel: test_gpu_parity.py::test_general_kernel_on_out_of_range_members)
    assert all(_merged_form_ok(p) for p in g["params"])
    planted, q_ref, gw_ref = g["params"], g["q"], g["gw"]
    n = 100000 - 37                                     # C2's size, with a ragged last CTA
    params, spots = _planted_batch(n, planted, 5)
    assert all(_merged_form_ok(p) for p in params[::997])
    eng = make_engine(catchment, precision=precision)
    small = eng.run(planted, discharge=False, scores=True, gw=True)
    sc_small, gw_small = small["scores"].cpu().numpy(), small["gw"].cpu().numpy()
    full = eng.run(params, discharge=False, scores=True, gw=True)
    sc, gw = full["scores"].cpu().numpy(), full["gw"].cpu().numpy()
    assert sc.shape == (n, 8) and np.isfinite(sc[:, :7]).all()
    # planted members: bit-identical to the launch of their own, whatever their neighbours
    if precision == "f64":
        assert np.array_equal(sc[spots], sc_small) and np.array_equal(gw[spots], gw_small)
    else:   # binary32 state keeps the compiler's contraction: same kernel here, so same bits too
        assert np.allclose(sc[spots][:, :7], sc_small[:, :7], rtol=0, atol=ATOL_F32)
    # ... and within the bar of the reference's own answers
    sc_ref = oscores.score_members(q_ref, gw_ref, catchment.obs, 0.12667)
    if precision == "f64":
        assert np.max(np.abs(sc[spots][:, :7] - sc_ref[:, :7]) / np.maximum(1.0, np.abs(sc_ref[:, :7]))) < 1e-10
        assert relmax(gw[spots], gw_ref) < 1e-9
    else:
        assert np.max(np.abs(sc[spots][:, :2] - sc_ref[:, :2])) < ATOL_F32      # NSE, KGE
    # the same launch again: same bits
    again = eng.run(params, discharge=False, scores=True, gw=True)
    assert np.array_equal(again["scores"].cpu().numpy(), sc, equal_nan=True)
    # row ranges on their own (the shards of an 8-GPU run): same bits as the slice
    from smartpy_b200.distributed import shard_bounds
    for rank in (0, 3, 7):
        lo, hi = shard_bounds(n, rank, 8)
        part = eng.run(params[lo:hi], discharge=False, scores=True, gw=True)
        assert np.array_equal(part["scores"].cpu().numpy(), sc[lo:hi], equal_nan=True)
        assert np.array_equal(part["gw"].cpu().numpy(), gw[lo:hi])
    # best member of the whole batch = arg-max of the table (first index on ties)
    best = eng.run(params, discharge=False, scores=True, best=("NSE", 1))
    torch.cuda.synchronize()
    assert int(best["best"][1].item()) == int(np.argmax(sc[:, 0])) and float(best["best"][0].item()) == sc[:, 0].max()


def test_c3_full_batch_planted_members_match_oracle(oracle_lib):
    """The per-GPU batch of the BASELINE target run: 1.25e6 members x (262,992 + 8,760) steps."""
    import bench
    from smartpy_b200.engine import BatchEngine
    from oracle import scores as oscores
    n = 1250000
    w = bench.make_workload("c3", 0, members=n)
    params = w["params"]
    spots = np.array([0, 63, 64, 127, 128, n // 2, n - 129, n - 1])
    eng = BatchEngine(w["rain"], w["peva"], w["area"], w["dt"], w["gap"], obs=w["obs"], extra=w["extra"],
                      warm_up_steps=w["warm_steps"], gw_constraint=w["gwc"])
    res = eng.run(params, discharge=False, scores=True, gw=True)
    sc, gw = res["scores"].cpu().numpy(), res["gw"].cpu().numpy()
    assert np.isfinite(sc[:, :7]).all()
    q_ref, gw_ref = oracle_lib.run_members(w["area"], w["dt"], w["rain"], w["peva"], params[spots], w["extra"],
                                           w["n_steps"], w["gap"], warm_up=365)
    sc_ref = oscores.score_members(q_ref, gw_ref, w["obs"], w["gwc"])
    assert np.max(np.abs(sc[spots][:, :7] - sc_ref[:, :7]) / np.maximum(1.0, np.abs(sc_ref[:, :7]))) < 1e-10
    assert np.array_equal(sc[spots][:, 7], sc_ref[:, 7])
    assert relmax(gw[spots], gw_ref) < 1e-9


@pytest.mark.parametrize("flags", [0, 0x10000])        # block mode, per-step path
def test_member_order_changes_no_bit(catchment, flags):
    """BatchEngine sorts the members of a launch by T on the device (member_order of the C ABI) so
    that the lanes of a warp agree on wet and dry steps; FLAG_NO_REORDER keeps the sample order.
    Same scores, gw and best member, bit for bit; and the C ABI refuses an order it cannot honour."""
    import ctypes
    import torch
    import bench
    from smartpy_b200 import _native
    from smartpy_b200.engine import FLAG_NO_REORDER, REORDER_MIN_MEMBERS
    g = load_golden("runs_members")
    n = 3 * REORDER_MIN_MEMBERS + 17
    params, spots = _planted_batch(n, g["params"], 9)
    params[5, 0] = params[6, 0] = params[n - 1, 0]              # ties on the sort key
    n_steps = 24 * 500
    a = make_engine(catchment, n_steps=n_steps, flags=flags)
    b = make_engine(catchment, n_steps=n_steps, flags=flags | FLAG_NO_REORDER)
    c0 = a.kernel_launches                                       # the library's own counter (global)
    ra = a.run(params, discharge=False, scores=True, gw=True, best=("KGE", 1))
    c1 = a.kernel_launches
    rb = b.run(params, discharge=False, scores=True, gw=True, best=("KGE", 1))
    c2 = a.kernel_launches
    assert c1 - c0 > c2 - c1 == 3                                # fast + general + best finalize; the sort ran only in `a`
    assert np.array_equal(ra["scores"].cpu().numpy(), rb["scores"].cpu().numpy(), equal_nan=True)
    assert np.array_equal(ra["gw"].cpu().numpy(), rb["gw"].cpu().numpy())
    assert int(ra["best"][1].item()) == int(rb["best"][1].item())
    assert float(ra["best"][0].item()) == float(rb["best"][0].item())
    # discharge wanted: the order is not used (the [t][member] stores must stay coalesced)
    before = a.kernel_launches
    a.run(params[:REORDER_MIN_MEMBERS], discharge=True, scores=False)
    assert a.kernel_launches - before == 2
    # the ABI itself: member_order with a discharge buffer is refused before any launch
    d = a._desc(8)
    d.params = d.rain = d.peva = d.area_m2 = 8
    d.member_order = 8
    d.discharge, d.ld_discharge = 8, 8
    assert a.lib.smart_batch_run_f64(ctypes.byref(d), None) == _native.ERR_BAD_ARG
    assert "member_order" in _native.last_error()

