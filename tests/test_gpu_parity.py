"""GPU parity tests: the CUDA path (through the C ABI, include/smart_b200.h) against
(a) outputs of the reference itself (tests/golden/*.npz) and (b) the CPU oracle on the same
seeded inputs.

Tolerances (BASELINE.json north_star): FP64 discharge within 1e-10 relative of the
reference; FP32 mode NSE/KGE within 1e-5 absolute.
"""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, EXTRA

pytestmark = pytest.mark.gpu

RTOL_Q = 1e-10      # north_star: FP64 discharge within 1e-10 relative
ATOL_F32 = 1e-5     # north_star: FP32 NSE/KGE within 1e-5 absolute


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den))


def make_engine(c, report='summary', warm_up_days=365, extra=EXTRA, obs=True, gwc=0.12667, precision='f64',
                flags=0, n_steps=None):
    from smartpy_b200.engine import BatchEngine, warm_up_length
    n = c.n_steps if n_steps is None else n_steps
    return BatchEngine(c.rain[:n], c.peva[:n], c.area, c.dt, c.gap, obs=c.obs[:n // c.gap] if obs else None,
                       extra=extra, warm_up_steps=warm_up_length(warm_up_days, c.dt) if warm_up_days else 0,
                       report=report, gw_constraint=gwc, precision=precision, flags=flags)


# ---------------------------------------------------------------- C1: the reference's own test configuration
@pytest.mark.parametrize("tag,kw", [
    ("summary", {}),
    ("raw", dict(report='raw')),
    ("nowarm", dict(warm_up_days=0)),
    ("noextra", dict(extra=None)),
    ("nowarm_noextra", dict(warm_up_days=0, extra=None)),
    ("warm30_raw", dict(warm_up_days=30, report='raw')),
])
def test_single_run_matches_reference(catchment, tag, kw):
    _torch()
    g = load_golden("runs_single")
    eng = make_engine(catchment, obs=False, **kw)
    res = eng.run(g["p_test"][None, :], discharge=True, scores=False, gw=True)
    q = res["discharge"][:, 0].cpu().numpy()
    assert q.shape == g["q_" + tag].shape
    assert relmax(q, g["q_" + tag]) < RTOL_Q
    assert abs(float(res["gw"][0]) - float(g["gw_" + tag])) < 1e-10 * float(g["gw_" + tag])


def test_printed_known_answers(catchment):
    """The reference's own known-answer test compares '%.6e' strings
    (tests/test_run_daily_to_hourly.py:135-143); examples/out/ExampleDaily/ExampleDaily.mod.flow
    holds the full 3653-day series at the same precision."""
    _torch()
    g = load_golden("runs_single")
    printed = load_golden("example_daily_printed")["mod_flow"]
    eng = make_engine(catchment, obs=False)
    q = eng.run(g["p_test"][None, :], discharge=True, scores=False)["discharge"][:, 0].cpu().numpy()
    assert ['%e' % v for v in q] == ['%e' % v for v in printed]
    # three of the 91 values spelled out in the reference test (first, and the last two)
    assert '%.6e' % q[-2] == '%.6e' % 6.7547748371e-01
    assert '%.6e' % q[-1] == '%.6e' % 8.3091723923e-01


# ---------------------------------------------------------------- members: LHS sample, range corners, .lhs rows
# default (block mode: one forcing row per day), hourly fast path, FORCE_GENERAL, NO_TMA, combinations
@pytest.mark.parametrize("flags", [0, 0x10000, 1, 2, 3, 0x10002, 0x10001])
def test_members_match_reference_and_oracle_scores(catchment, flags):
    _torch()
    from oracle import scores as oscores
    g = load_golden("runs_members")
    eng = make_engine(catchment, flags=flags)
    res = eng.run(g["params"], discharge=True, scores=True, gw=True)
    q = res["discharge"].cpu().numpy().T
    assert relmax(q, g["q"]) < RTOL_Q
    gw = res["gw"].cpu().numpy()
    assert relmax(gw, g["gw"]) < 1e-9
    sc = res["scores"].cpu().numpy()
    sc_ref = oscores.score_members(g["q"], g["gw"], catchment.obs, 0.12667)
    assert np.max(np.abs(sc[:, :7] - sc_ref[:, :7]) / np.maximum(1.0, np.abs(sc_ref[:, :7]))) < 1e-10
    assert np.array_equal(sc[:, 7], sc_ref[:, 7])
    # the reference's only artefact pinning the objective functions: the float32 .lhs file
    n0 = int(g["n_lhs"]) + int(g["n_corners"])
    file_scores = g["file_scores"].astype(np.float64)
    assert np.max(np.abs(sc[n0:, :5] - file_scores[:, :5])) < 5e-7
    assert np.max(np.abs(sc[n0:, 5:7] - file_scores[:, 5:7]) / np.abs(file_scores[:, 5:7])) < 5e-7
    assert np.array_equal(sc[n0:, 7], file_scores[:, 7])


@pytest.mark.parametrize("flags", [0, 0x10000])
def test_scores_only_equals_scores_with_discharge(catchment, flags):
    _torch()
    g = load_golden("runs_members")
    eng = make_engine(catchment, n_steps=24 * 400, flags=flags)
    a = eng.run(g["params"], discharge=True, scores=True)
    b = eng.run(g["params"], discharge=False, scores=True)
    assert np.array_equal(a["scores"].cpu().numpy(), b["scores"].cpu().numpy(), equal_nan=True)


def test_block_mode_is_independent_of_batch_composition(catchment):
    """A member's result must not depend on which other members share its warp (the warp-level
    shortcuts of the fast path only skip work that would not change any value)."""
    _torch()
    g = load_golden("runs_members")
    eng = make_engine(catchment, n_steps=24 * 600, obs=False)
    full = eng.run(g["params"], discharge=True, scores=False)["discharge"].cpu().numpy()
    rng = np.random.RandomState(0)
    perm = rng.permutation(len(g["params"]))
    shuffled = eng.run(g["params"][perm], discharge=True, scores=False)["discharge"].cpu().numpy()
    assert np.array_equal(shuffled, full[:, perm])
    alone = eng.run(g["params"][7:8], discharge=True, scores=False)["discharge"].cpu().numpy()
    assert np.array_equal(alone[:, 0], full[:, 7])


def test_best_member(catchment):
    _torch()
    g = load_golden("runs_members")
    eng = make_engine(catchment, n_steps=24 * 500)
    res = eng.run(np.tile(g["params"], (8, 1)), scores=True, best=('KGE', 1))
    sc = res["scores"].cpu().numpy()
    score, index = res["best"]
    assert int(index.item()) == int(np.argmax(sc[:, 1]))      # ties -> lowest index, like argmax
    assert float(score.item()) == sc[:, 1].max()
    res = eng.run(g["params"], scores=True, best=('RMSE', -1))
    sc = res["scores"].cpu().numpy()
    assert int(res["best"][1].item()) == int(np.argmin(sc[:, 6]))


# ---------------------------------------------------------------- daily time step: clamps + 95 % river cap fire
@pytest.mark.parametrize("tag,report,warm,gap", [
    ("g1_summary_w365", "summary", 365, 1),
    ("g13_summary_w0", "summary", 0, 13),
    ("g13_raw_w365", "raw", 365, 13),
    ("g13_summary_w26", "summary", 26, 13),
])
def test_daily_step_matches_reference(tag, report, warm, gap):
    _torch()
    from smartpy_b200.engine import BatchEngine
    g = load_golden("runs_daily")
    area = float(load_golden("catchment_processed")["area_m2"])
    eng = BatchEngine(g["rain"], g["peva"], area, 86400.0, gap, extra=EXTRA, warm_up_steps=warm, report=report)
    res = eng.run(g["params"], discharge=True, scores=False, gw=True)
    q = res["discharge"].cpu().numpy().T
    ref = g["q_" + tag]
    assert q.shape == ref.shape
    # a discharge that has decayed to ~1e-300 carries no relative information: floor the scale
    scale = np.maximum(np.abs(ref), 1e-12 * np.abs(ref).max(axis=1, keepdims=True))
    assert np.max(np.abs(q - ref) / scale) < RTOL_Q
    assert relmax(res["gw"].cpu().numpy(), g["gw_" + tag]) < 1e-9


def test_error_behaviour_matches_reference():
    _torch()
    from smartpy_b200.engine import BatchEngine
    g = load_golden("runs_daily")
    assert bool(g["summary_w365_g13_raises"])     # the reference raises ValueError (np.reshape)
    eng = BatchEngine(g["rain"], g["peva"], 1e8, 86400.0, 13, extra=EXTRA, warm_up_steps=365, report='summary')
    with pytest.raises(ValueError):
        eng.run(g["params"][:2], discharge=True, scores=False)
    eng = BatchEngine(g["rain"][:100], g["peva"][:100], 1e8, 86400.0, 1, warm_up_steps=101)
    with pytest.raises(Exception, match="warm-up"):           # structure.py:90-95
        eng.run(g["params"][:2], discharge=True, scores=False)
    with pytest.raises(Exception, match="unknown"):           # structure.py:70
        BatchEngine(g["rain"], g["peva"], 1e8, 86400.0, 1, report='mean')


# ---------------------------------------------------------------- smartcpp.allsteps contract
def test_allsteps_host_matches_reference():
    _torch()
    from smartpy_b200 import smartcpp_shim
    g = load_golden("runs_single")
    proc = load_golden("catchment_processed")
    rain = np.repeat(proc["rain_hourly_per_day"], 24)
    peva = np.repeat(proc["peva_hourly_per_day"], 24)
    q, gw, last = smartcpp_shim.allsteps(float(proc["area_m2"]), 3600.0, 4800, rain, peva, g["p_test"],
                                         g["allsteps_init"], 2, 1)
    assert relmax(q, g["allsteps_q_hourly"]) < RTOL_Q
    assert abs(gw - float(g["allsteps_gw"])) < 1e-10
    assert relmax(last[7:], g["allsteps_last"][7:]) < RTOL_Q
    one = load_golden("one_step")
    for i in (0, 3, 5, 17, 34):     # in range, out of range (i % 5 == 0), exact wet/dry ties (i % 17 == 0)
        out = np.asarray(smartcpp_shim.onestep(*one["cases"][i]))
        ref = one["outs"][i]
        scale = np.maximum(np.abs(ref), 1e-9 * np.abs(ref).max())
        assert np.max(np.abs(out - ref) / scale) < 1e-9


def test_one_step_known_answers_through_allsteps():
    """400 reference run_one_step cases (in and out of the parameter ranges) as 1-step runs."""
    _torch()
    from smartpy_b200.engine import BatchEngine
    g = load_golden("one_step")
    worst, worst_flux = 0.0, 0.0
    for dt in (3600.0, 86400.0, 900.0):
        sel = np.where(g["cases"][:, 1] == dt)[0]
        for i in sel[:40]:
            c, ref = g["cases"][i], g["outs"][i]
            eng = BatchEngine(c[2:3], c[3:4], c[0], dt, 1, report='raw')
            init = np.zeros((1, 19))
            init[0, 7:] = c[14:]
            res = eng.run(c[4:14][None, :], discharge=True, scores=False, gw=False, last_state=True,
                          initial_state=init)
            q = float(res["discharge"][0, 0])
            last = res["last_state"][0].cpu().numpy()
            # errors are judged against the magnitude of the quantities that entered the
            # arithmetic (a store that a leak with s' ~ 1 nearly empties keeps few relative digits)
            scale = max(abs(ref[6]), abs(c[25]) / dt, 1e-9)
            worst = max(worst, abs(q - ref[6]) / scale)
            # floor: 1e-3 mm of water over the catchment (a store the reference leaves at exactly 0
            # may hold ~1e-15 mm here: one ulp of a layer capacity, from the mm-state formulation)
            vs = np.maximum(np.maximum(np.abs(ref[7:]), np.abs(c[14:])), 1e-3 * c[0] / 1e3)
            worst = max(worst, float(np.max(np.abs(last[7:] - ref[7:]) / vs)))
            fs = np.maximum(np.abs(ref[:7]), 1e-9 * max(np.abs(ref[:7]).max(), 1e-9))
            worst_flux = max(worst_flux, float(np.max(np.abs(last[:7] - ref[:7]) / fs)))
    assert worst < 1e-11
    assert worst_flux < 1e-9


# ---------------------------------------------------------------- multi-catchment batches ([t][catchment] forcing)
@pytest.mark.parametrize("n_catch,mpc", [(6, 50), (3, 200), (10, 7)])
def test_multi_catchment_matches_oracle(catchment, oracle_lib, n_catch, mpc):
    _torch()
    from smartpy_b200.engine import BatchEngine, warm_up_length
    from oracle import scores as oscores
    rng = np.random.RandomState(123)
    n_steps = 24 * 120
    rain = np.stack([catchment.rain[k * 500:k * 500 + n_steps] * rng.uniform(0.5, 1.5) for k in range(n_catch)], 1)
    peva = np.stack([catchment.peva[k * 300:k * 300 + n_steps] * rng.uniform(0.8, 1.2) for k in range(n_catch)], 1)
    area = 10 ** rng.uniform(7, 9, n_catch)
    obs = np.stack([catchment.obs[k * 10:k * 10 + n_steps // 24] for k in range(n_catch)], 1)
    base = load_golden("runs_members")["params"]
    params = base[rng.randint(0, len(base), n_catch * mpc)]
    eng = BatchEngine(rain, peva, area, 3600.0, 24, obs=obs, extra=EXTRA,
                      warm_up_steps=warm_up_length(30, 3600.0), members_per_catchment=mpc)
    res = eng.run(params, discharge=True, scores=True, gw=True)
    q = res["discharge"].cpu().numpy().T
    sc = res["scores"].cpu().numpy()
    pick = rng.choice(n_catch * mpc, 24, replace=False)
    for m in pick:
        c = m // mpc
        q_ref, gw_ref = oracle_lib.run(area[c], 3600.0, rain[:, c].copy(), peva[:, c].copy(), params[m], EXTRA,
                                       n_steps, 24, warm_up=30)
        assert relmax(q[m], q_ref) < RTOL_Q
        s_ref = oscores.objectivefunction((q_ref, [gw_ref]), (obs[:, c], [None]))
        assert np.max(np.abs(sc[m, :7] - np.array(s_ref)) / np.maximum(1.0, np.abs(s_ref))) < 1e-9
        assert np.isnan(sc[m, 7])


# ---------------------------------------------------------------- ragged sizes: odd lengths, partial chunks, tails
@pytest.mark.parametrize("n_steps,gap,report,n_members", [
    (1, 1, 'raw', 1), (511, 1, 'summary', 3), (513, 3, 'raw', 129), (1025, 5, 'summary', 257), (2000, 7, 'raw', 5)])
def test_ragged_shapes_match_oracle(catchment, oracle_lib, n_steps, gap, report, n_members):
    _torch()
    from smartpy_b200.engine import BatchEngine
    base = load_golden("runs_members")["params"]
    params = np.resize(base, (n_members, 10))
    rain, peva = catchment.rain[100:100 + n_steps].copy(), catchment.peva[100:100 + n_steps].copy()
    warm = 0 if report == 'summary' and n_steps < gap else (n_steps // gap // 2) * gap
    for flags in (0, 2):
        eng = BatchEngine(rain, peva, catchment.area, 3600.0, gap, extra=EXTRA, warm_up_steps=warm, report=report,
                          flags=flags)
        res = eng.run(params, discharge=True, scores=False, gw=True)
        q = res["discharge"].cpu().numpy().T
        for m in (0, n_members - 1):
            q_ref, gw_ref = oracle_lib.run(catchment.area, 3600.0, rain, peva, params[m], EXTRA, n_steps, gap,
                                           report=report, warm_up=warm * 3600.0 / 86400.0)
            assert q[m].shape == q_ref.shape
            assert relmax(q[m], q_ref) < RTOL_Q


# ---------------------------------------------------------------- FP32 mode
def test_fp32_mode_scores_within_tolerance(catchment):
    _torch()
    from oracle import scores as oscores
    g = load_golden("runs_members")
    eng = make_engine(catchment, precision='f32')
    res = eng.run(g["params"], discharge=True, scores=True, gw=True)
    sc = res["scores"].cpu().numpy()
    sc_ref = oscores.score_members(g["q"], g["gw"], catchment.obs, 0.12667)
    assert np.max(np.abs(sc[:, 0] - sc_ref[:, 0])) < ATOL_F32     # NSE
    assert np.max(np.abs(sc[:, 1] - sc_ref[:, 1])) < ATOL_F32     # KGE
    assert res["discharge"].dtype == _torch().float32


# ---------------------------------------------------------------- host-pointer entry point
def test_batch_run_host_entry_point(catchment):
    _torch()
    from smartpy_b200 import _native
    lib = _native.load()
    g = load_golden("runs_members")
    n, days = 5, 200
    params = np.ascontiguousarray(g["params"][:n])
    rain = np.ascontiguousarray(catchment.rain[:days * 24])
    peva = np.ascontiguousarray(catchment.peva[:days * 24])
    obs = np.ascontiguousarray(catchment.obs[:days])
    area = np.array([catchment.area])
    q = np.zeros((days, n))
    sc = np.zeros((n, 8))
    gw = np.zeros(n)
    d = _native.BatchDesc()
    d.n_members, d.n_steps, d.n_warmup, d.n_catchments, d.members_per_catchment = n, days * 24, 24 * 30, 1, 1
    d.report_gap, d.report_type, d.dt_sec = 24, 1, 3600.0
    d.params, d.rain, d.peva = params.ctypes.data, rain.ctypes.data, peva.ctypes.data
    d.area_m2, d.obs = area.ctypes.data, obs.ctypes.data
    d.has_extra, d.aar, d.ro_ratio = 1, 1200.0, 0.45
    for k in range(5):
        d.ro_split[k] = EXTRA['r-o_split'][k]
    d.gw_constraint = float('nan')
    d.discharge, d.ld_discharge, d.scores, d.gw = q.ctypes.data, n, sc.ctypes.data, gw.ctypes.data
    _native.check(lib.smart_batch_run_host(ctypes.byref(d), 64, 0))
    import oracle
    q_ref, gw_ref = oracle.run_members(catchment.area, 3600.0, rain, peva, params, EXTRA, days * 24, 24, warm_up=30)
    assert relmax(q.T, q_ref) < RTOL_Q
    assert relmax(gw, gw_ref) < 1e-9


# ---------------------------------------------------------------- daily forcing in, equal split on the device
def test_device_disaggregation_is_bit_identical(catchment):
    """forcing_repeat=24 (smart_disaggregate) must reproduce the host split of
    timeframe.py:167-186 bit for bit, hence identical discharge."""
    torch = _torch()
    from smartpy_b200.engine import BatchEngine, warm_up_length
    proc = load_golden("catchment_processed")
    g = load_golden("runs_members")
    days = 400
    daily_rain = proc["rain_hourly_per_day"][:days] * 24.0     # not necessarily the raw totals: any daily series
    daily_peva = proc["peva_hourly_per_day"][:days] * 24.0
    kw = dict(extra=EXTRA, warm_up_steps=warm_up_length(30, 3600.0))
    a = BatchEngine(np.repeat(daily_rain / 24.0, 24), np.repeat(daily_peva / 24.0, 24), catchment.area, 3600.0, 24, **kw)
    b = BatchEngine(daily_rain, daily_peva, catchment.area, 3600.0, 24, forcing_repeat=24, **kw)
    assert torch.equal(a.rain, b.rain) and torch.equal(a.peva, b.peva)
    qa = a.run(g["params"], discharge=True, scores=False)["discharge"]
    qb = b.run(g["params"], discharge=True, scores=False)["discharge"]
    assert torch.equal(qa, qb)
    # [t][catchment] layout
    c = BatchEngine(np.stack([daily_rain, daily_rain * 0.5], 1), np.stack([daily_peva, daily_peva], 1),
                    [catchment.area, 2 * catchment.area], 3600.0, 24, forcing_repeat=24, members_per_catchment=20, **kw)
    assert torch.equal(c.rain[:, 0], a.rain) and torch.equal(c.rain[:, 1], torch.from_numpy(np.repeat(daily_rain * 0.5 / 24.0, 24)).to(c.rain.device))
    qc = c.run(g["params"], discharge=True, scores=False)["discharge"]
    assert torch.equal(qc[:, :20], qa[:, :20])


def test_general_kernel_on_out_of_range_members(catchment, oracle_lib):
    """Members outside the fast form's validity (S > 0.5, routing constants below dt, D > 1 ...)
    are routed to the branch-faithful kernel CTA by CTA, inside one batch with in-range members."""
    _torch()
    g = load_golden("runs_members")
    wild = np.array([
        [1.0, 0.5, 0.2, 0.3, 1.5, 60.0, 0.5, 2.0, 30.0, 0.2],
        [1.05, 0.9, 0.6, 0.5, 3.0, 20.0, 5.0, 10.0, 12.0, 3.0],
        [0.95, 0.1, 0.1, 0.9, 0.9, 100.0, 30.0, 30.0, 30.0, 30.0],
        [1.0, 0.5, 0.995, 1.0, 0.01, 50.0, 1.0, 48.0, 1200.0, 1.0],
    ])
    params = np.concatenate([g["params"][:70 % len(g["params"])], wild, np.resize(g["params"], (150, 10))])
    n_steps = 24 * 300
    eng = make_engine(catchment, n_steps=n_steps, obs=False, warm_up_days=20)
    q = eng.run(params, discharge=True, scores=False)["discharge"].cpu().numpy().T
    for m in list(range(28, 36)) + [0, 100, len(params) - 1]:
        q_ref, _ = oracle_lib.run(catchment.area, 3600.0, catchment.rain[:n_steps], catchment.peva[:n_steps], params[m],
                                  EXTRA, n_steps, 24, warm_up=20)
        scale = np.maximum(np.abs(q_ref), 1e-9 * np.abs(q_ref).max())
        assert np.max(np.abs(q[m] - q_ref) / scale) < RTOL_Q


def test_large_multi_catchment_batch_uses_wide_ctas(catchment, oracle_lib):
    """> 871k members: 128-member CTAs, three catchments per tile (needs > 48 KB of shared memory)."""
    _torch()
    from smartpy_b200.engine import BatchEngine
    n_catch, mpc, n_steps = 9000, 100, 48
    rng = np.random.RandomState(5)
    rain = np.ascontiguousarray(np.tile(catchment.rain[3000:3000 + n_steps, None], (1, n_catch)) * rng.uniform(0.5, 1.5, n_catch))
    peva = np.ascontiguousarray(np.tile(catchment.peva[3000:3000 + n_steps, None], (1, n_catch)))
    area = rng.uniform(1e7, 1e9, n_catch)
    base = load_golden("runs_members")["params"]
    params = np.resize(base, (n_catch * mpc, 10))
    eng = BatchEngine(rain, peva, area, 3600.0, 1, extra=EXTRA, report='raw', members_per_catchment=mpc)
    q = eng.run(params, discharge=True, scores=False)["discharge"]
    for m in (0, 123457, n_catch * mpc - 1):
        c = m // mpc
        q_ref, _ = oracle_lib.run(area[c], 3600.0, rain[:, c].copy(), peva[:, c].copy(), params[m], EXTRA, n_steps, 1,
                                  report='raw')
        assert relmax(q[:, m].cpu().numpy(), q_ref) < RTOL_Q


# ---------------------------------------------------------------- BASELINE config 3 shape: 30 years hourly
@pytest.mark.parametrize("flags", [0, 0x10000])     # block mode, per-step path
def test_thirty_year_synthetic_forcing_matches_oracle(oracle_lib, flags):
    """262,992 + 8,760 steps of the synthetic forcing bench.py uses for C3, a few
    members against the oracle, scores included: errors must not grow with the length of the run."""
    _torch()
    import bench
    from smartpy_b200.engine import BatchEngine
    from oracle import scores as oscores
    w = bench.make_workload("c3", 0, members=64)
    params = w["params"][[0, 9, 17, 31, 42, 63]]
    eng = BatchEngine(w["rain"], w["peva"], w["area"], w["dt"], w["gap"], obs=w["obs"], extra=w["extra"],
                      warm_up_steps=w["warm_steps"], gw_constraint=w["gwc"], flags=flags)
    res = eng.run(params, discharge=True, scores=True, gw=True)
    q = res["discharge"].cpu().numpy().T
    sc = res["scores"].cpu().numpy()
    q_ref, gw_ref = oracle_lib.run_members(w["area"], w["dt"], w["rain"], w["peva"], params, w["extra"], w["n_steps"],
                                           w["gap"], warm_up=365)
    assert relmax(q, q_ref) < RTOL_Q
    sc_ref = oscores.score_members(q_ref, gw_ref, w["obs"], w["gwc"])
    assert np.max(np.abs(sc[:, :7] - sc_ref[:, :7]) / np.maximum(1.0, np.abs(sc_ref[:, :7]))) < 1e-10
    assert np.array_equal(sc[:, 7], sc_ref[:, 7])
