"""CPU: pin the oracle (oracle/smart_oracle.c, oracle/scores.py) to the reference.

Every fixture under tests/golden was produced by running the unmodified reference
(tests/golden/make_golden.py).  Discharge must match BIT FOR BIT: the oracle restates
smartpy/structure.py operation for operation, in binary64, without FMA contraction."""
import numpy as np
import pytest

from conftest import load_golden, EXTRA


def test_one_step_bit_exact(oracle_lib):
    g = load_golden("one_step")
    for case, ref in zip(g["cases"], g["outs"]):
        assert np.array_equal(oracle_lib.onestep(case), ref)


@pytest.mark.parametrize("tag,kw", [
    ("summary", {}),
    ("raw", dict(report='raw')),
    ("nowarm", dict(warm_up=0)),
    ("noextra", dict(extra=None)),
    ("nowarm_noextra", dict(warm_up=0, extra=None)),
    ("warm30_raw", dict(warm_up=30, report='raw')),
])
def test_single_run_bit_exact(oracle_lib, catchment, tag, kw):
    g = load_golden("runs_single")
    args = dict(report='summary', warm_up=365, extra=EXTRA)
    args.update(kw)
    q, gw = oracle_lib.run(catchment.area, 3600.0, catchment.rain, catchment.peva, g["p_test"], args["extra"],
                           87672, 24, report=args["report"], warm_up=args["warm_up"])
    assert np.array_equal(q, g["q_" + tag])
    # np.sum's pairwise order over the fancy-indexed columns (structure.py:191) is not restated
    assert abs(gw - float(g["gw_" + tag])) <= 1e-12 * float(g["gw_" + tag])


def test_allsteps_bit_exact(oracle_lib, catchment):
    g = load_golden("runs_single")
    q, gw, last = oracle_lib.allsteps(catchment.area, 3600.0, 4800, catchment.rain, catchment.peva, g["p_test"],
                                      g["allsteps_init"], 2, 1)
    assert np.array_equal(q, g["allsteps_q_hourly"])
    assert np.array_equal(last, g["allsteps_last"])
    assert abs(gw - float(g["allsteps_gw"])) < 1e-13


def test_members_bit_exact(oracle_lib, catchment):
    g = load_golden("runs_members")
    sel = [0, 5, 23, 24, 25, 29, 30, 39]      # LHS rows, range corners, .lhs-file rows
    q, gw = oracle_lib.run_members(catchment.area, 3600.0, catchment.rain, catchment.peva, g["params"][sel], EXTRA,
                                   87672, 24, warm_up=365)
    assert np.array_equal(q, g["q"][sel])
    assert np.max(np.abs(gw - g["gw"][sel]) / g["gw"][sel]) < 1e-12


@pytest.mark.parametrize("tag,report,warm,gap", [
    ("g1_summary_w365", "summary", 365, 1),
    ("g13_summary_w0", "summary", 0, 13),
    ("g13_raw_w365", "raw", 365, 13),
    ("g13_summary_w26", "summary", 26, 13),
])
def test_daily_step_bit_exact(oracle_lib, catchment, tag, report, warm, gap):
    """Daily time step: the >= 0 clamps (structure.py:429-450) and the 95 % river cap (:492-498) fire."""
    g = load_golden("runs_daily")
    q, gw = oracle_lib.run_members(catchment.area, 86400.0, g["rain"], g["peva"], g["params"], EXTRA, 3653, gap,
                                   report=report, warm_up=warm)
    assert np.array_equal(q, g["q_" + tag])
    assert np.max(np.abs(gw - g["gw_" + tag]) / g["gw_" + tag]) < 1e-12


def test_error_behaviour(oracle_lib, catchment):
    g = load_golden("runs_daily")
    assert bool(g["summary_w365_g13_raises"])
    with pytest.raises(ValueError):
        oracle_lib.run(catchment.area, 86400.0, g["rain"], g["peva"], g["params"][0], EXTRA, 3653, 13,
                       report='summary', warm_up=365)
    with pytest.raises(Exception, match="warm-up"):
        oracle_lib.run(catchment.area, 86400.0, g["rain"][:100], g["peva"][:100], g["params"][0], EXTRA, 100, 1,
                       warm_up=101)
    with pytest.raises(Exception, match="unknown"):
        oracle_lib.run(catchment.area, 86400.0, g["rain"], g["peva"], g["params"][0], EXTRA, 3653, 1, report='mean')


def test_printed_known_answers():
    """Reference's own goldens at 7 significant digits: tests/test_run_daily_to_hourly.py:31-121
    (values below quoted from it) and examples/out/ExampleDaily/ExampleDaily.{mod,obs}.flow."""
    single = load_golden("runs_single")
    printed = load_golden("example_daily_printed")
    proc = load_golden("catchment_processed")
    q = single["q_summary"]
    assert ['%e' % v for v in q] == ['%e' % v for v in printed["mod_flow"]]
    assert '%.6e' % q[-2] == '%.6e' % 6.7547748371e-01      # 2016-12-30 09:00:00
    assert '%.6e' % q[-1] == '%.6e' % 8.3091723923e-01      # 2016-12-31 09:00:00
    obs = proc["nd_flow"]
    assert np.array_equal(np.isnan(obs), np.isnan(printed["obs_flow"]))
    ok = ~np.isnan(obs)
    assert ['%e' % v for v in obs[ok]] == ['%e' % v for v in printed["obs_flow"][ok]]
    assert int(np.isnan(obs).sum()) == 425


def test_scores_match_reference_lhs_file(catchment):
    """The only artefact of the reference that pins the objective functions (spotpy is not
    vendored): examples/out/ExampleDaily/ExampleDaily.SMART.lhs, float32."""
    from oracle import scores
    g = load_golden("runs_members")
    n0 = int(g["n_lhs"]) + int(g["n_corners"])
    mine = scores.score_members(g["q"][n0:], g["gw"][n0:], catchment.obs, 0.12667)
    ref = g["file_scores"].astype(np.float64)
    assert list(g["file_score_names"]) == scores.SCORE_NAMES
    assert np.max(np.abs(mine[:, :5] - ref[:, :5])) < 2e-7
    assert np.max(np.abs(mine[:, 5:7] - ref[:, 5:7]) / np.abs(ref[:, 5:7])) < 2e-7
    assert np.array_equal(mine[:, 7], ref[:, 7])


def test_scores_anchor_values(catchment):
    """Full-precision anchors measured on the reference during the survey."""
    from oracle import scores
    single = load_golden("runs_single")
    s = scores.objectivefunction((single["q_summary"], [float(single["gw_summary"])]), (catchment.obs, [None]))
    anchors = [0.39044538284151, 0.25208537324827, 0.93416082775515, 0.49907600668112, 0.44853228088603,
               -55.146771911397, 4.3106307613864]
    assert np.allclose(s, anchors, rtol=1e-12, atol=0)
    assert abs(float(single["gw_summary"]) - 0.0529869529029892) < 1e-15
    assert scores.groundwater_constraint([0.12667], [0.2]) == 1.0
    assert scores.groundwater_constraint([0.12667], [0.23]) == 0.0
