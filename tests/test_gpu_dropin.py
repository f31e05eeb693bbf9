"""GPU: the drop-in surface end to end -- files in, `SMART.simulate` / `montecarlo.LHS.run`
out -- mirroring the reference's two tests (tests/test_run_daily_to_hourly.py,
tests/test_run_mc_lhs.py), with values checked against reference-generated goldens."""
import gzip
import os
from datetime import datetime, timedelta

import numpy as np
import pytest

from conftest import load_golden, EXTRA
from test_host_logic import catchment_dir  # noqa: F401  (fixture: CSV files from the raw fixture)

pytestmark = pytest.mark.gpu


def _model(root):
    import smartpy_b200 as smartpy
    sm = smartpy.SMART(
        catchment='Catchment', catchment_area_m2=175.46 * 1E6,
        start=datetime.strptime('01/01/2007 09:00:00', '%d/%m/%Y %H:%M:%S'),
        end=datetime.strptime('31/12/2016 09:00:00', '%d/%m/%Y %H:%M:%S'),
        time_delta_simu=timedelta(hours=1), time_delta_save=timedelta(days=1),
        warm_up_days=365, in_format='csv', out_format='csv', root=root,
        gauged_area_m2=175.97 * 1E6)
    sm.extra = dict(EXTRA)
    return sm


def test_run_daily_to_hourly(catchment_dir):  # noqa: F811
    """The reference's TestRunDaily2Hourly, same calls, on the GPU."""
    sm = _model(catchment_dir)
    sm.parameters.set_parameters_with_file(''.join([sm.in_f, sm.catchment, '.parameters']))
    out = sm.simulate(sm.parameters.values)
    g = load_golden("runs_single")
    assert out[0] is sm.nd_discharge and out[1] == sm.gw_contribution
    assert np.max(np.abs(sm.nd_discharge - g["q_summary"]) / g["q_summary"]) < 1e-10
    assert abs(sm.gw_contribution - float(g["gw_summary"])) < 1e-10
    expected = {   # three of the 91 pinned values of tests/test_run_daily_to_hourly.py:31-121
        datetime.strptime('2016-12-30 09:00:00', '%Y-%m-%d %H:%M:%S'): 6.7547748371e-01,
        datetime.strptime('2016-12-31 09:00:00', '%Y-%m-%d %H:%M:%S'): 8.3091723923e-01,
    }
    for dt, val in expected.items():
        assert '%.6e' % sm.nd_discharge[sm.timeseries_report[1:].index(dt)] == '%.6e' % val
    raw = sm.simulate(sm.parameters.values, report='raw')
    assert np.max(np.abs(raw[0] - g["q_raw"]) / g["q_raw"]) < 1e-10
    with pytest.raises(Exception, match="unknown"):
        sm.simulate(sm.parameters.values, report='mean')
    sm.simulate(sm.parameters.values)
    sm.write_output_files(which='both')
    printed = load_golden("example_daily_printed")["mod_flow"]
    with open(os.path.join(sm.out_f, 'Catchment.mod.flow')) as f:
        lines = f.read().splitlines()
    assert lines[0] == 'DateTime,flow'
    assert [ln.split(',')[1] for ln in lines[1:]] == ['%e' % v for v in printed]


def test_run_mc_lhs(catchment_dir):  # noqa: F811
    """The reference's TestRunMonteCarloLHS (csv in/out, save_sim, compression), plus values."""
    from smartpy_b200 import montecarlo
    from oracle import scores as oscores
    import oracle
    np.random.seed(42)
    setup = montecarlo.LHS(catchment='Catchment', root_f=catchment_dir, in_format='csv', out_format='csv',
                           sample_size=5, parallel='seq', save_sim=True)
    setup.model.extra = dict(EXTRA)
    assert np.array_equal(setup.lhs_params.shape, (5, 10))
    setup.run(compression=True)
    db = setup.db_file + '.gz'
    assert os.path.exists(db) and not os.path.exists(setup.db_file)
    with gzip.open(db, 'rt') as f:
        lines = f.read().splitlines()
    header = lines[0].split(',')
    assert header[:8] == ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW']     # gw_constraint is set
    assert header[8:18] == ['T', 'C', 'H', 'D', 'S', 'Z', 'SK', 'FK', 'GK', 'RK']
    assert header[18] == '2007-01-01 09:00:00' and len(header) == 18 + 3653 and len(lines) == 6
    c = setup.model
    q_ref, gw_ref = oracle.run_members(c.area, 3600.0, c.nd_rain, c.nd_peva, setup.lhs_params, EXTRA, 87672, 24,
                                       warm_up=365)
    sc_ref = oscores.score_members(q_ref, gw_ref, c.nd_flow, 0.12667)
    for k, line in enumerate(lines[1:]):
        vals = line.split(',')
        want = ['%.6e' % np.float32(v) for v in list(sc_ref[k]) + list(setup.lhs_params[k]) + list(q_ref[k])]
        assert vals[:18] == want[:18]
        # simulated series: float32 text; allow the last printed digit to differ
        got = np.array(vals[18:], dtype=np.float64)
        assert np.max(np.abs(got - q_ref[k]) / q_ref[k]) < 2e-6
    # the spotpy-protocol methods still answer
    sim = setup.simulation(setup.lhs_params[0])
    assert np.max(np.abs(sim[0] - q_ref[0]) / q_ref[0]) < 1e-10
    of = setup.objectivefunction(sim, setup.evaluation())
    assert np.allclose(of, sc_ref[0], rtol=1e-9, atol=0)
    # conditioning of the scores that are still on the device
    rows, picked = setup.select_best('NSE', 2)
    assert rows.tolist() == np.argsort(sc_ref[:, 0], kind='stable')[-2:].tolist()
    assert np.array_equal(picked, setup.lhs_params[rows.cpu().numpy()])
    rows, _ = setup.select_behavioural({'KGE': ('max', (float(np.median(sc_ref[:, 1])),))})
    assert rows.tolist() == np.nonzero(sc_ref[:, 1] <= np.median(sc_ref[:, 1]))[0].tolist()
    # conditioning on top of that sample
    np.random.seed(42)
    montecarlo.LHS('Catchment', catchment_dir, 'csv', 'csv', sample_size=5).run()
    best = montecarlo.Best('Catchment', catchment_dir, 'csv', 'csv', target='NSE', nb_best=2)
    order = np.argsort(sc_ref[:, 0].astype(np.float32), kind='stable')[-2:]
    as_stored = np.array([[np.float32('%.6e' % np.float32(v)) for v in row] for row in setup.lhs_params[order]],
                         dtype=np.float32)      # the database keeps 7 significant digits of float32
    assert np.array_equal(best.best_params, as_stored)
    best.run()
    assert best.results['scores'].shape == (2, 8)
    glue = montecarlo.GLUE('Catchment', catchment_dir, 'csv', 'csv', conditioning={'KGE': ('min', (-10.0,))})
    assert glue.behavioural_params.shape == (5, 10)
    total = montecarlo.Total('Catchment', catchment_dir, 'csv', 'csv')
    total.model.extra = dict(EXTRA)          # as in the reference, `extra` is set on each new setup
    total.run()
    assert np.allclose(total.results['scores'][:, 0].cpu().numpy(), sc_ref[:, 0], atol=1e-4)


def test_best_and_glue_from_a_run_in_memory(catchment_dir):  # noqa: F811
    """Best.from_run / GLUE.from_run / Total.from_run: the score table a sampling run left on the
    device is conditioned there (smart_best_rows / smart_condition_rows) and the selected rows go
    back through the same engine -- no database round trip, binary64 parameters."""
    from smartpy_b200 import montecarlo
    np.random.seed(7)
    lhs = montecarlo.LHS('Catchment', catchment_dir, 'csv', 'csv', sample_size=60)
    lhs.model.extra = dict(EXTRA)
    with pytest.raises(Exception, match="call run"):
        montecarlo.Best.from_run(lhs, 'NSE', 3)
    lhs.run()
    sc = lhs.results['scores'].cpu().numpy()
    # Best: constraint + top-k, ascending with the best last, exact float64 rows of the sample
    kge_floor = float(np.median(sc[:, 1]))
    best = montecarlo.Best.from_run(lhs, target='NSE', nb_best=5, constraining={'KGE': ('min', (kge_floor,))})
    kept = np.nonzero(sc[:, 1] >= kge_floor)[0]
    expect = kept[np.argsort(sc[kept, 0], kind='stable')][-5:]
    assert best.best_rows.tolist() == expect.tolist()
    assert np.array_equal(best.best_params, lhs.sample_params[expect])
    assert best.db_file.endswith('Catchment.SMART.5best') and best.target_fn_index == [0]
    assert best.constraints_indices == [1] and best.constraints_types == ['min']
    best.model.extra = dict(EXTRA)
    best.run()
    again = best.results['scores'].cpu().numpy()
    assert np.array_equal(again, sc[expect], equal_nan=True)        # same period, same parameters: same bits
    with open(best.db_file) as f:
        assert len(f.read().splitlines()) == 6
    with pytest.raises(Exception, match="restrained sample size"):
        montecarlo.Best.from_run(lhs, 'NSE', 40, constraining={'KGE': ('min', (kge_floor,))})
    with pytest.raises(Exception, match="not recognised"):
        montecarlo.Best.from_run(lhs, 'nse', 2)
    # GLUE: behavioural rows in sample order
    glue = montecarlo.GLUE.from_run(lhs, {'NSE': ('min', (float(np.median(sc[:, 0])),)), 'PBias': ('inside', (-90.0, 90.0))})
    mask = (sc[:, 0] >= np.median(sc[:, 0])) & (sc[:, 5] >= -90.0) & (sc[:, 5] <= 90.0)
    assert glue.behavioural_rows.tolist() == np.nonzero(mask)[0].tolist()
    assert np.array_equal(glue.behavioural_params, lhs.sample_params[mask])
    # Total: the whole sample again
    total = montecarlo.Total.from_run(lhs)
    assert np.array_equal(total.sample_params, lhs.sample_params) and total.db_file.endswith('.SMART.total')
