#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (the reference lives at /root/reference, which does not
exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Everything written here is either (a) raw test data of the reference re-encoded as arrays
(tests/data/in/Catchment/*, examples/out/ExampleDaily/*), or (b) an output of the
reference's own functions (smartpy.SMART, smartpy.structure.run / run_all_steps /
run_one_step, montecarlo.lhs.LHS._get_params_from_lh) on stated inputs.  No reference
source is copied.  The fixtures pin oracle/ (tests/test_oracle_*.py) and, through the
oracle and directly, the CUDA path (tests/test_gpu_*.py).
"""
import os
import sys
import types
import csv
from datetime import datetime, timedelta

sys.dont_write_bytecode = True
REF = os.environ.get("SMART_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import smartpy  # noqa: E402
from smartpy import structure  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EPOCH = datetime(1970, 1, 1)

EXTRA = {'aar': 1200, 'r-o_ratio': 0.45, 'r-o_split': (0.10, 0.15, 0.15, 0.30, 0.30)}
NAMES = ['T', 'C', 'H', 'D', 'S', 'Z', 'SK', 'FK', 'GK', 'RK']


def secs(dt):
    return int((dt - EPOCH).total_seconds())


def read_raw_csv(path, val):
    """Raw rows -> (epoch seconds int64, float64 values with NaN for blank cells)."""
    ts, vs = [], []
    with open(path, 'r', encoding='utf8') as f:
        for row in csv.DictReader(f):
            ts.append(secs(datetime.strptime(row['DateTime'], "%Y-%m-%d %H:%M:%S")))
            vs.append(float(row[val]) if row[val] != '' else float('nan'))
    return np.asarray(ts, dtype=np.int64), np.asarray(vs, dtype=np.float64)


def ref_lhs_sample(n, seed):
    """The reference's own LHS routine (lhs.py:133-167), called unbound with spotpy stubbed."""
    stub = types.ModuleType('spotpy')
    sys.modules.setdefault('spotpy', stub)
    from smartpy.montecarlo.lhs import LHS
    fake = types.SimpleNamespace(
        model=types.SimpleNamespace(parameters=smartpy.parameters.Parameters()),
        param_names=NAMES)
    np.random.seed(seed)
    return LHS._get_params_from_lh(fake, n)


def main():
    in_dir = os.path.join(REF, 'tests', 'data', 'in', 'Catchment')
    out = {}

    # ---------------------------------------------------------------- raw inputs
    raw = {}
    for name in ('rain', 'peva', 'flow'):
        t, v = read_raw_csv(os.path.join(in_dir, 'Catchment.' + name), name)
        raw[name + '_t'] = t
        raw[name + '_v'] = v
    with open(os.path.join(in_dir, 'Catchment.parameters')) as f:
        par = {r['PAR_NAME']: float(r['PAR_VALUE']) for r in csv.DictReader(f)}
    raw['parameters'] = np.array([par[n] for n in NAMES])
    with open(os.path.join(in_dir, 'Catchment.sttngs')) as f:
        st = {r['ARGUMENT']: r['VALUE'] for r in csv.DictReader(f)}
    raw['sttngs_keys'] = np.array(list(st.keys()))
    raw['sttngs_vals'] = np.array(list(st.values()))
    np.savez_compressed(os.path.join(HERE, 'catchment_raw.npz'), **raw)

    # ---------------------------------------------------------------- reference SMART object
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, 'tests'))
    sm = smartpy.SMART(
        catchment='Catchment', catchment_area_m2=175.46 * 1E6,
        start=datetime.strptime('01/01/2007 09:00:00', '%d/%m/%Y %H:%M:%S'),
        end=datetime.strptime('31/12/2016 09:00:00', '%d/%m/%Y %H:%M:%S'),
        time_delta_simu=timedelta(hours=1), time_delta_save=timedelta(days=1),
        warm_up_days=365, in_format='csv', out_format='csv', root="data/",
        gauged_area_m2=175.97 * 1E6)
    os.chdir(cwd)
    sm.extra = dict(EXTRA)
    sm.parameters.set_parameters_with_file(os.path.join(in_dir, 'Catchment.parameters'))
    p_test = np.array([sm.parameters.values[n] for n in NAMES])

    assert sm.nd_rain.shape == (87672,)
    # hourly forcing is the daily value / 24 repeated 24x (timeframe.py:167-186): store per day
    rain_d = sm.nd_rain.reshape(-1, 24)
    peva_d = sm.nd_peva.reshape(-1, 24)
    assert (rain_d == rain_d[:, :1]).all() and (peva_d == peva_d[:, :1]).all()
    proc = dict(
        rain_hourly_per_day=rain_d[:, 0].copy(), peva_hourly_per_day=peva_d[:, 0].copy(),
        nd_flow=sm.nd_flow.copy(),
        simu_first=np.int64(secs(sm.timeseries[0])), simu_last=np.int64(secs(sm.timeseries[-1])),
        simu_len=np.int64(len(sm.timeseries)),
        save_first=np.int64(secs(sm.timeseries_report[0])), save_last=np.int64(secs(sm.timeseries_report[-1])),
        save_len=np.int64(len(sm.timeseries_report)),
        area_m2=np.float64(sm.area))
    np.savez_compressed(os.path.join(HERE, 'catchment_processed.npz'), **proc)

    # ---------------------------------------------------------------- single runs (C1) and variants
    def run(params, report='summary', warm_up=365, extra=EXTRA, sm=sm):
        d, gw = structure.run(sm.area, sm.delta_simu, sm.nd_rain, sm.nd_peva, np.asarray(params), extra,
                              sm.timeseries, sm.timeseries_report, report=report, warm_up=warm_up)
        return np.asarray(d, dtype=np.float64).copy(), float(gw)

    runs = {'p_test': p_test}
    for tag, kw in (('summary', {}), ('raw', dict(report='raw')), ('nowarm', dict(warm_up=0)),
                    ('noextra', dict(extra=None)), ('nowarm_noextra', dict(warm_up=0, extra=None)),
                    ('warm30_raw', dict(warm_up=30, report='raw'))):
        d, gw = run(p_test, **kw)
        runs['q_' + tag] = d
        runs['gw_' + tag] = np.float64(gw)
        print(tag, d[:2], gw)
    # simulate() through the facade must agree with structure.run
    d_f, gw_f = sm.simulate(sm.parameters.values)
    assert (d_f == runs['q_summary']).all() and gw_f == runs['gw_summary']

    # hourly Q_out + last state through run_all_steps (report_gap=1, raw) on the first 4800 steps
    init = np.zeros(19)
    init[7:12] = [1.0e5, 2.0e5, 3.0e5, 4.0e6, 5.0e6]
    init[12:18] = (p_test[5] / 12) / 1000 * sm.area
    init[18] = 5.0e4
    qh, gwh, last = structure.run_all_steps(sm.area, 3600.0, 4800, sm.nd_rain, sm.nd_peva, p_test, init, 2, 1)
    runs['allsteps_init'] = init
    runs['allsteps_q_hourly'] = np.asarray(qh).copy()
    runs['allsteps_gw'] = np.float64(gwh)
    runs['allsteps_last'] = np.asarray(last).copy()
    np.savez_compressed(os.path.join(HERE, 'runs_single.npz'), **runs)

    # ---------------------------------------------------------------- members: LHS(24, seed 42), corners, .lhs rows
    lhs24 = ref_lhs_sample(24, 42)
    rng = smartpy.parameters.Parameters().ranges
    lo = np.array([rng[n][0] for n in NAMES])
    hi = np.array([rng[n][1] for n in NAMES])
    corners = np.array([
        lo, hi,
        np.where(np.arange(10) % 2 == 0, lo, hi), np.where(np.arange(10) % 2 == 1, lo, hi),
        # S=0 is legal (lo corner); Z small + S large makes leaks matter; T at both ends
        [1.1, 0.0, 0.3, 0.0, 0.013, 15.0, 1.0, 48.0, 1200.0, 1.0],
        [0.9, 1.0, 0.0, 1.0, 0.013, 150.0, 240.0, 1440.0, 4800.0, 96.0],
    ])
    lhs_file = os.path.join(REF, 'examples', 'out', 'ExampleDaily', 'ExampleDaily.SMART.lhs')
    with open(lhs_file) as f:
        rows = list(csv.DictReader(f))
    score_cols = ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW']
    file_scores = np.array([[np.float32(r[c]) for c in score_cols] for r in rows], dtype=np.float32)
    file_params = np.array([[np.float32(r[c]) for c in NAMES] for r in rows], dtype=np.float32)

    members = np.concatenate([lhs24, corners, file_params.astype(np.float64)])
    q = np.zeros((members.shape[0], 3653))
    gw = np.zeros(members.shape[0])
    for i, p in enumerate(members):
        q[i], gw[i] = run(p)
        print('member', i, gw[i])
    np.savez_compressed(os.path.join(HERE, 'runs_members.npz'),
                        lhs24_seed42=lhs24, params=members, q=q, gw=gw,
                        n_lhs=np.int64(24), n_corners=np.int64(len(corners)), n_file=np.int64(len(rows)),
                        file_scores=file_scores, file_params=file_params,
                        file_score_names=np.array(score_cols))

    # ---------------------------------------------------------------- daily time step (clamps + 95% river cap fire)
    rain_daily = raw['rain_v']
    peva_daily = raw['peva_v']
    t0 = int(np.searchsorted(raw['rain_t'], secs(datetime(2007, 1, 1, 9))))
    rain_dd = rain_daily[t0:t0 + 3653].copy()
    peva_dd = peva_daily[t0:t0 + 3653].copy()
    ts = [None] * (3653 + 1)
    daily_members = np.concatenate([p_test[None, :], corners, lhs24[:6]])
    # a few deliberately out-of-range sets: S large (s' >= 1 leak predicates), tiny routing constants
    wild = np.array([
        [1.0, 0.5, 0.2, 0.3, 1.5, 60.0, 0.5, 2.0, 30.0, 0.2],
        [1.05, 0.9, 0.6, 0.5, 3.0, 20.0, 5.0, 10.0, 12.0, 3.0],
        [0.95, 0.1, 0.1, 0.9, 0.9, 100.0, 30.0, 30.0, 30.0, 30.0],
    ])
    daily_members = np.concatenate([daily_members, wild])
    dd = {'rain': rain_dd, 'peva': peva_dd, 'params': daily_members}

    def run_daily(p, report, warm_up, gap, extra=EXTRA):
        tsr = [None] * (3653 // gap + 1)
        d, g = structure.run(sm.area, timedelta(days=1), rain_dd, peva_dd, np.asarray(p), extra,
                             ts, tsr, report=report, warm_up=warm_up)
        return np.asarray(d, dtype=np.float64).copy(), float(g)

    for tag, (report, warm_up, gap) in (('g1_summary_w365', ('summary', 365, 1)),
                                        ('g13_summary_w0', ('summary', 0, 13)),
                                        ('g13_raw_w365', ('raw', 365, 13)),
                                        ('g13_summary_w26', ('summary', 26, 13))):
        qs, gs = [], []
        for p in daily_members:
            d, g = run_daily(p, report, warm_up, gap)
            qs.append(d)
            gs.append(g)
        dd['q_' + tag] = np.array(qs)
        dd['gw_' + tag] = np.array(gs)
        print(tag, np.array(gs)[:3])
    # reference behaviour when W % gap != 0 with 'summary': numpy reshape raises
    try:
        run_daily(p_test, 'summary', 365, 13)
        dd['summary_w365_g13_raises'] = np.bool_(False)
    except ValueError:
        dd['summary_w365_g13_raises'] = np.bool_(True)
    np.savez_compressed(os.path.join(HERE, 'runs_daily.npz'), **dd)

    # ---------------------------------------------------------------- one-step known answers (incl. out-of-range)
    r = np.random.RandomState(7)
    n = 400
    cases = np.zeros((n, 4 + 10 + 12))
    outs = np.zeros((n, 19))
    for i in range(n):
        area = 10 ** r.uniform(6, 9)
        dt = [3600.0, 86400.0, 900.0][i % 3]
        rain = 0.0 if r.rand() < 0.3 else r.gamma(0.7, 2.0)
        peva = 0.0 if r.rand() < 0.1 else r.uniform(0, 0.4)
        if i % 17 == 0:
            peva = rain * 1.0  # exact tie with T = 1 -> wet branch
        p = lo + r.rand(10) * (hi - lo)
        if i % 17 == 0:
            p[0] = 1.0
        if i % 5 == 0:  # out of range: s' >= 1 and fast reservoirs
            p[4] = r.uniform(0.5, 4.0)
            p[6:10] = r.uniform(0.05, 3.0, 4)
        z = p[5] / 6
        lv = r.uniform(0, 1, 6) * z
        lv[r.rand(6) < 0.2] = 0.0
        lv[r.rand(6) < 0.2] = z
        st_ = np.concatenate([r.uniform(0, 1e6, 5) * (r.rand(5) > 0.15), lv / 1e3 * area,
                              [r.uniform(0, 1e5) * (r.rand() > 0.15)]])
        cases[i] = np.concatenate([[area, dt, rain, peva], p, st_])
        outs[i] = structure.run_one_step(*cases[i])
    np.savez_compressed(os.path.join(HERE, 'one_step.npz'), cases=cases, outs=outs)

    # ---------------------------------------------------------------- the reference's printed goldens (7 digits)
    def read_flow_txt(path):
        with open(path) as f:
            rows_ = list(csv.DictReader(f))
        return np.array([float(r_['flow']) for r_ in rows_])

    ex = os.path.join(REF, 'examples', 'out', 'ExampleDaily')
    np.savez_compressed(os.path.join(HERE, 'example_daily_printed.npz'),
                        mod_flow=read_flow_txt(os.path.join(ex, 'ExampleDaily.mod.flow')),
                        obs_flow=read_flow_txt(os.path.join(ex, 'ExampleDaily.obs.flow')))
    print('done')


if __name__ == '__main__':
    main()
