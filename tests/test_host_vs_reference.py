"""CPU, build container only (skipped where /root/reference does not exist): the host-side
re-implementations (TimeFrame, the rescaling helpers, Parameters, the LHS sampler, the conditioning
rules) against the reference's own functions on randomised inputs -- results must be identical."""
import os
import sys
import types
from collections import OrderedDict
from datetime import datetime, timedelta

import numpy as np
import pytest

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "smartpy")),
                                reason="the reference is only mounted in the build container")


@pytest.fixture(scope="module")
def ref():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REFERENCE)
    sys.modules.setdefault("spotpy", types.ModuleType("spotpy"))     # only its import is needed
    import smartpy
    from smartpy import timeframe, parameters
    from smartpy.montecarlo import glue, best, lhs
    yield types.SimpleNamespace(smartpy=smartpy, timeframe=timeframe, parameters=parameters, glue=glue, best=best,
                                lhs=lhs)
    sys.path.remove(REFERENCE)
    for k in [k for k in sys.modules if k == "smartpy" or k.startswith("smartpy.")]:
        del sys.modules[k]


@pytest.mark.parametrize("start,end,simu,save", [
    (datetime(2007, 1, 1, 9), datetime(2007, 3, 1, 9), timedelta(hours=1), timedelta(days=1)),
    (datetime(2010, 6, 1, 0), datetime(2010, 6, 20, 0), timedelta(hours=6), timedelta(days=1)),
    (datetime(2012, 2, 27, 12), datetime(2012, 3, 5, 12), timedelta(days=1), timedelta(days=1)),
    (datetime(2015, 1, 1, 0), datetime(2015, 1, 3, 0), timedelta(minutes=15), timedelta(hours=3)),
])
def test_timeframe_matches_reference(ref, start, end, simu, save):
    from smartpy_b200.timeframe import TimeFrame
    mine, theirs = TimeFrame(start, end, simu, save), ref.timeframe.TimeFrame(start, end, simu, save)
    assert mine.get_series_simu() == theirs.get_series_simu()
    assert mine.get_series_save() == theirs.get_series_save()
    assert (mine.simu_start, mine.simu_end) == (theirs.simu_start, theirs.simu_end)


def test_timeframe_errors_match_reference(ref):
    from smartpy_b200.timeframe import TimeFrame
    for args in ((datetime(2007, 1, 2), datetime(2007, 1, 1), timedelta(hours=1), timedelta(days=1)),
                 (datetime(2007, 1, 1), datetime(2007, 1, 2, 5), timedelta(hours=1), timedelta(days=1)),
                 (datetime(2007, 1, 1), datetime(2007, 1, 3), timedelta(hours=5), timedelta(days=1))):
        with pytest.raises(Exception) as a:
            TimeFrame(*args)
        with pytest.raises(Exception) as b:
            ref.timeframe.TimeFrame(*args)
        assert str(a.value) == str(b.value)


@pytest.mark.parametrize("data_delta,simu_delta,shift_h", [
    (timedelta(days=1), timedelta(hours=1), 0), (timedelta(days=1), timedelta(hours=6), 6),
    (timedelta(hours=6), timedelta(hours=3), 0), (timedelta(days=1), timedelta(days=1), 0),
    (timedelta(hours=1), timedelta(hours=3), 0),
])
def test_cumulative_rescaling_matches_reference(ref, data_delta, simu_delta, shift_h):
    from smartpy_b200 import timeframe as mine
    rng = np.random.RandomState(1)
    start_data = datetime(2000, 1, 1, 9)
    n = 80
    data = {start_data + k * data_delta: np.float64(rng.gamma(0.7, 3.0)) for k in range(n)}
    end_data = start_data + (n - 1) * data_delta
    start_simu = start_data + 5 * data_delta - timedelta(hours=shift_h)
    start_simu = start_simu - (data_delta - simu_delta) if simu_delta < data_delta else start_simu
    n_simu = int((end_data - start_simu) / simu_delta) - 3
    end_simu = start_simu + n_simu * simu_delta
    res = mine.get_required_resolution(start_data, start_simu, data_delta, simu_delta)
    assert res == ref.timeframe.get_required_resolution(start_data, start_simu, data_delta, simu_delta)
    args = (data, start_data, end_data, data_delta, res, start_simu, end_simu, simu_delta)
    a = mine.rescale_time_resolution_of_regular_cumulative_data(*args)
    b = ref.timeframe.rescale_time_resolution_of_regular_cumulative_data(*args)
    assert list(a.keys()) == list(b.keys())
    assert list(a.values()) == list(b.values())          # bit-identical floats


def test_mean_rescaling_matches_reference(ref):
    from smartpy_b200 import timeframe as mine
    rng = np.random.RandomState(2)
    for stamp_hour in (0, 1, 9):
        data = OrderedDict()
        day = datetime(2006, 12, 20, stamp_hour)
        for k in range(120):
            if rng.rand() > 0.2:                       # gaps of one or more days
                data[day + timedelta(days=k)] = np.float64(rng.gamma(2.0, 3.0))
        start, end = datetime(2007, 1, 1, 9), datetime(2007, 4, 1, 9)
        a = mine.rescale_time_resolution_of_irregular_mean_data(data, start, end, timedelta(days=1), timedelta(hours=1))
        b = ref.timeframe.rescale_time_resolution_of_irregular_mean_data(data, start, end, timedelta(days=1),
                                                                         timedelta(hours=1))
        assert list(a.keys()) == list(b.keys())
        assert np.array_equal(np.array(list(a.values())), np.array(list(b.values())), equal_nan=True)
        hi_a = mine.increase_time_resolution_of_irregular_mean_data(data, timedelta(days=1), timedelta(hours=1))
        hi_b = ref.timeframe.increase_time_resolution_of_irregular_mean_data(data, timedelta(days=1), timedelta(hours=1))
        assert hi_a == hi_b


def test_parameters_and_lhs_match_reference(ref):
    from smartpy_b200.parameters import Parameters
    from smartpy_b200.montecarlo.lhs import latin_hypercube
    mine, theirs = Parameters(), ref.parameters.Parameters()
    assert mine.names == theirs.names and mine.ranges == theirs.ranges
    fake = types.SimpleNamespace(model=types.SimpleNamespace(parameters=theirs), param_names=theirs.names)
    for n, seed in ((7, 0), (1000, 5)):
        np.random.seed(seed)
        expected = ref.lhs.LHS._get_params_from_lh(fake, n)
        np.random.seed(seed)
        got = latin_hypercube(n, [mine.ranges[k] for k in mine.names])
        assert np.array_equal(got, expected)


def test_conditioning_matches_reference(ref):
    from smartpy_b200.montecarlo.glue import GLUE
    from smartpy_b200.montecarlo.best import Best
    rng = np.random.RandomState(4)
    params = rng.rand(300, 10).astype(np.float32)
    fns = rng.randn(300, 3).astype(np.float32)
    for vals, kinds in (([(0.1,), (-0.5, 0.5), (0.0,)], ['min', 'inside', 'max']),
                        ([(0.3,), (-1.0, 1.0), (0.2, 0.9)], ['max', 'outside', 'inside'])):
        assert np.array_equal(GLUE._get_behavioural_sets(params, fns, vals, kinds),
                              ref.glue.GLUE._get_behavioural_sets(params, fns, vals, kinds))
    a = Best._get_best_sets(params, fns[:, 1:], [(0.5,), (-1.0,)], ['max', 'min'], fns[:, :1], 9)
    b = ref.best.Best._get_best_sets(params, fns[:, 1:], [(0.5,), (-1.0,)], ['max', 'min'], fns[:, :1], 9)
    assert np.array_equal(a, b)


def test_sample_database_reader_matches_the_reference_algorithm_on_its_example_file():
    """read_sample_database against the reference's reader restated here (csv.DictReader row by row,
    strings to np.float32 arrays, montecarlo.py:247-262) on the reference's own example database
    (written by an older version: six extra columns, found by name)."""
    from csv import DictReader
    from smartpy_b200.montecarlo.database import read_sample_database
    from smartpy_b200.parameters import Parameters
    path = os.path.join(REFERENCE, "examples", "out", "ExampleDaily", "ExampleDaily.SMART.lhs")
    names = Parameters().names
    fns = ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW']
    params, obj_fns = read_sample_database(path, 'csv', names, fns)
    with open(path, 'r', encoding='utf8') as handle:
        rows = list(DictReader(handle))
    ref_params = np.array([[row[p] for p in names] for row in rows], dtype=np.float32)
    ref_fns = np.array([[row[f] for f in fns] for row in rows], dtype=np.float32)
    assert params.dtype == np.float32 and obj_fns.dtype == np.float32
    assert np.array_equal(params, ref_params) and np.array_equal(obj_fns, ref_fns)
    with pytest.raises(KeyError):
        read_sample_database(path, 'csv', names, ['NoSuchScore'])
