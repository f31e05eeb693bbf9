"""CPU: the host side of the drop-in API (no compute calls): time frames, rescaling, file
readers, parameters, settings, LHS sampler, conditioning, sharding, the C-ABI library."""
import ctypes
import os
import re
import subprocess
from datetime import datetime, timedelta

import numpy as np
import pytest

from conftest import load_golden, ROOT


# ---------------------------------------------------------------- files materialised from the raw fixture
@pytest.fixture(scope="module")
def catchment_dir(tmp_path_factory):
    from smartpy_b200.timeframe import from_seconds
    raw = load_golden("catchment_raw")
    root = tmp_path_factory.mktemp("smart")
    d = root / "in" / "Catchment"
    d.mkdir(parents=True)
    for name in ("rain", "peva", "flow"):
        with open(d / ("Catchment." + name), "w") as f:
            f.write("DateTime,%s\n" % name)
            for t, v in zip(raw[name + "_t"], raw[name + "_v"]):
                f.write("%s,%s\n" % (from_seconds(t).strftime("%Y-%m-%d %H:%M:%S"),
                                     "" if np.isnan(v) else repr(float(v))))
    with open(d / "Catchment.parameters", "w") as f:
        f.write("PAR_NAME,PAR_VALUE\n")
        for n, v in zip(['T', 'C', 'H', 'D', 'S', 'Z', 'SK', 'FK', 'GK', 'RK'], raw["parameters"]):
            f.write("%s,%r\n" % (n, float(v)))
    with open(d / "Catchment.sttngs", "w") as f:
        f.write("ARGUMENT,VALUE\n")
        for k, v in zip(raw["sttngs_keys"], raw["sttngs_vals"]):
            f.write("%s,%s\n" % (k, v))
    return str(root)


@pytest.fixture(scope="module")
def model(catchment_dir):
    import smartpy_b200
    return smartpy_b200.SMART('Catchment', 175.46e6, datetime(2007, 1, 1, 9), datetime(2016, 12, 31, 9),
                              timedelta(hours=1), timedelta(days=1), 365, 'csv', 'csv', catchment_dir,
                              gauged_area_m2=175.97e6)


def test_smart_ingest_matches_reference_bit_for_bit(model):
    """SMART.__init__ (smart.py:125-143): forcing disaggregation daily -> hourly and the
    observation rescaling must reproduce the reference's arrays exactly."""
    from smartpy_b200.timeframe import to_seconds
    proc = load_golden("catchment_processed")
    assert np.array_equal(model.nd_rain, np.repeat(proc["rain_hourly_per_day"], 24))
    assert np.array_equal(model.nd_peva, np.repeat(proc["peva_hourly_per_day"], 24))
    assert np.array_equal(model.nd_flow, proc["nd_flow"], equal_nan=True)
    assert len(model.timeseries) == int(proc["simu_len"]) and len(model.timeseries_report) == int(proc["save_len"])
    assert to_seconds(model.timeseries[0]) == int(proc["simu_first"])
    assert to_seconds(model.timeseries[-1]) == int(proc["simu_last"])
    assert to_seconds(model.timeseries_report[0]) == int(proc["save_first"])
    assert to_seconds(model.timeseries_report[-1]) == int(proc["save_last"])
    assert model.nd_rain[0] == model.nd_rain[23]            # a daily total split in 24 equal parts
    assert len(model.flow) == 3653 and list(model.flow)[0] == datetime(2007, 1, 1, 9)
    assert model.rain[model.timeseries[1]] == model.nd_rain[0]


def test_smart_api_surface(model, catchment_dir):
    import smartpy_b200
    assert smartpy_b200.__version__
    assert smartpy_b200.objfunctions.groundwater_constraint([0.5], [0.45]) == 1.0
    assert smartpy_b200.objfunctions.groundwater_constraint([0.5], [0.61]) == 0.0
    model.parameters.set_parameters_with_file(os.path.join(catchment_dir, 'in', 'Catchment', 'Catchment.parameters'))
    assert model.parameters.names == ['T', 'C', 'H', 'D', 'S', 'Z', 'SK', 'FK', 'GK', 'RK']
    assert np.array_equal(model.parameters.as_row(), load_golden("catchment_raw")["parameters"])
    assert model.parameters.ranges['GK'] == (1200.0, 4800.0)
    with pytest.raises(Exception, match="simulate"):
        model.get_simulation_array()
    assert model.get_evaluation_array() is model.nd_flow
    with pytest.raises(Exception, match="modelled flow"):
        model.write_output_files('modelled')
    model.write_output_files('observed')
    with open(os.path.join(model.out_f, 'Catchment.obs.flow')) as f:
        lines = f.read().splitlines()
    assert lines[0] == 'DateTime,flow' and len(lines) == 3654
    assert lines[1] == '2007-01-01 09:00:00,%e' % model.nd_flow[0]


def test_no_cpu_fallback(model):
    """Without a GPU the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    model.parameters.set_parameters_with_dict(dict(zip(model.parameters.names, load_golden("catchment_raw")["parameters"])))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.simulate(model.parameters.values)


def test_parameters_errors(tmp_path):
    from smartpy_b200.parameters import Parameters
    p = Parameters()
    with pytest.raises(Exception, match="no parameters file"):
        p.set_parameters_with_file(str(tmp_path / "missing.parameters"))
    f = tmp_path / "x.parameters"
    f.write_text("PAR_NAME,PAR_VALUE\nT,1.0\nC,abc\n")
    with pytest.raises(Exception, match="incorrect parameter value"):
        p.set_parameters_with_file(str(f))
    f.write_text("PAR_NAME,PAR_VALUE\nT,1.0\n")
    with pytest.raises(Exception, match="parameter C is not available"):
        p.set_parameters_with_file(str(f))
    f.write_text("NAME,VALUE\nT,1.0\n")
    with pytest.raises(Exception, match="PAR_NAME"):
        p.set_parameters_with_file(str(f))
    with pytest.raises(Exception, match="not available in the dictionary"):
        p.set_parameters_with_dict({'T': 1.0})


def test_settings_file(catchment_dir, tmp_path):
    from smartpy_b200.inout import get_dict_simulation_settings
    s = get_dict_simulation_settings(os.path.join(catchment_dir, 'in', 'Catchment', 'Catchment.sttngs'))
    assert s == (175.46 * 1e6, 175.97 * 1e6, datetime(2007, 1, 1, 9), datetime(2016, 12, 31, 9),
                 timedelta(hours=1), timedelta(days=1), 365, 0.12667)
    f = tmp_path / "a.sttngs"
    f.write_text("ARGUMENT,VALUE\ncatchment_area_km2,10\nstart_datetime,01/01/2007 09:00:00\n"
                 "end_datetime,02/01/2007 09:00:00\nsimu_timedelta_min,60\nreport_timedelta_min,1440\nwarm_up_days,0\n")
    s = get_dict_simulation_settings(str(f))
    assert s[1] == s[0] == 10e6 and s[-1] is None
    f.write_text("ARGUMENT,VALUE\ncatchment_area_km2,ten\n")
    with pytest.raises(Exception, match="CATCHMENT AREA could not be converted"):
        get_dict_simulation_settings(str(f))
    f.write_text("ARGUMENT,VALUE\ncatchment_area_km2,10\n")
    with pytest.raises(Exception, match="START is missing"):
        get_dict_simulation_settings(str(f))


# ---------------------------------------------------------------- time frames
def test_timeframe_series():
    from smartpy_b200.timeframe import TimeFrame
    tf = TimeFrame(datetime(2007, 1, 1, 9), datetime(2007, 1, 3, 9), timedelta(hours=1), timedelta(days=1))
    assert tf.get_series_save() == [datetime(2006, 12, 31, 9), datetime(2007, 1, 1, 9), datetime(2007, 1, 2, 9),
                                    datetime(2007, 1, 3, 9)]
    simu = tf.get_series_simu()
    assert simu[0] == datetime(2006, 12, 31, 9) and simu[1] == datetime(2006, 12, 31, 10)
    assert simu[-1] == datetime(2007, 1, 3, 9) and len(simu) == 73
    assert tf.get_simu_length() == 72 and tf.get_report_gap() == 24
    assert tf.get_gap_simu() == timedelta(hours=1) and tf.get_gap_report() == timedelta(days=1)
    with pytest.raises(Exception, match="Save Start is greater"):
        TimeFrame(datetime(2007, 1, 2), datetime(2007, 1, 1), timedelta(hours=1), timedelta(days=1))
    with pytest.raises(Exception, match="not compatible"):
        TimeFrame(datetime(2007, 1, 1), datetime(2007, 1, 2, 5), timedelta(hours=1), timedelta(days=1))
    with pytest.raises(Exception, match="multiple of Simulation Gap"):
        TimeFrame(datetime(2007, 1, 1), datetime(2007, 1, 2), timedelta(hours=5), timedelta(days=1))


def test_cumulative_rescaling_dict_api_agrees_with_array_core():
    from smartpy_b200 import timeframe as tfm
    day, hour = timedelta(days=1), timedelta(hours=1)
    start = datetime(2000, 1, 1, 9)
    vals = {start + k * day: float(k + 1) * 0.37 for k in range(6)}
    hi = tfm.increase_time_resolution_of_regular_cumulative_data(vals, start, start + 5 * day, day, hour)
    assert len(hi) == 6 * 24
    assert hi[start] == hi[start - 23 * hour] == 0.37 / 24
    lo = tfm.decrease_time_resolution_of_regular_cumulative_data(hi, start + day, start + 5 * day, day, hour)
    total = 0.0
    for _ in range(24):
        total += 2 * 0.37 / 24
    assert lo[start + day] == total                         # sequential re-aggregation, same order
    # 6-hourly simulation from daily data, shifted start: resolution = gcd(shift, gcd(24 h, 6 h))
    res = tfm.get_required_resolution(start, start - 18 * hour, day, 6 * hour)
    assert res == 6 * hour
    out = tfm.rescale_time_resolution_of_regular_cumulative_data(
        vals, start, start + 5 * day, day, res, start - 18 * hour, start + 5 * day, 6 * hour)
    assert out[start - 18 * hour] == 0.37 / 4 and out[start + 5 * day] == 6 * 0.37 / 4
    with pytest.raises(Exception, match="not multiples"):
        tfm.increase_time_resolution_of_regular_cumulative_data(vals, start, start + day, day, timedelta(hours=5))
    with pytest.raises(KeyError):
        tfm.rescale_time_resolution_of_regular_cumulative_data(
            vals, start, start + 5 * day, day, hour, start - 2 * day, start, hour)


def test_mean_data_rescaling_gaps_and_missing():
    from smartpy_b200 import timeframe as tfm
    from collections import OrderedDict
    day, hour = timedelta(days=1), timedelta(hours=1)
    t0 = datetime(2007, 1, 1, 0)
    obs = OrderedDict()
    for k, v in ((0, 1.0), (1, 2.0), (2, 3.0), (5, 6.0), (6, 7.0)):    # days 3 and 4 missing
        obs[t0 + k * day] = v
    out = tfm.rescale_time_resolution_of_irregular_mean_data(obs, t0 + 9 * hour, t0 + 9 * hour + 6 * day, day, hour)
    vals = list(out.values())
    # report stamp D 09:00 averages D-1 10:00 .. D 09:00: 15 h of the value stamped D 00:00 and
    # 9 h of the value stamped D+1 00:00 (replicated backwards); any uncovered hour -> NaN
    assert vals[0] == (15 * 1.0 + 9 * 2.0) / 24 and vals[1] == (15 * 2.0 + 9 * 3.0) / 24
    assert vals[5] == (15 * 6.0 + 9 * 7.0) / 24
    assert all(np.isnan(vals[k]) for k in (2, 3, 4, 6))
    ref = tfm.decrease_time_resolution_of_irregular_mean_data(
        tfm.increase_time_resolution_of_irregular_mean_data(obs, day, hour),
        t0 + 9 * hour, t0 + 9 * hour + 6 * day, hour, day)
    assert np.array_equal(np.array(vals), np.array(list(ref.values())), equal_nan=True)


def test_interval_checks():
    from smartpy_b200.timeframe import check_interval_in_list
    day = timedelta(days=1)
    t0 = datetime(2000, 1, 1)
    assert check_interval_in_list([t0, t0 + day, t0 + 2 * day], "f") == (t0, t0 + 2 * day, day)
    with pytest.raises(Exception, match="Inconsistent Interval"):
        check_interval_in_list([t0, t0 + day, t0 + 3 * day], "f")


# ---------------------------------------------------------------- Monte Carlo host logic
def test_lhs_sampler_reproduces_reference_stream():
    from smartpy_b200.montecarlo.lhs import latin_hypercube
    from smartpy_b200.parameters import Parameters
    p = Parameters()
    bounds = [p.ranges[n] for n in p.names]
    np.random.seed(42)
    sample = latin_hypercube(24, bounds)
    assert np.array_equal(sample, load_golden("runs_members")["lhs24_seed42"])
    # stratification: every column has exactly one value in each of the N strata
    big = latin_hypercube(1000, bounds, rng=np.random.RandomState(1))
    for k, (lo, hi) in enumerate(bounds):
        strata = np.floor((big[:, k] - lo) / (hi - lo) * 1000).astype(int)
        assert sorted(strata.tolist()) == list(range(1000))


def test_conditioning_rules():
    from smartpy_b200.montecarlo.glue import GLUE
    from smartpy_b200.montecarlo.best import Best
    params = np.arange(50, dtype=np.float32).reshape(5, 10)
    fns = np.array([[0.1, 5.0], [0.5, 3.0], [0.7, 1.0], [0.9, 9.0], [0.3, 2.0]], dtype=np.float32)
    keep = GLUE._get_behavioural_sets(params, fns, [(0.3,), (1.5, 6.0)], ['min', 'inside'])
    assert np.array_equal(keep, params[[1, 4]])
    keep = GLUE._get_behavioural_sets(params, fns[:, :1], [(0.5,)], ['max'])
    assert np.array_equal(keep, params[[0, 1, 4]])
    with pytest.raises(Exception, match="inconsistent"):
        GLUE._get_behavioural_sets(params, fns[:, :1], [(2.0, 1.0)], ['inside'])
    with pytest.raises(Exception, match="not in the database"):
        GLUE._get_behavioural_sets(params, fns[:, :1], [(2.0,)], ['above'])
    best = Best._get_best_sets(params, fns[:, 1:], [(8.0,)], ['max'], fns[:, :1], 2)
    assert np.array_equal(best, params[[1, 2]])              # ascending by target, best last
    with pytest.raises(Exception, match="restrained sample size"):
        Best._get_best_sets(params, fns[:, 1:], [(2.5,)], ['max'], fns[:, :1], 3)


def test_shard_bounds_cover_rows_once():
    from smartpy_b200.distributed import shard_bounds
    for n, world in ((10, 3), (8, 8), (5, 8), (10 ** 7, 8), (1, 2)):
        spans = [shard_bounds(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------- the C ABI library
def test_library_exports_every_declared_symbol():
    from smartpy_b200 import _native, _build
    _build.build()
    lib = _native.load()
    header = open(os.path.join(ROOT, "include", "smart_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|size_t|const char \*)\s*\*?\s*(smart_\w+)\(", header, flags=re.M))
    assert declared == set(_native.SYMBOLS)
    io_header = open(os.path.join(ROOT, "include", "smart_b200_io.h")).read()
    declared_io = set(re.findall(r"^(?:int|int64_t|size_t|const char \*)\s*\*?\s*(smart_\w+)\(", io_header, flags=re.M))
    assert declared_io == set(_native.IO_SYMBOLS)
    for name in declared | declared_io:
        assert hasattr(lib, name), name
    assert lib.smart_version() == 200
    # argument validation happens before any CUDA call, so it is testable without a GPU
    d = _native.BatchDesc()
    assert lib.smart_batch_run_f64(ctypes.byref(d), None) == _native.ERR_BAD_ARG
    assert "n_members" in _native.last_error()


def test_member_order_sizes_are_host_functions():
    """smart_member_order_len / _workspace_bytes need no device: the fast group padded to a CTA
    boundary (128) plus the other group always fit; the workspace holds the sort's padded pairs."""
    from smartpy_b200 import _native, _build
    _build.build()
    lib = _native.load()
    for n in (1, 127, 128, 129, 100000, 1250000):
        slots = lib.smart_member_order_len(n)
        assert slots % 128 == 0 and slots >= n + 127 and slots <= n + 255
        padded = 1 << max(n - 1, 0).bit_length()
        assert lib.smart_member_order_workspace_bytes(n) >= padded * 12
    # argument validation before any CUDA call
    assert lib.smart_member_order(None, 5, 3600.0, None, None, None) == _native.ERR_BAD_ARG
    assert lib.smart_fold_blocks(None, 48, 1, 24, None, None, None) == _native.ERR_BAD_ARG
    assert lib.smart_launch_count() >= 0


def test_launch_workspace_size_is_a_host_function():
    """smart_batch_workspace_bytes needs no device.  Single-catchment launches get room for the relay
    (2 tickets + progress counters + 26 doubles per thread of the launch, include/smart_b200.h) behind
    the per-CTA slots of the best-member search; multi-catchment launches, launches that return the
    state, and batches beyond the relay's range only get the best-member slots."""
    import ctypes
    from smartpy_b200 import _native, _build
    _build.build()
    lib = _native.load()

    def desc(n, catchments=1, best=0, slots=0, last_state=False):
        d = _native.BatchDesc()
        d.n_members, d.n_steps, d.n_catchments = n, 240, catchments
        d.members_per_catchment = n // catchments
        d.best_sign = best
        d.member_order_len = slots
        d.member_order = 8 if slots else None
        d.last_state = 8 if last_state else None
        return d

    def pad(x):
        return (x + 127) // 128 * 128

    n = 100000
    groups = -(-n // 64)                         # 64-member CTAs for a batch of this size
    relay = 128 + pad(4 * groups) + 8 * groups * 26 * 64
    assert lib.smart_batch_workspace_bytes(ctypes.byref(desc(n))) == relay
    assert lib.smart_batch_workspace_bytes(ctypes.byref(desc(n, best=1))) == pad(16 * groups) + relay
    slots = lib.smart_member_order_len(n)        # the launch has one thread per slot of member_order
    groups = slots // 64
    assert lib.smart_batch_workspace_bytes(ctypes.byref(desc(n, slots=slots))) == \
        128 + pad(4 * groups) + 8 * groups * 26 * 64
    assert lib.smart_batch_workspace_bytes(ctypes.byref(desc(n, last_state=True))) == 0
    assert lib.smart_batch_workspace_bytes(ctypes.byref(desc(1000000, catchments=10000))) == 0
    blocks = 1000000 // 32                       # one-warp CTAs for whole warps per catchment
    assert lib.smart_batch_workspace_bytes(ctypes.byref(desc(1000000, catchments=10000, best=-1))) == 16 * blocks
    assert lib.smart_batch_workspace_bytes(ctypes.byref(desc(4000000))) == 0      # beyond the relay's range
    assert lib.smart_batch_workspace_bytes(None) == 0
    assert _native.flag_relay_segs(9) == 9 << 8 and _native.flag_relay_segs(300) == (300 & 0xff) << 8


def test_binary64_unit_is_built_without_implicit_contraction():
    """The determinism of binary64 results across kernel instantiations rests on -fmad=false for
    smart_kernels.cu (tests/test_gpu_fullsize.py); the FP32-state unit keeps the default."""
    from smartpy_b200 import _build
    units = dict(_build.UNITS)
    assert units["smart_kernels.cu"] == ["-fmad=false"]
    assert units["smart_kernels_f32.cu"] == []
    src = open(os.path.join(ROOT, "smartpy_b200", "csrc", "smart_kernels_f32.cu")).read()
    assert "#define SMART_TU_F32" in src and '#include "smart_kernels.cu"' in src


def test_selection_and_sampler_entry_points_validate_before_touching_the_device():
    """smart_condition_rows / smart_best_rows / smart_lhs_rows reject bad arguments with
    SMART_ERR_BAD_ARG and a message, before any CUDA call (so this runs without a GPU)."""
    from smartpy_b200 import _native
    lib = _native.load()
    cond = (_native.Condition * 1)(_native.Condition(0, 9, 0.0, 0.0))              # unknown kind
    assert lib.smart_condition_rows(8, 10, 8, cond, 1, 8, 8, 8, None) == _native.ERR_BAD_ARG
    assert "kind" in _native.last_error()
    cond = (_native.Condition * 1)(_native.Condition(8, 1, 0.0, 0.0))              # column outside the table
    assert lib.smart_condition_rows(8, 10, 8, cond, 1, 8, 8, 8, None) == _native.ERR_BAD_ARG
    assert "column" in _native.last_error()
    assert lib.smart_condition_rows(8, 10, 8, cond, _native.MAX_CONDITIONS + 1, 8, 8, 8, None) == _native.ERR_BAD_ARG
    assert lib.smart_best_rows(8, 10, 8, 0, None, 0, 11, 8, 8, 8, None) == _native.ERR_BAD_ARG   # k > n_rows
    assert "k must be" in _native.last_error()
    assert lib.smart_best_rows(8, 10, 8, 8, None, 0, 1, 8, 8, 8, None) == _native.ERR_BAD_ARG    # target column
    bounds = np.array([[0.0, 1.0]] * 3)
    assert lib.smart_lhs_rows(1, 10, 5, 6, 3, bounds.ctypes.data, 8, None) == _native.ERR_BAD_ARG   # rows past n_total
    assert lib.smart_lhs_rows(1, 10, 0, 10, 17, bounds.ctypes.data, 8, None) == _native.ERR_BAD_ARG  # too many columns
    assert "smart_lhs_rows" in _native.last_error()
    assert lib.smart_condition_workspace_bytes(1000, 10) >= 8 * 1000
    # the ctypes mirror of smart_condition: 24 bytes, doubles at 8 and 16
    assert ctypes.sizeof(_native.Condition) == 24
    assert (_native.Condition.lo.offset, _native.Condition.hi.offset) == (8, 16)


def test_descriptor_layout_matches_header(tmp_path):
    """ctypes mirror vs the C struct: size and every field offset, checked with gcc."""
    from smartpy_b200 import _native
    fields = [f[0] for f in _native.BatchDesc._fields_]
    src = tmp_path / "layout.c"
    body = "\n".join('printf("%s %%zu\\n", offsetof(smart_batch_desc, %s));' % (f, f) for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "smart_b200.h"\nint main(void){\n'
                   'printf("sizeof %%zu\\n", sizeof(smart_batch_desc));\n%s\nreturn 0;}\n' % body)
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["sizeof"]) == ctypes.sizeof(_native.BatchDesc)
    for f in fields:
        assert int(out[f]) == getattr(_native.BatchDesc, f).offset, f


def test_device_conditioning_argument_checks():
    """conditioning.py: the reference's error messages are raised before any device work, and a
    CPU tensor is refused (there is no CPU path)."""
    import torch
    from smartpy_b200.montecarlo import conditioning
    names = ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW']
    t = torch.zeros((20, 8), dtype=torch.float64)
    with pytest.raises(Exception, match="not recognised"):
        conditioning.best_rows(t, names, 'XYZ', 1)
    with pytest.raises(Exception, match="higher than the sample size"):
        conditioning.best_rows(t, names, 'NSE', 21)
    with pytest.raises(Exception, match="one and only one"):
        conditioning.behavioural_rows(t, names, {'NSE': ('min', (0.1, 0.2))})
    with pytest.raises(Exception, match="inconsistent"):
        conditioning.behavioural_rows(t, names, {'NSE': ('inside', (0.3, 0.2))})
    with pytest.raises(Exception, match="not in the database"):
        conditioning.behavioural_rows(t, names, {'NSE': ('about', (0.3,))})
    with pytest.raises(RuntimeError, match="no CPU path"):
        conditioning.behavioural_rows(t, names, {'NSE': ('min', (0.1,))})


def test_sample_database_round_trip(tmp_path):
    """database.SampleDatabase writes the reference's format ('%.6e' of the float32 value, header =
    objective functions + parameters + report stamps, optional gzip) in bulk; read_sample_database
    finds the columns by name and returns float32 -- checked against a row-by-row restatement of
    the reference's writer (montecarlo.py:226-231) and reader (:247-262)."""
    import gzip
    from csv import DictReader
    from datetime import datetime, timedelta
    from smartpy_b200.montecarlo.database import SampleDatabase, read_sample_database, database_path
    rng = np.random.RandomState(0)
    fns, names = ['NSE', 'KGE', 'GW'], ['T', 'C', 'H']
    scores, params, sims = rng.randn(7, 3) * 10, rng.rand(7, 3) * 100, rng.rand(7, 4) * 1e-3
    stamps = [datetime(2007, 1, 1, 9) + k * timedelta(days=1) for k in range(4)]
    path = database_path(str(tmp_path) + os.sep, 'Catchment', 'lhs', 'csv')
    assert path.endswith('Catchment.SMART.lhs') and database_path('x/', 'C', 'glue', 'netcdf') == 'x/C.SMART.glue.nc'
    db = SampleDatabase(path, 'csv', fns + names, stamps).open()
    db.write_rows(scores[:3], params[:3], sims[:3])
    db.write_rows(scores[3:], params[3:], sims[3:])
    with pytest.raises(ValueError):
        db.write_rows(scores, params)                                 # the series are part of every row
    db.close()
    lines = open(path).read().splitlines()
    assert lines[0] == 'NSE,KGE,GW,T,C,H,2007-01-01 09:00:00,2007-01-02 09:00:00,2007-01-03 09:00:00,2007-01-04 09:00:00'
    for r in range(7):                                                # the reference's per-row formatting
        values = np.concatenate([scores[r], params[r], sims[r]])
        assert lines[1 + r] == ','.join('%.6e' % np.float32(v) for v in values)
    p, f = read_sample_database(path, 'csv', ['H', 'T'], ['GW', 'NSE'])
    rows = list(DictReader(open(path)))
    assert np.array_equal(p, np.array([[row['H'], row['T']] for row in rows], dtype=np.float32))
    assert np.array_equal(f, np.array([[row['GW'], row['NSE']] for row in rows], dtype=np.float32))
    # gzip: <file>.gz replaces the file
    db = SampleDatabase(path, 'csv', fns + names).open()
    db.write_rows(scores, params)
    db.close(compression=True)
    assert not os.path.exists(path) and gzip.open(path + '.gz', 'rt').readline().strip() == 'NSE,KGE,GW,T,C,H'
    p2, _ = read_sample_database(path, 'csv', names, fns, gzipped=True)
    assert np.array_equal(p2, np.array([['%.6e' % np.float32(v) for v in row] for row in params], dtype=np.float32))
    with pytest.raises(Exception, match="netCDF4"):
        SampleDatabase(path, 'netcdf', fns + names)
    with pytest.raises(Exception, match="netCDF4"):
        read_sample_database(path, 'netcdf', names, fns)


def test_no_netcdf_code_is_left_unexecuted():
    """'netcdf' takes the reference's "package netCDF4 missing" exit everywhere (inout.py, database.py);
    nothing imports netCDF4."""
    import smartpy_b200.inout as inout
    from smartpy_b200.montecarlo import database, montecarlo
    for module in (inout, database, montecarlo):
        assert 'Dataset' not in vars(module)
    with pytest.raises(Exception, match="netCDF4"):
        inout.read_rain_file('nowhere.nc', 'netcdf')
    with pytest.raises(Exception, match="netCDF4"):
        inout.write_flow_file_from_nds([], [], 'nowhere', 'netcdf')


def test_bench_only_uses_captures_of_the_running_sources():
    """bench.py folds ncu evidence into its line only when the capture was taken from the sources
    that run (round 1 injected an instruction count of older sources)."""
    import bench
    sha = bench.csrc_hash()
    assert len(sha) == 16 and sha == bench.csrc_hash()
    table = {"c2:100000:f64:0": {"csrc_sha": sha, "fp64_inst_per_step": 28.0},
             "c3:1250000:f64:0": {"csrc_sha": "0" * 16, "fp64_inst_per_step": 30.0}}
    cap, ok = bench.capture_for(table, "c2:100000:f64:0", sha)
    assert ok and cap["fp64_inst_per_step"] == 28.0
    cap, ok = bench.capture_for(table, "c3:1250000:f64:0", sha)
    assert not ok and cap["csrc_sha"] == "0" * 16          # stale: reported as such, not used
    cap, ok = bench.capture_for(table, "c5:100000:f32:0", sha)
    assert not ok and cap == {}


def test_sample_database_text_is_numpys_byte_for_byte():
    """include/smart_b200_io.h: the library formats and parses the rows of the sample database
    (montecarlo.py:211-231 writer, :233-262 reader) on all host cores.  The bytes must be those of
    numpy.savetxt(fmt='%.6e') -- Python's correctly rounded '%.6e' of the float32 value -- and the
    values read back those of text -> binary64 -> float32, on ordinary values, the whole exponent
    range, exact ties of the seventh digit, neighbours of the powers of ten, random bit patterns and
    the non-finite values (the GW column is NaN when the constraint is off)."""
    import io
    from smartpy_b200 import _native, _build
    from smartpy_b200.montecarlo import database as db
    _build.build()
    lib = _native.load()
    assert all(hasattr(lib, name) for name in _native.IO_SYMBOLS)
    rng = np.random.RandomState(3)

    def check(table):
        table = np.ascontiguousarray(table, dtype=np.float32)
        text = db.format_rows(table)
        ref = io.BytesIO()
        np.savetxt(ref, table, fmt='%.6e', delimiter=',')
        assert text == ref.getvalue()
        cols = list(range(table.shape[1]))[::-1] + [0]
        back = db.parse_rows(text, table.shape[1], cols)
        want = np.loadtxt(io.BytesIO(text), delimiter=',', dtype=np.float64, ndmin=2).astype(np.float32)[:, cols]
        assert np.array_equal(back, want, equal_nan=True)

    check(np.abs(rng.randn(5000, 18)) * 50)
    check(rng.randn(5000, 7) * 10.0 ** rng.randint(-44, 38, (5000, 7)))
    k = rng.randint(1000000, 8388607, (5000, 3))
    check(k + 0.5)                       # seventh digit exactly half way: ties to even
    check((k + 0.5) / 1024.0)
    check((k + 0.5) * 4096.0)
    check(rng.randint(0, 2 ** 32, (20000, 5)).astype(np.uint32).view(np.float32))
    powers = np.array([[10.0 ** e for e in range(-45, 39)]], dtype=np.float32)
    check(powers)
    check(np.nextafter(powers, np.float32(0)))
    check(np.nextafter(powers, np.float32(np.inf)))
    check(np.array([[0.0, -0.0, np.nan, np.inf, -np.inf, 1e-45, 3.4028235e38, 9.9999995e6, 99999996.0, 0.1]]))
    # text this library did not write (more digits, no exponent, spaces) goes through strtod
    loose = b"1.5,-2.25e3, 7\n0.1234567891234,1e-400,1e400\n"
    got = db.parse_rows(loose, 3, [0, 1, 2])
    assert np.array_equal(got, np.array([[1.5, -2250.0, 7.0], [0.1234567891234, 0.0, np.inf]], dtype=np.float32))
    with pytest.raises(ValueError):
        db.parse_rows(b"1.0,2.0\n3.0\n", 2, [0, 1])
    # a block boundary inside write_rows / an empty table
    assert db.format_rows(np.zeros((0, 4), dtype=np.float32)) == b""
    assert lib.smart_csv_format_f32(None, 1, 1, 1, None, 0, 0) == _native.ERR_BAD_ARG
