"""ctypes binding of oracle/smart_oracle.c (TEST INFRASTRUCTURE -- see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsmart_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile smart_oracle.c with gcc (no FMA contraction).  Idempotent."""
    src = os.path.join(_HERE, "smart_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
             "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        lib = ctypes.CDLL(_SO)
        lib.smart_oracle_onestep.argtypes = [_dp, _dp]
        lib.smart_oracle_onestep.restype = None
        lib.smart_oracle_allsteps.argtypes = [
            ctypes.c_double, ctypes.c_double, ctypes.c_long, _dp, _dp, _dp, _dp,
            ctypes.c_int, ctypes.c_long, _dp, _dp, _dp, _dp]
        lib.smart_oracle_allsteps.restype = ctypes.c_int
        lib.smart_oracle_run.argtypes = [
            ctypes.c_double, ctypes.c_double, ctypes.c_long, _dp, _dp, _dp,
            ctypes.c_int, ctypes.c_double, ctypes.c_double, _dp,
            ctypes.c_double, ctypes.c_int, ctypes.c_long, _dp, _dp]
        lib.smart_oracle_run.restype = ctypes.c_int
        _lib = lib
    return _lib


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def onestep(values26):
    """run_one_step (structure.py:200-264): 26 inputs -> 19 outputs."""
    a, pa = _c(values26)
    assert a.shape == (26,)
    out = np.zeros(19)
    _load().smart_oracle_onestep(pa, out.ctypes.data_as(_dp))
    return out


def allsteps(area_m2, delta_sec, length_simu, nd_rain, nd_peva, nd_parameters, nd_initial,
             report_type, report_gap, return_storage=False):
    """run_all_steps (structure.py:149-197): -> (discharge, gw, last[19]) [+ storage]."""
    rain, prain = _c(nd_rain)
    peva, ppeva = _c(nd_peva)
    par, ppar = _c(nd_parameters)
    ini, pini = _c(nd_initial)
    assert len(rain) >= length_simu and len(peva) >= length_simu
    n_rep = length_simu // report_gap if report_type == 1 else -(-length_simu // report_gap)
    q = np.zeros(max(n_rep, 0))
    gw = ctypes.c_double(0.0)
    last = np.zeros(19)
    storage = np.zeros((length_simu + 1, 19)) if return_storage else None
    rc = _load().smart_oracle_allsteps(
        float(area_m2), float(delta_sec), int(length_simu), prain, ppeva, ppar, pini,
        int(report_type), int(report_gap), q.ctypes.data_as(_dp), ctypes.byref(gw),
        last.ctypes.data_as(_dp),
        storage.ctypes.data_as(_dp) if return_storage else None)
    if rc == -2:
        raise ValueError("cannot reshape: length_simu is not a multiple of report_gap")
    if rc:
        raise MemoryError("oracle allocation failed")
    if return_storage:
        return q, gw.value, last, storage
    return q, gw.value, last


def run(area_m2, delta_sec, nd_rain, nd_peva, nd_parameters, extra, simu_length, report_gap,
        report='summary', warm_up=0):
    """run (structure.py:30-146) with array-length arguments instead of datetime lists:
    simu_length = len(timeseries) - 1, report_gap as computed at structure.py:75."""
    if report == 'summary':
        report_type = 1
    elif report == 'raw':
        report_type = 2
    else:
        raise Exception('Reporting type \'{}\' unknown.'.format(report))
    rain, prain = _c(nd_rain)
    peva, ppeva = _c(nd_peva)
    par, ppar = _c(nd_parameters)
    has_extra = 1 if extra else 0
    split = np.asarray(extra['r-o_split'], dtype=np.float64) if extra else np.zeros(5)
    n_rep = simu_length // report_gap if report_type == 1 else -(-simu_length // report_gap)
    q = np.zeros(n_rep)
    gw = ctypes.c_double(0.0)
    rc = _load().smart_oracle_run(
        float(area_m2), float(delta_sec), int(simu_length), prain, ppeva, ppar,
        has_extra, float(extra['aar']) if extra else 0.0, float(extra['r-o_ratio']) if extra else 0.0,
        split.ctypes.data_as(_dp), float(warm_up), report_type, int(report_gap),
        q.ctypes.data_as(_dp), ctypes.byref(gw))
    if rc == -3:
        raise Exception("The warm-up duration (i.e. {} days) cannot exceed the length of the "
                        "simulation period".format(warm_up))
    if rc == -2:
        raise ValueError("cannot reshape: run length is not a multiple of report_gap")
    if rc:
        raise MemoryError("oracle allocation failed")
    return q, gw.value


def run_members(area_m2, delta_sec, nd_rain, nd_peva, params, extra, simu_length, report_gap,
                report='summary', warm_up=0):
    """Loop `run` over the rows of params[N,10] -> (discharge[N, n_report], gw[N])."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    qs, gws = [], []
    for p in params:
        q, gw = run(area_m2, delta_sec, nd_rain, nd_peva, p, extra, simu_length, report_gap,
                    report=report, warm_up=warm_up)
        qs.append(q)
        gws.append(gw)
    return np.array(qs), np.array(gws)
