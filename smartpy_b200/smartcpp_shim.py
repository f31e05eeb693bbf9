"""A ``smartcpp``-compatible module backed by the B200 kernel.

The reference's only native plug point is ``import smartcpp`` (smartpy/structure.py:22-27):
when that module has ``allsteps`` the reference calls it instead of ``run_all_steps``
(structure.py:56-62, call sites :118-121 and :143-146); ``onestep`` replaces ``run_one_step``
(structure.py:171-174, :182-187).  Making the UNMODIFIED reference run on the GPU is one line
before importing it::

    import sys, smartpy_b200.smartcpp_shim as shim; sys.modules['smartcpp'] = shim
    import smartpy

Each call is a batch of one member through ``smart_allsteps_host`` (host pointers in, host
results out, synchronous) -- the same contract as the C++ module it stands in for.  For
throughput use ``smartpy_b200.SMART.simulate_batch`` / ``montecarlo`` instead.
"""
import ctypes

import numpy as np

from . import _native

__version__ = '0.2.0'   # >= 0.2.0 means "has allsteps" to the reference (structure.py:57)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def allsteps(area_m2, delta_sec, length_simu, nd_rain, nd_peva, nd_parameters, nd_initial,
             report_type, report_gap, device=0):
    """-> (discharge ndarray, groundwater_component float, last 19-vector ndarray)."""
    lib = _native.load()
    rain = np.ascontiguousarray(nd_rain, dtype=np.float64)
    peva = np.ascontiguousarray(nd_peva, dtype=np.float64)
    par = np.ascontiguousarray(nd_parameters, dtype=np.float64)
    ini = np.ascontiguousarray(nd_initial, dtype=np.float64)
    length_simu, report_type, report_gap = int(length_simu), int(report_type), int(report_gap)
    if rain.size < length_simu or peva.size < length_simu:
        raise IndexError("forcing arrays are shorter than length_simu")
    if par.shape != (_native.N_PARAMS,) or ini.shape != (_native.N_VARS,):
        raise ValueError("nd_parameters must have 10 values and nd_initial 19")
    n_rep = length_simu // report_gap if report_type == _native.REPORT_SUMMARY else -(-length_simu // report_gap)
    discharge = np.zeros(max(n_rep, 0), dtype=np.float64)
    gw = np.zeros(1, dtype=np.float64)
    last = np.zeros(_native.N_VARS, dtype=np.float64)
    rc = lib.smart_allsteps_host(float(area_m2), float(delta_sec), length_simu, _ptr(rain), _ptr(peva),
                                 _ptr(par), _ptr(ini), report_type, report_gap,
                                 _ptr(discharge), _ptr(gw), _ptr(last), int(device))
    _native.check(rc)
    return discharge, float(gw[0]), last


def onestep(area_m2, time_delta_sec, c_in_rain, c_in_peva,
            c_p_t, c_p_c, c_p_h, c_p_d, c_p_s, c_p_z, c_p_sk, c_p_fk, c_p_gk, r_p_rk,
            c_s_v_h2o_ove, c_s_v_h2o_dra, c_s_v_h2o_int, c_s_v_h2o_sgw, c_s_v_h2o_dgw,
            c_s_v_h2o_ly1, c_s_v_h2o_ly2, c_s_v_h2o_ly3, c_s_v_h2o_ly4, c_s_v_h2o_ly5, c_s_v_h2o_ly6,
            r_s_v_riv):
    """One model step -> the 19-tuple of structure.py:259-264 (a 1-step run on the device)."""
    initial = np.zeros(_native.N_VARS, dtype=np.float64)
    initial[7:] = [c_s_v_h2o_ove, c_s_v_h2o_dra, c_s_v_h2o_int, c_s_v_h2o_sgw, c_s_v_h2o_dgw,
                   c_s_v_h2o_ly1, c_s_v_h2o_ly2, c_s_v_h2o_ly3, c_s_v_h2o_ly4, c_s_v_h2o_ly5, c_s_v_h2o_ly6,
                   r_s_v_riv]
    params = np.array([c_p_t, c_p_c, c_p_h, c_p_d, c_p_s, c_p_z, c_p_sk, c_p_fk, c_p_gk, r_p_rk], dtype=np.float64)
    _, _, last = allsteps(area_m2, time_delta_sec, 1, np.array([c_in_rain], dtype=np.float64),
                          np.array([c_in_peva], dtype=np.float64), params, initial, 2, 1)
    return tuple(last.tolist())
