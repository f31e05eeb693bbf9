"""CPU: the CUDA step header (smartpy_b200/csrc/smart_step.cuh) compiled for the HOST with g++
(tools/host_emulate.cpp, a development aid -- never loaded by smartpy_b200) against the
reference-generated goldens.  This checks the arithmetic of the three step formulations (general,
per-step fast, block mode) without a GPU; the GPU tests check the kernels themselves.  On B200
the kernels reproduce these host numbers to the last digit (profiles/r01_parity_report.txt)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libemul.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-DSMART_HOST_EMULATION", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tools", "host_emulate.cpp")])
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.emulate_run.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_long, ctypes.c_long, dp, dp, dp,
                                ctypes.c_int, ctypes.c_double, dp, ctypes.c_int, ctypes.c_int, dp, dp]
    return lib


@pytest.mark.parametrize("mode,label,bound", [(0, "general", 2e-12), (2, "per-step fast", 4e-12), (3, "block mode", 8e-12)])
def test_step_formulations_match_reference_goldens(emulator, catchment, mode, label, bound):
    dp = ctypes.POINTER(ctypes.c_double)
    g = load_golden("runs_members")
    split = np.array([0.10, 0.15, 0.15, 0.30, 0.30])
    rain, peva = np.ascontiguousarray(catchment.rain), np.ascontiguousarray(catchment.peva)
    worst_q = worst_gw = 0.0
    for i, p in enumerate(g["params"]):
        q = np.zeros(3653)
        gw = ctypes.c_double()
        p = np.ascontiguousarray(p)
        emulator.emulate_run(mode, catchment.area, 3600.0, 87672, 8760, rain.ctypes.data_as(dp), peva.ctypes.data_as(dp),
                             p.ctypes.data_as(dp), 1, 1200 * 0.45, split.ctypes.data_as(dp), 1, 24,
                             q.ctypes.data_as(dp), ctypes.byref(gw))
        worst_q = max(worst_q, float(np.max(np.abs(q - g["q"][i]) / g["q"][i])))
        worst_gw = max(worst_gw, abs(gw.value - g["gw"][i]) / g["gw"][i])
    assert worst_q < bound, (label, worst_q)        # far inside the 1e-10 bar of BASELINE.json
    assert worst_gw < 1e-11


@pytest.mark.parametrize("gap,report,n_days", [(24, "raw", 3653), (1, "raw", 500), (1, "summary", 300), (6, "summary", 500),
                                                (8, "raw", 500)])
def test_block_sub_mode_matches_oracle(emulator, catchment, oracle_lib, gap, report, n_days):
    """Blocks of 24 constant-forcing steps with reports INSIDE the block (kModeBlockSub of
    run_timeline: hourly output of a model forced with daily totals, or 'raw' reporting), against
    the oracle's hour-by-hour run (structure.py:181-195 handles every gap with one loop)."""
    dp = ctypes.POINTER(ctypes.c_double)
    emulator.emulate_run_sub.argtypes = [ctypes.c_int] + emulator.emulate_run.argtypes[1:]
    g = load_golden("runs_members")
    split = np.array([0.10, 0.15, 0.15, 0.30, 0.30])
    n = 24 * n_days
    rain, peva = np.ascontiguousarray(catchment.rain[:n]), np.ascontiguousarray(catchment.peva[:n])
    rtype = 1 if report == "summary" else 2
    worst_q = worst_gw = 0.0
    for i in (0, 3, 11, 17, 25, 33, 39):
        p = np.ascontiguousarray(g["params"][i])
        q = np.zeros(n // gap)
        gw = ctypes.c_double()
        emulator.emulate_run_sub(24, catchment.area, 3600.0, n, 24 * 30, rain.ctypes.data_as(dp), peva.ctypes.data_as(dp),
                                 p.ctypes.data_as(dp), 1, 1200 * 0.45, split.ctypes.data_as(dp), rtype, gap,
                                 q.ctypes.data_as(dp), ctypes.byref(gw))
        q_ref, gw_ref = oracle_lib.run(catchment.area, 3600.0, rain, peva, p, catchment.extra, n, gap, report=report,
                                       warm_up=30)
        assert q_ref.shape == q.shape
        worst_q = max(worst_q, float(np.max(np.abs(q - q_ref) / q_ref)))
        worst_gw = max(worst_gw, abs(gw.value - gw_ref) / gw_ref)
    assert worst_q < 1e-11, worst_q
    assert worst_gw < 1e-11, worst_gw


@pytest.mark.parametrize("mode,label", [(3, "block mode"), (2, "per-step fast")])
def test_binary32_state_keeps_nse_kge_within_the_bar(emulator, catchment, mode, label):
    """The FP32 mode of the kernels (binary32 stores, the soil kept as deficits z - level, see
    fast_wet_soil_deficit) compiled for the host: NSE and KGE of the 40 golden members against the
    reference's own discharge within 1e-5 absolute (BASELINE.json north_star).  The GPU unit may
    contract a * b + c differently, so this pins the formulation, not the bits; the GPU tests
    repeat it over the whole C2 batch and at the C3 length."""
    from oracle import scores as oscores
    dp = ctypes.POINTER(ctypes.c_double)
    emulator.emulate_run_f32.argtypes = emulator.emulate_run.argtypes
    g = load_golden("runs_members")
    split = np.array([0.10, 0.15, 0.15, 0.30, 0.30])
    rain, peva = np.ascontiguousarray(catchment.rain), np.ascontiguousarray(catchment.peva)
    worst_nse = worst_kge = worst_q = 0.0
    for i, p in enumerate(g["params"]):
        q = np.zeros(3653)
        gw = ctypes.c_double()
        p = np.ascontiguousarray(p)
        emulator.emulate_run_f32(mode, catchment.area, 3600.0, 87672, 8760, rain.ctypes.data_as(dp),
                                 peva.ctypes.data_as(dp), p.ctypes.data_as(dp), 1, 1200 * 0.45, split.ctypes.data_as(dp),
                                 1, 24, q.ctypes.data_as(dp), ctypes.byref(gw))
        ref = oscores.objectivefunction((g["q"][i], [g["gw"][i]]), (catchment.obs, [None]))
        got = oscores.objectivefunction((q, [gw.value]), (catchment.obs, [None]))
        worst_nse = max(worst_nse, abs(got[0] - ref[0]))
        worst_kge = max(worst_kge, abs(got[1] - ref[1]))
        floor = 1e-6 * g["q"][i].max()
        worst_q = max(worst_q, float(np.max(np.abs(q - g["q"][i]) / np.maximum(g["q"][i], floor))))
    assert worst_nse < 1e-5 and worst_kge < 1e-5, (label, worst_nse, worst_kge)
    assert worst_q < 5e-3, (label, worst_q)      # low flows of the range-corner members (floor: 1e-6 of the peak)
