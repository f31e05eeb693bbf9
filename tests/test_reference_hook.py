"""CPU, build container only (skipped where /root/reference does not exist): the reference's own
`smartcpp` hook (smartpy/structure.py:22-27, :56-62, :118-121, :143-146) accepts a module with the
call contract that smartpy_b200.smartcpp_shim implements.  The GPU cannot run here, so the hook is
driven with a stand-in that has the shim's exact signatures and forwards to the CPU oracle; the
unmodified reference must then reproduce its own golden discharge through it."""
import importlib
import inspect
import os
import sys
import types

import numpy as np
import pytest

from conftest import load_golden, EXTRA

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "smartpy")),
                                reason="the reference is only mounted in the build container")


def test_unmodified_reference_runs_through_the_shim_contract(catchment, oracle_lib):
    from smartpy_b200 import smartcpp_shim
    calls = []

    def allsteps(area_m2, delta_sec, length_simu, nd_rain, nd_peva, nd_parameters, nd_initial,
                 report_type, report_gap):
        calls.append((length_simu, report_type, report_gap, len(nd_rain)))
        return oracle_lib.allsteps(area_m2, delta_sec, length_simu, nd_rain, nd_peva, nd_parameters, nd_initial,
                                   report_type, report_gap)

    # same positional parameters as the shim (which adds only an optional device ordinal)
    shim_params = list(inspect.signature(smartcpp_shim.allsteps).parameters)
    assert shim_params[:9] == list(inspect.signature(allsteps).parameters)
    assert len(inspect.signature(smartcpp_shim.onestep).parameters) == 26

    fake = types.ModuleType("smartcpp")
    fake.allsteps = allsteps
    saved = {k: sys.modules.get(k) for k in ("smartcpp", "smartpy", "smartpy.structure")}
    sys.dont_write_bytecode = True
    sys.path.insert(0, REFERENCE)
    try:
        for k in [k for k in sys.modules if k == "smartpy" or k.startswith("smartpy.")]:
            del sys.modules[k]
        sys.modules["smartcpp"] = fake
        structure = importlib.import_module("smartpy.structure")
        assert structure.smart_in_cpp
        g = load_golden("runs_single")
        from datetime import timedelta
        timeseries = [None] * (87672 + 1)
        timeseries_report = [None] * (3653 + 1)
        q, gw = structure.run(catchment.area, timedelta(hours=1), catchment.rain, catchment.peva, g["p_test"], EXTRA,
                              timeseries, timeseries_report, report='summary', warm_up=365)
        assert np.array_equal(q, g["q_summary"])
        assert calls == [(8760, 1, 24, 87672), (87672, 1, 24, 87672)]     # warm-up, then the main run
    finally:
        sys.path.remove(REFERENCE)
        for k in [k for k in sys.modules if k == "smartpy" or k.startswith("smartpy.")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
