"""GPU: the relay (include/smart_b200.h, `workspace_bytes`; smart_batch_kernel in smart_kernels.cu).

A batch of about one wave runs as a relay: the timeline is cut into segments, a CTA advances one
group of members through one segment and parks the group's state for whichever CTA takes the next
one.  The numbers parked are the numbers an uninterrupted walk carries in registers and shared
memory, so a member's result must be the SAME BITS however the timeline is cut -- in every mode of
the time loop (block, block-sub, per-step), both precisions, both step forms, with or without
discharge output, with the best-member search -- and equal to the reference's own answers.
"""
import numpy as np
import pytest

from conftest import load_golden
from test_gpu_parity import make_engine, relmax, RTOL_Q

pytestmark = pytest.mark.gpu


def _flags(segs, extra=0):
    from smartpy_b200 import _native
    from smartpy_b200.engine import FLAG_NO_RELAY
    return (FLAG_NO_RELAY if segs == 0 else _native.flag_relay_segs(segs)) | extra


def _run(catchment, params, segs, extra_flags=0, **kw):
    run_kw = dict(discharge=kw.pop("discharge", True), scores=kw.pop("scores", True), gw=True)
    best = kw.pop("best", None)
    eng = make_engine(catchment, flags=_flags(segs, extra_flags), **kw)
    res = eng.run(params, best=best, **run_kw)
    out = {k: v.cpu().numpy() for k, v in res.items() if k != "best"}
    if best:
        out["best"] = (float(res["best"][0].cpu()), int(res["best"][1].cpu()))
    return out


def _same(a, b):
    for k in a:
        if k == "best":
            assert a[k] == b[k]
        else:
            assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("segs", [2, 3, 7, 64, 255])
def test_relay_block_mode_same_bits_and_reference_answers(catchment, segs):
    """Block mode (C2's mode): 40 members with reference-generated answers, any number of segments."""
    g = load_golden("runs_members")
    plain = _run(catchment, g["params"], 0)
    relay = _run(catchment, g["params"], segs)
    _same(plain, relay)
    assert relmax(relay["discharge"].T, g["q"]) < RTOL_Q


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("mode", ["perstep", "general", "raw", "nowarm", "no_tma"])
def test_relay_same_bits_in_every_mode(catchment, precision, mode):
    from smartpy_b200 import _native
    from smartpy_b200.engine import FLAG_NO_BLOCK_MODE
    g = load_golden("runs_members")
    extra, kw = 0, {}
    if mode == "perstep":
        extra = FLAG_NO_BLOCK_MODE
    elif mode == "general":
        extra = _native.FLAG_FORCE_GENERAL
    elif mode == "raw":
        kw = dict(report="raw")
    elif mode == "nowarm":
        kw = dict(warm_up_days=0)
    elif mode == "no_tma":
        extra = _native.FLAG_NO_TMA | FLAG_NO_BLOCK_MODE
    plain = _run(catchment, g["params"], 0, extra, precision=precision, **kw)
    for segs in (2, 5, 33):
        _same(plain, _run(catchment, g["params"], segs, extra, precision=precision, **kw))


def test_relay_hourly_reports_inside_constant_forcing_days(catchment):
    """Block-sub mode with the output cursor (hourly discharge of a model forced with daily totals)."""
    from smartpy_b200.engine import BatchEngine
    g = load_golden("runs_members")
    n = 24 * 400

    def run(segs):
        eng = BatchEngine(catchment.rain[:n], catchment.peva[:n], catchment.area, catchment.dt, 1, extra=catchment.extra,
                          warm_up_steps=24 * 30, report="raw", flags=_flags(segs))
        assert eng._repeat == 24
        res = eng.run(g["params"], discharge=True, scores=False, gw=True)
        return {k: v.cpu().numpy() for k, v in res.items()}

    plain = run(0)
    for segs in (2, 6):
        _same(plain, run(segs))


def test_relay_with_grouped_members_wild_members_and_best_member(catchment):
    """A batch large enough for the member grouping (idle slots, members that need the branch-faithful
    form in CTAs of their own: two kernels relay side by side), with the best-member search."""
    import bench
    n = 6000
    params = bench.lhs_rows(n, 3)
    params[17::400, 6] = 0.1          # SK * 3600 < dt: outside the merged form's domain
    kw = dict(discharge=False, best=("NSE", 1))
    plain = _run(catchment, params, 0, **kw)
    for segs in (4, 32):
        _same(plain, _run(catchment, params, segs, **kw))
    assert plain["best"][1] == int(np.nanargmax(plain["scores"][:, 0]))


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_relay_is_what_a_c2_sized_batch_runs_and_changes_no_bit(catchment, precision):
    """C2's size (C5 with binary32 state): the library chooses the relay on its own (no flag); same
    bits as the plain launch."""
    import bench
    n = 100000
    params = bench.lhs_rows(n, 42)
    auto = _run(catchment, params, 0, discharge=False, precision=precision)      # plain launch
    eng = make_engine(catchment, precision=precision)                            # library's choice
    res = eng.run(params, discharge=False, scores=True, gw=True)
    assert np.array_equal(res["scores"].cpu().numpy(), auto["scores"], equal_nan=True)
    assert np.array_equal(res["gw"].cpu().numpy(), auto["gw"])


def test_relay_through_the_host_entry_point(catchment):
    """smart_batch_run_host sizes its arena with smart_batch_workspace_bytes: relay by flag, same bits."""
    import ctypes
    from smartpy_b200 import _native
    lib = _native.load()
    g = load_golden("runs_members")
    params = np.ascontiguousarray(g["params"], dtype=np.float64)
    n, T = params.shape[0], catchment.n_steps
    rain, peva = np.ascontiguousarray(catchment.rain), np.ascontiguousarray(catchment.peva)
    area = np.array([catchment.area])

    def run(flags):
        d = _native.BatchDesc()
        d.n_members, d.n_steps, d.n_warmup = n, T, 8760
        d.n_catchments, d.members_per_catchment = 1, 1
        d.report_gap, d.report_type, d.flags, d.dt_sec = 24, _native.REPORT_SUMMARY, flags, 3600.0
        d.params, d.rain, d.peva, d.area_m2 = (x.ctypes.data for x in (params, rain, peva, area))
        d.has_extra, d.aar, d.ro_ratio = 1, 1200.0, 0.45
        for k, v in enumerate((0.10, 0.15, 0.15, 0.30, 0.30)):
            d.ro_split[k] = v
        d.gw_constraint = float("nan")
        q = np.empty((T // 24, n))
        gw = np.empty(n)
        d.discharge, d.ld_discharge, d.gw = q.ctypes.data, n, gw.ctypes.data
        _native.check(lib.smart_batch_run_host(ctypes.byref(d), 64, 0))
        return q, gw

    q0, gw0 = run(0)
    q1, gw1 = run(_native.flag_relay_segs(9))
    assert np.array_equal(q0, q1) and np.array_equal(gw0, gw1)
    assert relmax(q1.T, g["q"]) < RTOL_Q
