"""The C ABI from plain C (examples/abi_client.c): builds and links on the CPU box; on the GPU it
runs a case through smart_batch_run_host (per-step and block-mode forcing) against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden, EXTRA


@pytest.fixture(scope="module")
def client(tmp_path_factory):
    from smartpy_b200 import _build
    _build.build()
    exe = str(tmp_path_factory.mktemp("abi") / "abi_client")
    libdir = os.path.join(ROOT, "smartpy_b200")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "examples", "abi_client.c"), "-L", libdir, "-lsmart_b200",
                           "-Wl,-rpath," + libdir, "-lm"])
    return exe


def test_c_client_links_and_validates(client):
    out = subprocess.check_output([client], text=True)
    assert "smart_version 200" in out and "empty descriptor -> -1" in out


@pytest.mark.gpu
@pytest.mark.parametrize("repeat", [1, 24])
def test_c_client_matches_oracle(client, catchment, oracle_lib, tmp_path, repeat):
    from oracle import scores as oscores
    g = load_golden("runs_members")
    n, days, warm = 7, 300, 24 * 30
    params = np.ascontiguousarray(g["params"][:n])
    rain, peva = catchment.rain[:days * 24], catchment.peva[:days * 24]
    obs = np.ascontiguousarray(catchment.obs[:days])
    case, result = tmp_path / "case.bin", tmp_path / "result.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("7q", n, days * 24, warm, 24, 1, repeat, 1))
        f.write(struct.pack("3d", 3600.0, catchment.area, float("nan")))
        f.write(params.tobytes())
        f.write(np.ascontiguousarray(rain[::repeat]).tobytes())     # one row per block of `repeat` steps
        f.write(np.ascontiguousarray(peva[::repeat]).tobytes())
        f.write(obs.tobytes())
    subprocess.check_call([client, str(case), str(result)])
    raw = np.fromfile(result, dtype=np.float64)
    q = raw[:days * n].reshape(days, n).T
    gw = raw[days * n:days * n + n]
    sc = raw[days * n + n:].reshape(n, 8)
    q_ref, gw_ref = oracle_lib.run_members(catchment.area, 3600.0, rain, peva, params, EXTRA, days * 24, 24, warm_up=30)
    assert np.max(np.abs(q - q_ref) / q_ref) < 1e-10
    assert np.max(np.abs(gw - gw_ref) / gw_ref) < 1e-9
    sc_ref = oscores.score_members(q_ref, gw_ref, obs, None)
    assert np.max(np.abs(sc[:, :7] - sc_ref[:, :7]) / np.maximum(1.0, np.abs(sc_ref[:, :7]))) < 1e-10
    assert np.isnan(sc[:, 7]).all()
