"""GPU: randomised parity sweep -- parameter sets drawn from ranges wider than the reference's
defaults (including sets that must take the branch-faithful kernel), synthetic forcing with dry
spells, storms and zero-PET days, block-constant and genuinely hourly, against the oracle."""
import numpy as np
import pytest

from conftest import EXTRA

pytestmark = pytest.mark.gpu


def _params(rng, n, wild_fraction):
    lo = np.array([0.8, 0.0, 0.0, 0.0, 0.0, 5.0, 1.0, 1.5, 2.0, 1.0])
    hi = np.array([1.2, 1.0, 0.5, 1.0, 0.05, 300.0, 300.0, 2000.0, 6000.0, 120.0])
    p = lo + rng.rand(n, 10) * (hi - lo)
    p[rng.rand(n) < 0.1, 1] = 0.0          # C = 0: nothing is taken below an empty layer
    p[rng.rand(n) < 0.1, 1] = 1.0
    p[rng.rand(n) < 0.1, 4] = 0.0          # S = 0: no leaks at all
    p[rng.rand(n) < 0.1, 2] = 0.0          # H = 0
    wild = rng.rand(n) < wild_fraction     # outside the fast form's validity
    p[wild, 6:10] = rng.uniform(0.1, 2.0, (int(wild.sum()), 4))      # routing constants around / below dt
    p[wild, 4] = rng.uniform(0.0, 3.0, int(wild.sum()))              # S up to 3: s' can exceed 1
    return p


def _forcing(rng, days, block_constant):
    wet = rng.rand(days) < 0.6
    rain_d = np.where(wet, rng.gamma(0.7, 6.0, days), 0.0)
    rain_d[rng.rand(days) < 0.03] *= 15.0                              # storms: saturate every layer
    peva_d = np.maximum(0.0, 1.5 + 1.5 * np.sin(np.arange(days) / 58.0) + rng.normal(0, 0.5, days))
    dry_spell = slice(days // 3, days // 3 + 45)
    rain_d[dry_spell] = 0.0                                            # long enough to empty the soil
    peva_d[dry_spell] += 3.0
    rain, peva = np.repeat(rain_d / 24, 24), np.repeat(peva_d / 24, 24)
    if not block_constant:
        rain = rain * rng.uniform(0.0, 2.0, rain.size)
        peva = peva * rng.uniform(0.5, 1.5, peva.size)
    return np.ascontiguousarray(rain), np.ascontiguousarray(peva)


@pytest.mark.parametrize("block_constant,seed", [(True, 1), (True, 2), (False, 3)])
def test_random_members_and_forcing_match_oracle(oracle_lib, block_constant, seed):
    import torch
    assert torch.cuda.is_available()
    from smartpy_b200.engine import BatchEngine
    rng = np.random.RandomState(seed)
    days, n = 150, 320
    rain, peva = _forcing(rng, days, block_constant)
    params = _params(rng, n, wild_fraction=0.15)
    area = 3.3e8
    eng = BatchEngine(rain, peva, area, 3600.0, 24, extra=EXTRA, warm_up_steps=24 * 20)
    assert (eng._repeat == 24) == block_constant
    res = eng.run(params, discharge=True, scores=False, gw=True)
    q = res["discharge"].cpu().numpy().T
    gw = res["gw"].cpu().numpy()
    q_ref, gw_ref = oracle_lib.run_members(area, 3600.0, rain, peva, params, EXTRA, days * 24, 24, warm_up=20)
    # relative to each value, with a floor of 1e-9 of the member's peak flow (values that have
    # decayed by many orders of magnitude carry no relative information)
    scale = np.maximum(np.abs(q_ref), 1e-9 * np.abs(q_ref).max(axis=1, keepdims=True))
    err = np.abs(q - q_ref) / scale
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() < 1e-10, (err.max(), worst, params[worst[0]])
    ok = np.isfinite(gw_ref)
    assert np.max(np.abs(gw[ok] - gw_ref[ok]) / np.maximum(gw_ref[ok], 1e-12)) < 1e-8
