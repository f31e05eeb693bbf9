"""GPU: the device Latin Hypercube sampler (smart_lhs_rows) against a numpy
restatement of the algorithm written in csrc/smart_sample.cu -- bit for bit (integer work + four
IEEE operations) -- and the properties the reference's sampler has (lhs.py:133-167): one value per
stratum in every column, values inside the bounds, row ranges generated independently."""
import numpy as np
import pytest

from test_host_logic import catchment_dir  # noqa: F401

pytestmark = pytest.mark.gpu

U = np.uint64


def _mix(z):
    z = z.astype(np.uint64)
    z ^= z >> U(30)
    z *= U(0xBF58476D1CE4E5B9)
    z ^= z >> U(27)
    z *= U(0x94D049BB133111EB)
    z ^= z >> U(31)
    return z


def lhs_restated(seed, n_total, row_first, n_rows, bounds):
    bounds = np.asarray(bounds, dtype=np.float64)
    h = 1
    while 4 ** h < n_total:
        h += 1
    mask = U((1 << h) - 1)
    rows = np.arange(row_first, row_first + n_rows, dtype=np.uint64)
    out = np.empty((n_rows, bounds.shape[0]))
    with np.errstate(over='ignore'):
        for p in range(bounds.shape[0]):
            key = _mix(np.array([(seed + 0x9E3779B97F4A7C15 * (p + 1)) % 2 ** 64], dtype=np.uint64))[0]
            x = rows.copy()
            todo = np.ones(n_rows, dtype=bool)
            while todo.any():
                left, right = x[todo] >> U(h), x[todo] & mask
                for rnd in range(6):
                    f = _mix(key ^ (right + U((0xD6E8FEB86659FD93 * (rnd + 1)) % 2 ** 64))) & mask
                    left, right = right, left ^ f
                x[todo] = (left << U(h)) | right
                todo = x >= U(n_total)
            jitter = (_mix(key ^ _mix(rows + U(0x632BE59BD9B4E019))) >> U(11)).astype(np.float64) * 2.0 ** -53
            q = (x.astype(np.float64) + jitter) / float(n_total)
            out[:, p] = q * (bounds[p, 1] - bounds[p, 0]) + bounds[p, 0]
    return out


def _bounds():
    from smartpy_b200.parameters import Parameters
    p = Parameters()
    return [list(p.ranges[n]) for n in p.names]


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (5, 7), (2000, 5), (65536, 42), (100003, 2 ** 63 + 11)])
def test_device_lhs_matches_restatement_and_is_stratified(n, seed):
    from smartpy_b200.montecarlo.conditioning import latin_hypercube_device
    bounds = _bounds()
    sample = latin_hypercube_device(n, bounds, seed=seed).cpu().numpy()
    assert sample.shape == (n, 10)
    assert np.array_equal(sample, lhs_restated(seed, n, 0, n, bounds))
    for k, (lo, hi) in enumerate(bounds):
        assert sample[:, k].min() >= lo and sample[:, k].max() <= hi
        strata = np.floor((sample[:, k] - lo) / (hi - lo) * n).astype(np.int64)
        strata = np.clip(strata, 0, n - 1)      # (s + u)/n*w + lo rounds; the exact check is the one above
        assert np.array_equal(np.sort(strata), np.arange(n)), k


def test_device_lhs_row_ranges_tile_the_sample():
    """What a sharded run relies on: rank r's rows are exactly rows lo..hi of the whole sample."""
    from smartpy_b200.montecarlo.conditioning import latin_hypercube_device
    from smartpy_b200.distributed import shard_bounds
    bounds, n = _bounds(), 30011
    whole = latin_hypercube_device(n, bounds, seed=3).cpu().numpy()
    for world in (2, 8):
        parts = []
        for rank in range(world):
            lo, hi = shard_bounds(n, rank, world)
            parts.append(latin_hypercube_device(n, bounds, seed=3, row_first=lo, n_rows=hi - lo).cpu().numpy())
        assert np.array_equal(np.concatenate(parts), whole)
    assert not np.array_equal(whole, latin_hypercube_device(n, bounds, seed=4).cpu().numpy())


def test_device_lhs_columns_look_independent():
    """The keyed bijections of different columns must not be correlated (a broken key schedule
    would put all columns on the same permutation)."""
    from smartpy_b200.montecarlo.conditioning import latin_hypercube_device
    n = 200000
    sample = latin_hypercube_device(n, [[0.0, 1.0]] * 10, seed=9).cpu().numpy()
    corr = np.corrcoef(sample.T)
    off = corr[~np.eye(10, dtype=bool)]
    assert np.abs(off).max() < 5.0 / np.sqrt(n)
    # and successive rows of one column are not ordered: lag-1 autocorrelation ~ 0
    for k in range(10):
        c = np.corrcoef(sample[:-1, k], sample[1:, k])[0, 1]
        assert abs(c) < 5.0 / np.sqrt(n)
    # 2-D occupancy: a 10 x 10 grid over columns (0, 1) is filled uniformly (chi-square, 99 dof)
    cells = np.floor(sample[:, 0] * 10).astype(int) * 10 + np.floor(sample[:, 1] * 10).astype(int)
    counts = np.bincount(cells, minlength=100)
    chi2 = ((counts - n / 100.0) ** 2 / (n / 100.0)).sum()
    assert chi2 < 180.0


def test_lhs_class_with_device_sampler(catchment_dir):  # noqa: F811
    from smartpy_b200 import montecarlo
    setup = montecarlo.LHS('Catchment', catchment_dir, 'csv', 'csv', sample_size=64, sampler='device', seed=12)
    assert setup.sample_params.shape == (64, 10)
    assert np.array_equal(setup.sample_params, lhs_restated(12, 64, 0, 64, _bounds()))
    setup.run()
    assert setup.results['scores'].shape[0] == 64


def test_device_lhs_argument_errors():
    from smartpy_b200.montecarlo.conditioning import latin_hypercube_device
    with pytest.raises(Exception, match="smart_lhs_rows"):
        latin_hypercube_device(10, _bounds(), row_first=5, n_rows=6)
    with pytest.raises(Exception, match="smart_lhs_rows"):
        latin_hypercube_device(10, [[0.0, 1.0]] * 17)
