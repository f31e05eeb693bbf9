"""GPU: the conditioning kernels (smart_condition_rows / smart_best_rows)
against the numpy rules of the reference kept in GLUE/Best (glue.py:246-289, best.py:243-287):
index-exact, including ties, NaN scores, -0.0, constraints and k from 1 to the whole sample."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAMES = ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW']


def _table(n, seed, ties=False):
    rng = np.random.RandomState(seed)
    scores = rng.randn(n, 8)
    scores[:, 7] = rng.rand(n) > 0.5
    if ties:
        scores[:, 1] = np.round(scores[:, 1], 1)          # many equal values on the target
        scores[rng.randint(0, n, max(1, n // 50)), 1] = np.nan
        scores[rng.randint(0, n, max(1, n // 50)), 1] = -0.0
        scores[rng.randint(0, n, max(1, n // 50)), 1] = 0.0
        scores[rng.randint(0, n, 3), 1] = np.inf
        scores[rng.randint(0, n, 3), 1] = -np.inf
    return scores


def _best_ref(scores, mask, col, k):
    kept = np.nonzero(mask)[0]
    order = np.argsort(scores[kept, col], kind='stable')      # NaN last, ties by row
    return kept[order][-k:]


@pytest.mark.parametrize("n", [1, 31, 1000, 4097, 300000])
def test_behavioural_rows_match_numpy_rules(n):
    import torch
    from smartpy_b200.montecarlo import conditioning
    from smartpy_b200.montecarlo.montecarlo import condition_mask
    scores = _table(n, 3 + n)
    scores[::7, 0] = np.nan
    t = torch.from_numpy(scores).cuda()
    cases = [
        {'NSE': ('min', (0.2,)), 'PBias': ('inside', (-1.0, 1.0)), 'GW': ('equal', (1.0,))},
        {'KGE': ('max', (0.0,))},
        {'RMSE': ('outside', (-0.5, 0.5))},                    # the reference's literal rule: nothing passes
        {'NSE': ('min', (-100.0,))},
        {},
    ]
    for cond in cases:
        cols = [NAMES.index(k) for k in cond]
        mask = condition_mask(scores[:, cols], [cond[k][1] for k in cond], [cond[k][0] for k in cond])
        rows = conditioning.behavioural_rows(t, NAMES, cond).cpu().numpy()
        assert np.array_equal(rows, np.nonzero(mask)[0]), cond
    # a column slice of a wider table (what MonteCarlo.results holds) is used in place
    rows = conditioning.behavioural_rows(t[:, :7], NAMES[:7], {'KGE': ('max', (0.0,))}).cpu().numpy()
    assert np.array_equal(rows, np.nonzero(scores[:, 1] <= 0.0)[0])


@pytest.mark.parametrize("n,ks", [(1, [1]), (50, [1, 2, 49, 50]), (5000, [1, 7, 1024, 2048, 2049, 5000]),
                                  (300000, [1, 100, 4096, 70000, 300000])])
def test_best_rows_match_stable_argsort(n, ks):
    import torch
    from smartpy_b200.montecarlo import conditioning
    from smartpy_b200.montecarlo.montecarlo import condition_mask
    for ties in (False, True):
        scores = _table(n, 11 + n, ties=ties)
        t = torch.from_numpy(scores).cuda()
        for k in ks:
            rows = conditioning.best_rows(t, NAMES, 'KGE', k).cpu().numpy()
            assert np.array_equal(rows, _best_ref(scores, np.ones(n, bool), 1, k)), (n, k, ties)
        cond = {'RMSE': ('max', (0.5,)), 'GW': ('equal', (1.0,))}
        mask = condition_mask(scores[:, [6, 7]], [(0.5,), (1.0,)], ['max', 'equal'])
        kept = int(mask.sum())
        for k in sorted({1, max(1, kept // 3), kept}):
            if kept == 0:
                continue
            rows = conditioning.best_rows(t, NAMES, 'KGE', k, cond).cpu().numpy()
            assert np.array_equal(rows, _best_ref(scores, mask, 1, k)), (n, k, ties, 'constrained')
        if kept < n:
            with pytest.raises(Exception, match="restrained sample size"):
                conditioning.best_rows(t, NAMES, 'KGE', kept + 1, cond)


def test_best_rows_agree_with_the_file_based_best_class():
    import torch
    from smartpy_b200.montecarlo import conditioning
    from smartpy_b200.montecarlo.best import Best
    scores = _table(500, 3)
    params = np.random.RandomState(4).rand(500, 10)
    t = torch.from_numpy(scores).cuda()
    rows = conditioning.best_rows(t, NAMES, 'KGE', 7, {'RMSE': ('max', (0.5,))}).cpu().numpy()
    ref = Best._get_best_sets(params, scores[:, [6]], [(0.5,)], ['max'], scores[:, [1]], 7)
    assert np.array_equal(params[rows], ref)


def test_select_is_stream_ordered_and_repeatable():
    """Same answer when the two selections are queued back to back on a side stream."""
    import torch
    from smartpy_b200.montecarlo import conditioning
    scores = _table(100000, 9, ties=True)
    t = torch.from_numpy(scores).cuda()
    first = conditioning.best_rows(t, NAMES, 'KGE', 5000).cpu().numpy()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        again = conditioning.best_rows(t, NAMES, 'KGE', 5000)
        rows = conditioning.behavioural_rows(t, NAMES, {'GW': ('equal', (1.0,))})
    side.synchronize()
    assert np.array_equal(again.cpu().numpy(), first)
    assert np.array_equal(rows.cpu().numpy(), np.nonzero(scores[:, 7] == 1.0)[0])
