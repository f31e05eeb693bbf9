"""Conditioning of a scored sample on the device.

The reference conditions a sample after a round trip through a float32 text/NetCDF file
(``montecarlo.py:233-262``) with numpy masks and a full ``argsort`` (``glue.py:246-289``,
``best.py:243-287``).  With the scores already on the GPU (``MonteCarlo.results``) the same
selections are a handful of tensor operations; the functions below work on any torch tensor
(CUDA or CPU) and follow the reference's rules, including its tie order (ascending sort, best
last) and its literal 'outside' rule.
"""
import numpy as np


def _torch():
    import torch
    return torch


def condition_mask_tensor(scores, columns, conditions_val, conditions_typ):
    """Boolean tensor over the rows of scores[N, k']: AND of the (kind, values) conditions on the
    given columns.  Same kinds and error messages as glue.py:246-286."""
    torch = _torch()
    mask = torch.ones((scores.shape[0],), dtype=torch.bool, device=scores.device)
    for col, values, kind in zip(columns, conditions_val, conditions_typ):
        x = scores[:, col]
        if kind in ('equal', 'min', 'max'):
            if len(values) != 1:
                raise Exception("The tuple for \"{}\" condition does not contain one and only one "
                                "element.".format(kind))
            sel = (x == values[0]) if kind == 'equal' else (x >= values[0]) if kind == 'min' else (x <= values[0])
        elif kind in ('inside', 'outside'):
            if len(values) != 2:
                raise Exception("The tuple for \"{}\" condition does not contain two and only two "
                                "elements.".format(kind))
            if not values[1] > values[0]:
                raise Exception("The two elements of the tuple for \"{}\" are inconsistent.".format(kind))
            if kind == 'inside':
                sel = (x >= values[0]) & (x <= values[1])
            else:
                sel = (x <= values[0]) & (x >= values[1])
        else:
            raise Exception("The type of threshold \"{}\" is not in the database.".format(kind))
        mask &= sel
    return mask


def behavioural_rows(scores, obj_fn_names, conditioning):
    """Row indices (ascending) of the behavioural sets -- GLUE (glue.py:222-289)."""
    torch = _torch()
    try:
        columns = [obj_fn_names.index(fn) for fn in conditioning]
    except ValueError:
        raise Exception("One of the names of objective functions for conditioning in GLUE is not recognised."
                        "Please check for typos and case sensitive issues.")
    mask = condition_mask_tensor(scores, columns, [conditioning[fn][1] for fn in conditioning],
                                 [conditioning[fn][0] for fn in conditioning])
    return torch.nonzero(mask, as_tuple=False)[:, 0]


def best_rows(scores, obj_fn_names, target, nb_best, constraining=None):
    """Row indices of the nb_best sets on `target` after the optional constraints, in the
    reference's order: ascending on the target, best last (best.py:221-287)."""
    torch = _torch()
    try:
        t_col = obj_fn_names.index(target)
    except ValueError:
        raise Exception("The objective function {} for conditioning in Best is not recognised."
                        "Please check for typos and case sensitive issues.".format(target))
    n = scores.shape[0]
    if nb_best > n:
        raise Exception('The number of best models requested is higher than the sample size.')
    constraining = constraining or {}
    try:
        columns = [obj_fn_names.index(fn) for fn in constraining]
    except ValueError:
        raise Exception("One of the names of constraints in Best is not recognised."
                        "Please check for typos and case sensitive issues.")
    mask = condition_mask_tensor(scores, columns, [constraining[fn][1] for fn in constraining],
                                 [constraining[fn][0] for fn in constraining])
    kept = torch.nonzero(mask, as_tuple=False)[:, 0]
    if nb_best > kept.numel():
        raise Exception('The number of best models requested is higher than the restrained sample size.')
    values = scores[kept, t_col]
    # top-k instead of a full sort; re-sorted ascending so that the best set comes last
    top = torch.topk(values, nb_best, largest=True, sorted=True)
    order = torch.flip(top.indices, dims=[0])
    return kept[order]


def latin_hypercube_device(sample_size, bounds, device=None, generator=None):
    """Latin Hypercube sample [sample_size, n_params] built on the device:
    one stratum per row and column, strata visited in the order of a random permutation obtained
    by sorting Philox random keys, jitter uniform within the stratum -- the construction of
    lhs.py:133-167 without the host round trip.  Not the reference's random stream (use
    montecarlo.lhs.latin_hypercube for that); the stratification property is what is kept."""
    torch = _torch()
    bounds_t = torch.as_tensor(np.asarray(bounds, dtype=np.float64), device=device)
    n_params = bounds_t.shape[0]
    keys = torch.rand((n_params, sample_size), dtype=torch.float64, device=device, generator=generator)
    strata = torch.argsort(keys, dim=1).to(torch.float64)
    jitter = torch.rand((n_params, sample_size), dtype=torch.float64, device=device, generator=generator)
    quantiles = (strata + jitter) / sample_size
    lower, width = bounds_t[:, 0:1], (bounds_t[:, 1:2] - bounds_t[:, 0:1])
    return (quantiles * width + lower).t().contiguous()
