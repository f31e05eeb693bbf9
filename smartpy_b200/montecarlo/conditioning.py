"""Conditioning of a scored sample on the device.

The reference conditions a sample after a round trip through a float32 text/NetCDF file
(``montecarlo.py:233-262``) with numpy masks and a full ``argsort`` (``glue.py:246-289``,
``best.py:243-287``).  With the scores already in HBM (``MonteCarlo.results``) the same
selections are two calls into the CUDA library: ``smart_condition_rows`` (predicate mask +
ordered compaction) and ``smart_best_rows`` (radix select of the k best + sort of those k only),
``include/smart_b200.h``.  They follow the reference's rules, including its order (ascending,
best last), numpy's NaN-sorts-last rule and the literal 'outside' rule.  ``Best.from_run`` /
``GLUE.from_run`` use them on the score table of a run that is still in memory; the file-based
constructors (sample read back as float32 from the database) use the numpy form of the same
rules, ``condition_mask``.
"""
import numpy as np

from .. import _native


def _torch():
    import torch
    return torch


def parse_conditions(conditions_val, conditions_typ):
    """[(kind, lo, hi)] from the reference's (values tuple, kind) pairs, with the reference's checks and
    messages (glue.py:246-286, best.py:243-277): 'equal' / 'min' / 'max' take one value, 'inside' /
    'outside' two increasing ones."""
    parsed = []
    for values, kind in zip(conditions_val, conditions_typ):
        if kind in ('equal', 'min', 'max'):
            if len(values) != 1:
                raise Exception("The tuple for \"{}\" condition does not contain one and only one "
                                "element.".format(kind))
            parsed.append((kind, float(values[0]), 0.0))
        elif kind in ('inside', 'outside'):
            if len(values) != 2:
                raise Exception("The tuple for \"{}\" condition does not contain two and only two "
                                "elements.".format(kind))
            if not values[1] > values[0]:
                raise Exception("The two elements of the tuple for \"{}\" are inconsistent.".format(kind))
            parsed.append((kind, float(values[0]), float(values[1])))
        else:
            raise Exception("The type of threshold \"{}\" is not in the database.".format(kind))
    return parsed


_HOST_RULES = {
    'equal': lambda x, lo, hi: x == lo,
    'min': lambda x, lo, hi: x >= lo,
    'max': lambda x, lo, hi: x <= lo,
    'inside': lambda x, lo, hi: (x >= lo) & (x <= hi),
    'outside': lambda x, lo, hi: (x <= lo) & (x >= hi),      # the reference's literal rule (glue.py:278)
}


def condition_mask(obj_fns, conditions_val, conditions_typ):
    """Host form (numpy) of the selection rules, for the sample read back from a database file:
    boolean mask over the rows of obj_fns[N, k], one (values, kind) condition per column (thresholds
    enter the comparisons as Python floats, as in the reference)."""
    mask = np.ones((obj_fns.shape[0],), dtype=bool)
    for column, (kind, lo, hi) in zip(obj_fns.T, parse_conditions(conditions_val, conditions_typ)):
        mask &= _HOST_RULES[kind](column, lo, hi)
    return mask


def _conditions(columns, conditions_val, conditions_typ):
    """ctypes array of smart_condition for the device kernels."""
    if len(columns) > _native.MAX_CONDITIONS:
        raise Exception("At most {} conditions can be combined.".format(_native.MAX_CONDITIONS))
    conds = (_native.Condition * max(1, len(columns)))()
    for i, (col, (kind, lo, hi)) in enumerate(zip(columns, parse_conditions(conditions_val, conditions_typ))):
        conds[i] = _native.Condition(int(col), _native.COND_KINDS[kind], lo, hi)
    return conds


def _score_table(scores):
    """(tensor, rows, leading dimension) of a float64 CUDA score table whose rows are unit-stride
    (a column slice of the kernel's [N, 8] output is fine)."""
    torch = _torch()
    if not (torch.is_tensor(scores) and scores.is_cuda):
        raise RuntimeError("device conditioning needs the CUDA score table of a batch run "
                           "(smartpy_b200 has no CPU path; use GLUE/Best on the sample database instead)")
    if scores.dim() != 2 or scores.dtype != torch.float64:
        raise ValueError("scores must be a [N, k] float64 tensor")
    if scores.shape[0] > 0 and (scores.stride(1) != 1 or scores.stride(0) < scores.shape[1]):
        scores = scores.contiguous()
    ld = scores.stride(0) if scores.shape[0] > 1 else max(scores.shape[1], 1)
    return scores, scores.shape[0], ld


def _rows_where(scores, columns, conditions_val, conditions_typ):
    torch = _torch()
    conds = _conditions(columns, conditions_val, conditions_typ)
    scores, n, ld = _score_table(scores)
    lib = _native.load()
    with torch.cuda.device(scores.device):
        rows = torch.empty((n,), dtype=torch.int64, device=scores.device)
        count = torch.zeros((1,), dtype=torch.int64, device=scores.device)
        work = torch.empty((max(1, lib.smart_condition_workspace_bytes(n, 0)),), dtype=torch.uint8,
                           device=scores.device)
        _native.check(lib.smart_condition_rows(
            scores.data_ptr(), n, ld, conds, len(columns), rows.data_ptr(), count.data_ptr(),
            work.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return rows[:int(count.item())]


def behavioural_rows(scores, obj_fn_names, conditioning):
    """Row indices (ascending) of the behavioural sets -- GLUE (glue.py:222-289)."""
    try:
        columns = [obj_fn_names.index(fn) for fn in conditioning]
    except ValueError:
        raise Exception("One of the names of objective functions for conditioning in GLUE is not recognised."
                        "Please check for typos and case sensitive issues.")
    return _rows_where(scores, columns, [conditioning[fn][1] for fn in conditioning],
                       [conditioning[fn][0] for fn in conditioning])


def best_rows(scores, obj_fn_names, target, nb_best, constraining=None):
    """Row indices of the nb_best sets on `target` after the optional constraints, in the
    reference's order: ascending on the target, best last (best.py:221-287)."""
    torch = _torch()
    try:
        t_col = obj_fn_names.index(target)
    except ValueError:
        raise Exception("The objective function {} for conditioning in Best is not recognised."
                        "Please check for typos and case sensitive issues.".format(target))
    if nb_best > scores.shape[0]:
        raise Exception('The number of best models requested is higher than the sample size.')
    constraining = constraining or {}
    try:
        columns = [obj_fn_names.index(fn) for fn in constraining]
    except ValueError:
        raise Exception("One of the names of constraints in Best is not recognised."
                        "Please check for typos and case sensitive issues.")
    conds = _conditions(columns, [constraining[fn][1] for fn in constraining],
                        [constraining[fn][0] for fn in constraining])
    scores, n, ld = _score_table(scores)
    nb_best = int(nb_best)
    if nb_best < 1:
        return torch.empty((0,), dtype=torch.int64, device=scores.device)
    lib = _native.load()
    with torch.cuda.device(scores.device):
        rows = torch.empty((nb_best,), dtype=torch.int64, device=scores.device)
        kept = torch.zeros((1,), dtype=torch.int64, device=scores.device)
        work = torch.empty((lib.smart_condition_workspace_bytes(n, nb_best),), dtype=torch.uint8,
                           device=scores.device)
        _native.check(lib.smart_best_rows(
            scores.data_ptr(), n, ld, t_col, conds, len(columns), nb_best, rows.data_ptr(), kept.data_ptr(),
            work.data_ptr(), torch.cuda.current_stream().cuda_stream))
        if nb_best > int(kept.item()):
            raise Exception('The number of best models requested is higher than the restrained sample size.')
    return rows


def latin_hypercube_device(sample_size, bounds, device=None, seed=0, row_first=0, n_rows=None):
    """Rows [row_first, row_first + n_rows) of a Latin Hypercube sample [sample_size, n_params],
    generated on the device by ``smart_lhs_rows``: one stratum per row and
    column as in lhs.py:133-167, the strata order given by a keyed bijection instead of a host
    permutation, so each rank of a sharded run generates only its own rows and the union is one
    stratified sample.  Not the reference's random stream (montecarlo.lhs.latin_hypercube keeps
    that one); the stratification property is what is kept.  Returns a CUDA float64 tensor."""
    torch = _torch()
    bounds = np.ascontiguousarray(bounds, dtype=np.float64)
    if bounds.ndim != 2 or bounds.shape[1] != 2:
        raise ValueError("bounds must be [n_params, 2]")
    n_rows = int(sample_size) - int(row_first) if n_rows is None else int(n_rows)
    device = torch.device('cuda' if device is None else device)
    if device.type != 'cuda':
        raise RuntimeError("latin_hypercube_device runs on a CUDA device (smartpy_b200 has no CPU path; "
                           "montecarlo.lhs.latin_hypercube is the host sampler)")
    lib = _native.load()
    with torch.cuda.device(device):
        out = torch.empty((max(n_rows, 0), bounds.shape[0]), dtype=torch.float64, device=device)
        _native.check(lib.smart_lhs_rows(int(seed) & 0xFFFFFFFFFFFFFFFF, int(sample_size), int(row_first), n_rows,
                                         bounds.shape[0], bounds.ctypes.data, out.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
    return out
