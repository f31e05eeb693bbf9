"""Latin Hypercube sampling of the SMART parameter space -- drop-in for the reference's
``smartpy/montecarlo/lhs.py:31-167`` (McKay et al., 1979, doi:10.2307/1268522)."""
import numpy as np

from .montecarlo import MonteCarlo


def latin_hypercube(sample_size, bounds, rng=None):
    """[sample_size, n_params] float64 sample: column p takes one value in each of the
    sample_size equal-probability strata of U(bounds[p][0], bounds[p][1]), strata visited in a
    random order.

    With rng=None the draws come from numpy's global legacy generator in the same sequence as
    the reference (one rand(N, P) block, then one permutation(N) per column; lhs.py:149-154),
    so ``np.random.seed(s)`` reproduces the reference's sample bit for bit.
    """
    bounds = np.asarray(bounds, dtype=np.float64)
    n_params = bounds.shape[0]
    rand = np.random.rand if rng is None else rng.rand
    permutation = np.random.permutation if rng is None else rng.permutation
    jitter = rand(sample_size, n_params)
    strata = np.empty((sample_size, n_params), dtype=np.float64)
    for p in range(n_params):
        strata[:, p] = permutation(sample_size)
    quantiles = (strata + jitter) / sample_size
    # inverse CDF of the uniform distribution: loc + scale * q
    lower, width = bounds[:, 0], bounds[:, 1] - bounds[:, 0]
    return quantiles * width[None, :] + lower[None, :]


class LHS(MonteCarlo):
    """LHS is the available to perform a sampling in the SMART parameter
    space using a Latin Hypercube sampling `McKay et al. (2000)
    <https:doi.org/10.1080/00401706.2000.10485979>`_.

    Same arguments as the reference: catchment, root_f, in_format, out_format, sample_size,
    parallel='seq' | 'mpi', save_sim=False, settings_filename=None.
    """

    def __init__(self, catchment, root_f, in_format, out_format,
                 sample_size,
                 parallel='seq', save_sim=False, settings_filename=None):
        MonteCarlo.__init__(self, catchment, root_f, in_format, out_format,
                            parallel=parallel, save_sim=save_sim, func='lhs', settings_filename=settings_filename)
        # generate a sample of parameter sets using Latin Hypercube Sampling
        self.lhs_params = self._get_params_from_lh(sample_size)
        self._set_sample(self.lhs_params)

    def _get_params_from_lh(self, sample_size):
        ranges = self.model.parameters.ranges
        bounds = [[ranges[p][0], ranges[p][1]] for p in self.param_names]
        return latin_hypercube(sample_size, bounds)
