"""Best-n conditioning of an existing sample -- drop-in for ``smartpy/montecarlo/best.py:30-287``:
optionally constrain the sample, keep the nb_best sets on a target objective function and re-run
them (through the batch kernel) on this period.

Two ways in:

* ``Best(catchment, root_f, in_format, out_format, target, nb_best, constraining, ...)`` -- the
  reference's constructor: the sample comes back from the ``.SMART.lhs`` database file (float32
  text, ``montecarlo.py:233-262``) and is conditioned on the host with the reference's rules;
* ``Best.from_run(sampling, target, nb_best, constraining, ...)`` -- the sampling run is still in
  memory: its float64 score table is conditioned ON THE DEVICE (``smart_best_rows``: predicate
  mask + radix select of the nb_best rows, no full sort, no file round trip) and the selected
  rows go back through the same engine.  Differences with the file path: scores and parameters
  are the run's binary64 values, not their ``'%.6e'``-of-float32 images, so members whose target
  values differ by less than that rounding can be ranked differently, and the re-run uses the
  exact parameters.
"""
import numpy as np

from .montecarlo import MonteCarlo
from .conditioning import condition_mask, best_rows


def _columns(names, wanted, message):
    try:
        return [names.index(name) for name in wanted]
    except ValueError:
        raise Exception(message)


class Best(MonteCarlo):
    def __init__(self, catchment, root_f, in_format, out_format,
                 target, nb_best, constraining=None,
                 parallel='seq', save_sim=False, settings_filename=None,
                 decompression_csv=False):
        MonteCarlo.__init__(self, catchment, root_f, in_format, out_format,
                            parallel=parallel, save_sim=save_sim, func='{}best'.format(nb_best),
                            settings_filename=settings_filename)
        self.sampling_run_file = self._sampling_run_file()
        self.sampled_params, self.sampled_obj_fns = self._get_sampled_sets_from_file(
            self.sampling_run_file, self.param_names, self.obj_fn_names, decompression_csv)
        self._read_conditions(target, constraining)
        self.best_params = self._get_best_sets(
            self.sampled_params, self.sampled_obj_fns[:, self.constraints_indices],
            self.constraints_values, self.constraints_types,
            self.sampled_obj_fns[:, self.target_fn_index], nb_best)
        self._set_sample(self.best_params)

    @classmethod
    def from_run(cls, sampling, target, nb_best, constraining=None,
                 parallel=None, save_sim=False, settings_filename=None):
        """Condition the scores `sampling.run()` left on the device and set up the re-run of the
        nb_best sets (optionally on another period: settings_filename)."""
        self = cls.__new__(cls)
        self._sibling(sampling, '{}best'.format(nb_best), parallel, save_sim, settings_filename)
        self.sampling_run_file = None
        self.sampled_params, self.sampled_obj_fns = sampling.sample_params, sampling.results['scores']
        self._read_conditions(target, constraining)
        self.best_rows = best_rows(sampling.results['scores'], sampling.obj_fn_names, target, nb_best, constraining)
        self.best_params = sampling.sample_params[self.best_rows.cpu().numpy()]
        self._set_sample(self.best_params)
        return self

    def _read_conditions(self, target, constraining):
        self.target_fn_index = _columns(
            self.obj_fn_names, [target],
            "The objective function {} for conditioning in Best is not recognised."
            "Please check for typos and case sensitive issues.".format(target))
        constraining = constraining or {}
        self.constraints_indices = _columns(
            self.obj_fn_names, constraining,
            "One of the names of constraints in Best is not recognised."
            "Please check for typos and case sensitive issues.")
        self.constraints_types = [constraining[fn][0] for fn in constraining]
        self.constraints_values = [constraining[fn][1] for fn in constraining]

    @staticmethod
    def _get_best_sets(params, constraints_fns, constraints_val, constraints_typ, sort_fn, nb_best):
        """Host form (the sample as read from a database file): rows of params passing the
        constraints, the nb_best largest on sort_fn, ascending with the best last (best.py:221-287)."""
        shapes_ok = (
            (constraints_fns.ndim == 2, 'The matrix containing the constraint functions is not 2D.'),
            (params.ndim == 2, 'The matrix containing the parameters is not 2D.'),
            (constraints_fns.shape[0] == params.shape[0],
             'The matrices containing constraint functions and parameters have different sample sizes.'),
            (constraints_fns.ndim == 2 and constraints_fns.shape[1] == len(constraints_val) == len(constraints_typ),
             'The constraint function matrix and the conditions matrices do not have compatible dimensions.'),
            (sort_fn.shape[0] == params.shape[0],
             'The matrices containing objective functions and parameters have different sample sizes.'),
            (nb_best <= params.shape[0], 'The number of best models requested is higher than the sample size.'),
        )
        for ok, message in shapes_ok:
            if not ok:
                raise Exception(message)
        kept = condition_mask(constraints_fns, constraints_val, constraints_typ)
        if nb_best > int(kept.sum()):
            raise Exception('The number of best models requested is higher than the restrained sample size.')
        # numpy's default argsort, as best.py:287: ascending, best (largest) last
        ranking = np.argsort(sort_fn[kept, 0])
        return params[kept][ranking][-nb_best:]
