"""Best-n conditioning of an existing sample -- drop-in for ``smartpy/montecarlo/best.py:30-287``:
optionally constrain the sample, keep the nb_best sets on a target objective function and
re-run them (through the batch kernel) on this period."""
import numpy as np

from .montecarlo import MonteCarlo, condition_mask


class Best(MonteCarlo):
    def __init__(self, catchment, root_f, in_format, out_format,
                 target, nb_best, constraining=None,
                 parallel='seq', save_sim=False, settings_filename=None,
                 decompression_csv=False):
        MonteCarlo.__init__(self, catchment, root_f, in_format, out_format,
                            parallel=parallel, save_sim=save_sim, func='{}best'.format(nb_best),
                            settings_filename=settings_filename)

        self.sampling_run_file = \
            ''.join([self.model.out_f, catchment, '.SMART.lhs.nc']) if self.out_format == 'netcdf' else \
            ''.join([self.model.out_f, catchment, '.SMART.lhs'])
        self.sampled_params, self.sampled_obj_fns = self._get_sampled_sets_from_file(
            self.sampling_run_file, self.param_names, self.obj_fn_names, decompression_csv)

        try:
            self.target_fn_index = [self.obj_fn_names.index(target)]
        except ValueError:
            raise Exception("The objective function {} for conditioning in Best is not recognised."
                            "Please check for typos and case sensitive issues.".format(target))

        if constraining:
            try:
                self.constraints_indices = [self.obj_fn_names.index(fn) for fn in constraining]
            except ValueError:
                raise Exception("One of the names of constraints in Best is not recognised."
                                "Please check for typos and case sensitive issues.")
            self.constraints_types = [constraining[fn][0] for fn in constraining]
            self.constraints_values = [constraining[fn][1] for fn in constraining]
        else:
            self.constraints_indices, self.constraints_types, self.constraints_values = [], [], []

        self.best_params = self._get_best_sets(
            self.sampled_params, self.sampled_obj_fns[:, self.constraints_indices],
            self.constraints_values, self.constraints_types,
            self.sampled_obj_fns[:, self.target_fn_index], nb_best)
        self._set_sample(self.best_params)

    @staticmethod
    def _get_best_sets(params, constraints_fns, constraints_val, constraints_typ, sort_fn, nb_best):
        if constraints_fns.ndim != 2:
            raise Exception('The matrix containing the constraint functions is not 2D.')
        if params.ndim != 2:
            raise Exception('The matrix containing the parameters is not 2D.')
        if constraints_fns.shape[0] != params.shape[0]:
            raise Exception('The matrices containing constraint functions and parameters have different sample sizes.')
        if not ((constraints_fns.shape[1] == len(constraints_val)) and
                (constraints_fns.shape[1] == len(constraints_typ))):
            raise Exception('The constraint function matrix and the conditions matrices '
                            'do not have compatible dimensions.')
        if sort_fn.shape[0] != params.shape[0]:
            raise Exception('The matrices containing objective functions and parameters have different sample sizes.')
        if nb_best > params.shape[0]:
            raise Exception('The number of best models requested is higher than the sample size.')

        constrained = condition_mask(constraints_fns, constraints_val, constraints_typ)
        kept_params, kept_target = params[constrained, :], sort_fn[constrained, 0]
        if nb_best > kept_params.shape[0]:
            raise Exception('The number of best models requested is higher than the restrained sample size.')
        # ascending sort, best (largest) last -- same order and tie handling as best.py:287
        return kept_params[np.argsort(kept_target)][-nb_best:]
