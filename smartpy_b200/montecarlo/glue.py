"""GLUE conditioning of an existing sample -- drop-in for ``smartpy/montecarlo/glue.py:30-289``
(Beven & Binley 1992, doi:10.1002/HYP.3360060305): keep the behavioural parameter sets of a
previous LHS run and re-run them (through the batch kernel) on this period."""
from .montecarlo import MonteCarlo, condition_mask


class GLUE(MonteCarlo):
    """conditioning: dict {objective function name: (kind, (value[, value]))} with kind in
    'equal', 'min', 'max', 'inside', 'outside'."""

    def __init__(self, catchment, root_f, in_format, out_format,
                 conditioning,
                 parallel='seq', save_sim=False, settings_filename=None,
                 decompression_csv=False):
        MonteCarlo.__init__(self, catchment, root_f, in_format, out_format,
                            parallel=parallel, save_sim=save_sim, func='glue', settings_filename=settings_filename)

        # collect the sampling sets from the Monte Carlo simulation (LHS sampling)
        self.sampling_run_file = \
            ''.join([self.model.out_f, catchment, '.SMART.lhs.nc']) if self.out_format == 'netcdf' else \
            ''.join([self.model.out_f, catchment, '.SMART.lhs'])
        self.sampled_params, self.sampled_obj_fns = self._get_sampled_sets_from_file(
            self.sampling_run_file, self.param_names, self.obj_fn_names, decompression_csv)

        try:
            self.objective_fn_indices = [self.obj_fn_names.index(fn) for fn in conditioning]
        except ValueError:
            raise Exception("One of the names of objective functions for conditioning in GLUE is not recognised."
                            "Please check for typos and case sensitive issues.")
        self.conditions_types = [conditioning[fn][0] for fn in conditioning]
        self.conditions_values = [conditioning[fn][1] for fn in conditioning]

        # extract behavioural sets from sampling sets
        self.behavioural_params = self._get_behavioural_sets(
            self.sampled_params, self.sampled_obj_fns[:, self.objective_fn_indices],
            self.conditions_values, self.conditions_types)
        self._set_sample(self.behavioural_params)

    @staticmethod
    def _get_behavioural_sets(params, obj_fns, conditions_val, conditions_typ):
        if obj_fns.ndim != 2:
            raise Exception('The matrix containing the objective functions is not 2D.')
        if params.ndim != 2:
            raise Exception('The matrix containing the parameters is not 2D.')
        if obj_fns.shape[0] != params.shape[0]:
            raise Exception('The matrices containing objective functions and parameters have different sample sizes.')
        if not ((obj_fns.shape[1] == len(conditions_val)) and (obj_fns.shape[1] == len(conditions_typ))):
            raise Exception('The objective function matrix and the conditions matrices '
                            'do not have compatible dimensions.')
        return params[condition_mask(obj_fns, conditions_val, conditions_typ), :]
