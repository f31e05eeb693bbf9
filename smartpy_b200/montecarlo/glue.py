"""GLUE conditioning of an existing sample -- drop-in for ``smartpy/montecarlo/glue.py:30-289``
(Beven & Binley 1992, doi:10.1002/HYP.3360060305): keep the behavioural parameter sets of a
previous LHS run and re-run them (through the batch kernel) on this period.

``GLUE(...)`` reads the sample back from the ``.SMART.lhs`` database like the reference;
``GLUE.from_run(sampling, conditioning, ...)`` conditions the score table a sampling run left on
the device (``smart_condition_rows``: predicate mask + ordered compaction) -- see best.py for how
the two differ (binary64 values instead of their float32 text images).
"""
from .montecarlo import MonteCarlo
from .conditioning import condition_mask, behavioural_rows

_UNKNOWN = ("One of the names of objective functions for conditioning in GLUE is not recognised."
            "Please check for typos and case sensitive issues.")


class GLUE(MonteCarlo):
    """conditioning: dict {objective function name: (kind, (value[, value]))} with kind in
    'equal', 'min', 'max', 'inside', 'outside'."""

    def __init__(self, catchment, root_f, in_format, out_format,
                 conditioning,
                 parallel='seq', save_sim=False, settings_filename=None,
                 decompression_csv=False):
        MonteCarlo.__init__(self, catchment, root_f, in_format, out_format,
                            parallel=parallel, save_sim=save_sim, func='glue', settings_filename=settings_filename)
        self.sampling_run_file = self._sampling_run_file()
        self.sampled_params, self.sampled_obj_fns = self._get_sampled_sets_from_file(
            self.sampling_run_file, self.param_names, self.obj_fn_names, decompression_csv)
        self._read_conditions(conditioning)
        self.behavioural_params = self._get_behavioural_sets(
            self.sampled_params, self.sampled_obj_fns[:, self.objective_fn_indices],
            self.conditions_values, self.conditions_types)
        self._set_sample(self.behavioural_params)

    @classmethod
    def from_run(cls, sampling, conditioning, parallel=None, save_sim=False, settings_filename=None):
        """Condition the scores `sampling.run()` left on the device and set up the re-run of the
        behavioural sets (optionally on another period: settings_filename)."""
        self = cls.__new__(cls)
        self._sibling(sampling, 'glue', parallel, save_sim, settings_filename)
        self.sampling_run_file = None
        self.sampled_params, self.sampled_obj_fns = sampling.sample_params, sampling.results['scores']
        self._read_conditions(conditioning)
        self.behavioural_rows = behavioural_rows(sampling.results['scores'], sampling.obj_fn_names, conditioning)
        self.behavioural_params = sampling.sample_params[self.behavioural_rows.cpu().numpy()]
        self._set_sample(self.behavioural_params)
        return self

    def _read_conditions(self, conditioning):
        try:
            self.objective_fn_indices = [self.obj_fn_names.index(fn) for fn in conditioning]
        except ValueError:
            raise Exception(_UNKNOWN)
        self.conditions_types = [conditioning[fn][0] for fn in conditioning]
        self.conditions_values = [conditioning[fn][1] for fn in conditioning]

    @staticmethod
    def _get_behavioural_sets(params, obj_fns, conditions_val, conditions_typ):
        """Host form (the sample as read from a database file): rows of params passing every
        condition (glue.py:222-289)."""
        if obj_fns.ndim != 2:
            raise Exception('The matrix containing the objective functions is not 2D.')
        if params.ndim != 2:
            raise Exception('The matrix containing the parameters is not 2D.')
        if obj_fns.shape[0] != params.shape[0]:
            raise Exception('The matrices containing objective functions and parameters have different sample sizes.')
        if not obj_fns.shape[1] == len(conditions_val) == len(conditions_typ):
            raise Exception('The objective function matrix and the conditions matrices '
                            'do not have compatible dimensions.')
        return params[condition_mask(obj_fns, conditions_val, conditions_typ)]
