"""Re-run of a whole existing sample on another period, no conditioning -- drop-in for
``smartpy/montecarlo/total.py:27-144``.  ``Total.from_run`` takes the sample of a run that is
still in memory (its binary64 rows) instead of the database file."""
from .montecarlo import MonteCarlo


class Total(MonteCarlo):
    def __init__(self, catchment, root_f, in_format, out_format,
                 parallel='seq', save_sim=False, settings_filename=None, decompression_csv=False):
        MonteCarlo.__init__(self, catchment, root_f, in_format, out_format,
                            parallel=parallel, save_sim=save_sim, func='total', settings_filename=settings_filename)
        self.sampling_run_file = self._sampling_run_file()
        self.sampled_params, self.sampled_obj_fns = self._get_sampled_sets_from_file(
            self.sampling_run_file, self.param_names, self.obj_fn_names, decompression_csv)
        self._set_sample(self.sampled_params)

    @classmethod
    def from_run(cls, sampling, parallel=None, save_sim=False, settings_filename=None):
        self = cls.__new__(cls)
        self._sibling(sampling, 'total', parallel, save_sim, settings_filename)
        self.sampling_run_file = None
        self.sampled_params, self.sampled_obj_fns = sampling.sample_params, sampling.results['scores']
        self._set_sample(self.sampled_params)
        return self
