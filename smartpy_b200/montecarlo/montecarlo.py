"""Monte Carlo base class -- the caller of the hot path.  Same surface as the reference's
``smartpy/montecarlo/montecarlo.py:43-262`` (constructor arguments, ``run(compression)``, the
spotpy-protocol methods ``parameters / simulation / evaluation / objectivefunction / save``, the
sample database format), different engine: ``run()`` does not iterate sample by sample through
spotpy (montecarlo.py:153-154).  The whole sample goes through ONE kernel launch per batch with
the objective functions fused (``BatchEngine.run``) and the database is written in bulk
(``database.SampleDatabase``).  spotpy is therefore not required.  ``parallel='mpi'`` keeps its
meaning of "several processes share the sample", mapped onto ``torch.distributed``: rows are
sharded over the ranks of the process group, the scores all-gathered (NCCL on the GPU box).
"""
from os import sep

import numpy as np

from ..smart import SMART
from ..inout import get_dict_simulation_settings
from ..objfunctions import groundwater_constraint
from .. import distributed as dist_utils
from .conditioning import condition_mask  # noqa: F401  (re-exported: the host form of the selection rules)
from .database import SampleDatabase, database_path, read_sample_database

SCORE_COLUMNS = ('NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE')

# members per kernel launch when the simulated series are kept (bounds device memory:
# n_report x batch x 8 B); scores-only runs go in one launch
_SAVE_SIM_BATCH = 1 << 16


class MonteCarlo(object):
    def __init__(self, catchment, root_f, in_format, out_format,
                 parallel, save_sim, func, settings_filename):
        # where this experiment lives (kept so that Best / GLUE .from_run can set up a sibling)
        self.catchment, self.root_f = catchment, root_f
        self.in_format, self.out_format = in_format, out_format
        self.settings_filename = settings_filename
        self.func = func
        self.parallel = parallel
        self.p = parallel == 'mpi'
        self.save_sim = save_sim

        # the simulation period and catchment description come from the .sttngs file
        settings_file = sep.join([root_f, 'in', catchment, '']) + (settings_filename or catchment + '.sttngs')
        (area, gauged_area, start, end, delta_simu, delta_report,
         warm_up, gw_constraint) = get_dict_simulation_settings(settings_file)
        self.model = SMART(catchment, area, start, end, delta_simu, delta_report, warm_up,
                           in_format, out_format, root_f, gauged_area)

        # objective functions: the seven scores, plus the groundwater constraint when the
        # settings give one (montecarlo.py:66-74)
        self.constraints = {'gw': gw_constraint}
        self.param_names = self.model.parameters.names
        self.obj_fn_names = list(SCORE_COLUMNS) + (['GW'] if gw_constraint else [])

        # the sample: sub-classes register it with _set_sample()
        self.sample_params = None
        self.p_map = None
        self.params = None

        self.db_file = database_path(self.model.out_f, catchment, func, out_format)
        self.database = None
        self.precision = 'f64'
        #: device tensors of the last run: 'scores' [N, n_obj_fns], 'gw' [N]
        self.results = None

        # the observations the scores are computed against go next to the database (montecarlo.py:88)
        self.model.write_output_files(which='observed', parallel=self.p)

    def _set_sample(self, sample):
        """Register the [N, 10] sample (row order = database order)."""
        self.sample_params = np.ascontiguousarray(sample, dtype=np.float64)
        # the reference keeps a {tuple(row): index} dict (lhs.py:117) to place out-of-order MPI
        # results; rows never leave their order here, so the map is built lazily on demand
        self.p_map = _LazyRowMap(self.sample_params)
        self.params = [(name, self.sample_params[:, k]) for k, name in enumerate(self.param_names)]

    # ------------------------------------------------------------------ the spotpy protocol, kept for callers
    def parameters(self):
        try:
            import spotpy
        except ImportError:
            raise Exception('parameters() hands the sample to spotpy, which is not installed; '
                            'use .sample_params or run() instead.')
        return spotpy.parameter.generate([spotpy.parameter.List(name, column) for name, column in self.params])

    def simulation(self, vector):
        discharge, groundwater_component = self.model.simulate(dict(zip(self.param_names, vector)))
        return discharge, [groundwater_component]

    def evaluation(self):
        return self.model.nd_flow, [self.constraints['gw']]

    def objectivefunction(self, simulation, evaluation):
        """Scores of ONE simulated series, computed on the device by the same fused routine as
        the batch path (no CPU scoring code exists in this package)."""
        from ..engine import score_series
        scores = score_series(np.asarray(simulation[0], dtype=np.float64), np.asarray(evaluation[0]))
        if self.constraints['gw']:
            scores.append(groundwater_constraint(evaluation=evaluation[1], simulation=simulation[1]))
        return scores

    def save(self, obj_fns, parameters, simulations, *args, **kwargs):
        """One database row (the per-sample call of the reference, montecarlo.py:211-231)."""
        if self.database is None:
            self.database = self._new_database().open()
        sim = np.asarray(simulations[0], dtype=np.float64)[None, :] if self.save_sim else None
        self.database.write_rows(np.asarray(obj_fns, dtype=np.float64)[None, :],
                                 np.asarray(parameters, dtype=np.float64)[None, :], sim)

    def _new_database(self):
        stamps = self.model.timeseries_report[1:] if self.save_sim else ()
        return SampleDatabase(self.db_file, self.out_format, self.obj_fn_names + self.param_names, stamps)

    # ------------------------------------------------------------------ run
    def run(self, compression=None):
        """Run the simulations for the sample of parameter sets and write the database.

        compression: True gzips the 'csv' database (montecarlo.py:132-177).
        """
        import torch
        n_obj = len(self.obj_fn_names)
        rank, world = dist_utils.rank_world() if self.p else (0, 1)
        if world > 1:
            # ONE sample for the whole group: every rank built its own in its constructor (numpy's
            # global generator is per process), rank 0's is the one that is run and written
            self._set_sample(dist_utils.broadcast_rows(self.sample_params))
        n_rows = self.sample_params.shape[0]
        lo, hi = dist_utils.shard_bounds(n_rows, rank, world)
        engine = self.model.get_engine(report='summary', gw_constraint=self.constraints['gw'],
                                       precision=self.precision)
        writer = rank == 0          # with several ranks only the first one owns the database file
        if writer:
            self.database = self._new_database().open()
        batch = _SAVE_SIM_BATCH if self.save_sim else max(hi - lo, 1)
        blocks, series = [], []
        for first in range(lo, hi, batch):
            rows = self.sample_params[first:min(first + batch, hi)]
            block = torch.empty((rows.shape[0], 9), dtype=torch.float64, device=engine.device)
            res = engine.run(rows, discharge=self.save_sim, scores=True, gw=True, out={'block': block})
            blocks.append(block)
            if self.save_sim:
                sims = res['discharge'].t().contiguous()
                if world == 1:
                    self.database.write_rows(block[:, :n_obj].cpu().numpy(), rows, sims.cpu().numpy())
                else:
                    series.append(sims)
        table = torch.cat(blocks) if blocks else torch.empty((0, 9), dtype=torch.float64, device=engine.device)
        if world > 1:
            table = dist_utils.all_gather_rows(table, n_rows)
        self.results = {'scores': table[:, :n_obj], 'gw': table[:, 8]}
        if writer and not (self.save_sim and world == 1):
            scores_host = table[:, :n_obj].cpu().numpy()
        if world > 1 and self.save_sim:
            # the simulated series of every shard travel to the writer rank by rank, in row order
            mine = torch.cat(series) if series else torch.empty((0, engine.n_report), dtype=torch.float64,
                                                                device=engine.device)
            for r in range(world):
                r_lo, r_hi = dist_utils.shard_bounds(n_rows, r, world)
                sims = dist_utils.send_rows_to_first(mine, r, r_hi - r_lo)
                if writer:
                    self.database.write_rows(scores_host[r_lo:r_hi], self.sample_params[r_lo:r_hi], sims.cpu().numpy())
        elif writer and not self.save_sim:
            self.database.write_rows(scores_host, self.sample_params)
        if writer:
            self.database.close(compression)
        if world > 1:
            dist_utils.barrier()

    # ------------------------------------------------------------------ conditioning on the device
    def select_behavioural(self, conditioning):
        """GLUE's selection applied to the scores of the last run(), on the device: returns
        (row indices tensor, parameter rows [M, 10] float64)."""
        from .conditioning import behavioural_rows
        if self.results is None:
            raise Exception("No results to condition: call run() first.")
        rows = behavioural_rows(self.results['scores'], self.obj_fn_names, conditioning)
        return rows, self.sample_params[rows.cpu().numpy()]

    def select_best(self, target, nb_best, constraining=None):
        """Best's selection applied to the scores of the last run(), on the device (top-k): returns
        (row indices tensor, parameter rows [nb_best, 10] float64), best last."""
        from .conditioning import best_rows
        if self.results is None:
            raise Exception("No results to condition: call run() first.")
        rows = best_rows(self.results['scores'], self.obj_fn_names, target, nb_best, constraining)
        return rows, self.sample_params[rows.cpu().numpy()]

    def _sibling(self, source, func, parallel, save_sim, settings_filename):
        """Set this (bare) instance up as a new experiment on the catchment of `source`, a run that is
        still in memory -- what .from_run of Best / GLUE need before they pick their sample."""
        MonteCarlo.__init__(self, source.catchment, source.root_f, source.in_format, source.out_format,
                            parallel=source.parallel if parallel is None else parallel, save_sim=save_sim,
                            func=func, settings_filename=settings_filename or source.settings_filename)
        if source.results is None:
            raise Exception("No results to condition: call run() on the sampling experiment first.")
        if source.obj_fn_names != self.obj_fn_names and not set(self.obj_fn_names) <= set(source.obj_fn_names):
            raise Exception("The sampling experiment does not hold the objective functions of this one.")

    # ------------------------------------------------------------------ reading a sample database back
    def _get_sampled_sets_from_file(self, file_location, param_names, obj_fn_names, decompression_csv):
        """-> (params float32 [N, 10], obj_fns float32 [N, k]) (montecarlo.py:233-262)."""
        return read_sample_database(file_location, self.out_format, param_names, obj_fn_names,
                                    gzipped=decompression_csv)

    def _sampling_run_file(self):
        """The database of the LHS run that Best / GLUE / Total condition or re-run."""
        return database_path(self.model.out_f, self.catchment, 'lhs', self.out_format)


class _LazyRowMap(object):
    """{tuple(row): index} built on first use (the reference builds it eagerly, lhs.py:117)."""

    def __init__(self, rows):
        self._rows = rows
        self._map = None

    def _build(self):
        if self._map is None:
            self._map = {tuple(self._rows[r, :].tolist()): r for r in range(self._rows.shape[0])}
        return self._map

    def __getitem__(self, key):
        return self._build()[key]

    def __len__(self):
        return self._rows.shape[0]

    def __iter__(self):
        return iter(self._build())

    def __contains__(self, key):
        return key in self._build()
