"""Monte Carlo base class -- same surface as the reference's
``smartpy/montecarlo/montecarlo.py:43-262`` (constructor arguments, ``run(compression)``, the
spotpy-protocol methods ``parameters / simulation / evaluation / objectivefunction / save``,
the sample database formats), but ``run()`` does not iterate sample by sample through
spotpy (montecarlo.py:153-154): the whole sample goes through ONE kernel launch per batch
with the objective functions fused (``BatchEngine.run``), and the database is written in
bulk.  spotpy is therefore not required; ``parallel='mpi'`` is accepted for signature
compatibility and mapped onto ``torch.distributed`` when a process group is initialised
(rows are sharded over ranks, scores all-gathered over NCCL).
"""
from csv import DictReader
import gzip
from io import open
from os import sep, remove, rename
import shutil

import numpy as np

try:
    from netCDF4 import Dataset
except ImportError:
    Dataset = None

from ..smart import SMART
from ..inout import get_dict_simulation_settings
from ..objfunctions import groundwater_constraint
from ..version import __version__
from .. import distributed as dist_utils

_NETCDF_OUT = ("The use of 'netcdf' as the output file format requires the package 'netCDF4', "
               "please install it and retry, or choose another file format.")

# members per kernel launch when the simulated series are kept (bounds device memory:
# n_report x batch x 8 B); scores-only runs go in one launch
_SAVE_SIM_BATCH = 1 << 16


def condition_mask(obj_fns, conditions_val, conditions_typ):
    """Boolean mask over the rows of obj_fns[N, k] for k (kind, values) conditions -- the
    selection rules shared by GLUE (glue.py:246-286) and Best (best.py:243-277)."""
    mask = np.ones((obj_fns.shape[0],), dtype=bool)
    for column, values, kind in zip(obj_fns.T, conditions_val, conditions_typ):
        if kind in ('equal', 'min', 'max'):
            if len(values) != 1:
                raise Exception("The tuple for \"{}\" condition does not contain one and only one "
                                "element.".format(kind))
            if kind == 'equal':
                selection = column == values[0]
            elif kind == 'min':
                selection = column >= values[0]
            else:
                selection = column <= values[0]
        elif kind in ('inside', 'outside'):
            if len(values) != 2:
                raise Exception("The tuple for \"{}\" condition does not contain two and only two "
                                "elements.".format(kind))
            if not values[1] > values[0]:
                raise Exception("The two elements of the tuple for \"{}\" are inconsistent.".format(kind))
            if kind == 'inside':
                selection = (column >= values[0]) & (column <= values[1])
            else:
                selection = (column <= values[0]) & (column >= values[1])   # as written at glue.py:278
        else:
            raise Exception("The type of threshold \"{}\" is not in the database.".format(kind))
        mask &= selection
    return mask


class MonteCarlo(object):
    def __init__(self, catchment, root_f, in_format, out_format,
                 parallel, save_sim, func, settings_filename):
        in_f = sep.join([root_f, 'in', catchment, sep])

        # collect the simulation information from the .sttngs file
        settings_file = ''.join([in_f, settings_filename if settings_filename else catchment + '.sttngs'])
        c_area, g_area, start, end, delta_simu, delta_report, warm_up, gw_constraint = \
            get_dict_simulation_settings(settings_file)

        # generate an instance of the SMART model class
        self.model = SMART(catchment, c_area, start, end, delta_simu, delta_report, warm_up,
                           in_format, out_format, root_f,
                           g_area)

        # set the technical aspects of the simulation
        self.parallel = parallel
        self.p = True if parallel == 'mpi' else False
        self.save_sim = save_sim

        # possible additional constraint for SMART on the portion of base flow in runoff
        self.constraints = {'gw': gw_constraint}

        self.param_names = self.model.parameters.names
        self.obj_fn_names = \
            ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW'] \
            if self.constraints['gw'] else \
            ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE']

        # the sample: sub-classes set sample_params [N, 10]; p_map / params kept for compatibility
        self.sample_params = None
        self.p_map = None
        self.params = None

        # database
        self.out_format = out_format
        self.db_file = \
            self.model.out_f + '{}.SMART.{}.nc'.format(catchment, func) if self.out_format == 'netcdf' else \
            self.model.out_f + '{}.SMART.{}'.format(catchment, func)
        self.database = None
        self.precision = 'f64'
        #: device tensors of the last run: 'scores' [N, n_obj_fns], 'gw' [N] (and 'discharge' if save_sim)
        self.results = None

        # write out the observed discharge data used for the objective functions
        self.model.write_output_files(which='observed', parallel=self.p)

    def _set_sample(self, sample):
        """Register the [N, 10] sample (row order = database order)."""
        self.sample_params = np.ascontiguousarray(sample, dtype=np.float64)
        # the reference keeps a {tuple(row): index} dict (lhs.py:117) to place out-of-order MPI
        # results; rows never leave their order here, so the map is built lazily on demand
        self.p_map = _LazyRowMap(self.sample_params)
        self.params = [(name, self.sample_params[:, k]) for k, name in enumerate(self.param_names)]

    # ------------------------------------------------------------------ database
    def _init_db(self):
        n_samples = self.sample_params.shape[0]
        if self.out_format == 'netcdf':
            if not Dataset:
                raise Exception(_NETCDF_OUT)
            self.database = Dataset(self.db_file, 'w', format='NETCDF4', parallel=self.p)
            self.database.description = "Monte Carlo Simulation outputs with SMARTpy v{}.".format(__version__)
            self.database.createDimension('NbSamples', n_samples)
            self.database.createDimension('NbParameters', len(self.model.parameters.names))
            self.database.createDimension('NbObjFunctions', len(self.obj_fn_names))
            params = self.database.createVariable('Parameters', np.float32, ('NbSamples', 'NbParameters'))
            params.units = ', '.join(self.model.parameters.names)
            objfns = self.database.createVariable('ObjFunctions', np.float32, ('NbSamples', 'NbObjFunctions'))
            objfns.units = ', '.join(self.obj_fn_names)
            if self.save_sim:
                stamps = self.model.timeseries_report[1:]
                self.database.createDimension('DateTime', len(stamps))
                times = self.database.createVariable('DateTime', np.float64, ('DateTime',))
                times.units = "seconds since 1970-01-01 00:00:00.0"
                simu = self.database.createVariable('Simulations', np.float32, ('NbSamples', 'DateTime'))
                simu.units = "Discharge in m3/s"
                timestamps = (np.array(stamps, dtype='datetime64[s]') - np.datetime64('1970-01-01T00:00:00')) / \
                    np.timedelta64(1, 's')
                self.database.variables['DateTime'][0:len(stamps)] = timestamps
        else:
            self.database = open(self.db_file, 'w', newline='', encoding='utf8')
            simu_steps = [dt.strftime('%Y-%m-%d %H:%M:%S') for dt in self.model.timeseries_report[1:]] \
                if self.save_sim else []
            self.database.write(','.join(self.obj_fn_names + self.param_names + simu_steps) + '\n')

    def _save_block(self, first_row, obj_fns, parameters, simulations):
        """Bulk form of save(): rows first_row .. first_row + n of the database."""
        n = obj_fns.shape[0]
        if self.out_format == 'netcdf':
            self.database.variables['Parameters'][first_row:first_row + n, :] = parameters
            self.database.variables['ObjFunctions'][first_row:first_row + n, :] = obj_fns
            if self.save_sim:
                self.database.variables['Simulations'][first_row:first_row + n, :] = simulations
        else:
            block = [obj_fns, parameters] + ([simulations] if self.save_sim else [])
            table = np.concatenate(block, axis=1).astype(np.float32)   # '%.6e' of float32, montecarlo.py:226-231
            np.savetxt(self.database, table, fmt='%.6e', delimiter=',')

    # ------------------------------------------------------------------ the spotpy protocol, kept for callers
    def parameters(self):
        try:
            import spotpy
        except ImportError:
            raise Exception('parameters() hands the sample to spotpy, which is not installed; '
                            'use .sample_params or run() instead.')
        return spotpy.parameter.generate([spotpy.parameter.List(name, column) for name, column in self.params])

    def simulation(self, vector):
        discharge, groundwater_component = self.model.simulate(
            {name: vector[k] for k, name in enumerate(self.param_names)})
        return (
            discharge, [groundwater_component]
        )

    def evaluation(self):
        return (
            self.model.nd_flow, [self.constraints['gw']]
        )

    def objectivefunction(self, simulation, evaluation):
        """Scores of ONE simulated series, computed on the device by the same fused routine as
        the batch path (no CPU scoring code exists in this package)."""
        from ..engine import score_series
        scores = score_series(np.asarray(simulation[0], dtype=np.float64), np.asarray(evaluation[0]))
        if self.constraints['gw']:
            return scores + [groundwater_constraint(evaluation=evaluation[1], simulation=simulation[1])]
        return scores

    def save(self, obj_fns, parameters, simulations, *args, **kwargs):
        parameters = np.asarray(parameters, dtype=np.float64)
        index = self.p_map[tuple(parameters.tolist())] if self.out_format == 'netcdf' else 0
        sim = np.asarray(simulations[0], dtype=np.float64)[None, :] if self.save_sim else None
        self._save_block(index, np.asarray(obj_fns, dtype=np.float64)[None, :], parameters[None, :], sim)

    # ------------------------------------------------------------------ run
    def run(self, compression=None):
        """Run the simulations for the sample of parameter sets.

        compression: for 'csv' a bool (gzip the database); for 'netcdf' a bool or the zlib
        complevel 1-9 (True = 6); None = no compression (montecarlo.py:132-177).
        """
        n_obj = len(self.obj_fn_names)
        rank, world = dist_utils.rank_world() if self.p else (0, 1)
        writer = rank == 0          # with several ranks only the first one owns the database file
        if writer:
            self._init_db()
        lo, hi = dist_utils.shard_bounds(self.sample_params.shape[0], rank, world)
        engine = self.model.get_engine(report='summary', gw_constraint=self.constraints['gw'],
                                       precision=self.precision)
        batch = _SAVE_SIM_BATCH if self.save_sim else max(hi - lo, 1)
        scores_parts, gw_parts = [], []
        for first in range(lo, hi, batch):
            rows = self.sample_params[first:min(first + batch, hi)]
            res = engine.run(rows, discharge=self.save_sim, scores=True, gw=True)
            scores_parts.append(res['scores'])
            gw_parts.append(res['gw'])
            if self.save_sim and world == 1:
                self._save_block(first, res['scores'][:, :n_obj].cpu().numpy(), rows,
                                 res['discharge'].t().cpu().numpy())
        import torch
        scores = torch.cat(scores_parts) if scores_parts else torch.empty((0, 8), dtype=torch.float64)
        gw = torch.cat(gw_parts) if gw_parts else torch.empty((0,), dtype=torch.float64)
        if world > 1:
            scores, gw = dist_utils.all_gather_rows(scores, gw, self.sample_params.shape[0])
        self.results = {'scores': scores[:, :n_obj], 'gw': gw}
        if writer:
            if not self.save_sim or world > 1:
                self._save_block(0, scores[:, :n_obj].cpu().numpy(), self.sample_params, None)
            self.database.close()
        if world > 1:
            dist_utils.barrier()
        if not writer:
            return

        # if compression argument given, the file created will be compressed
        if self.out_format == 'netcdf':
            if compression is True:
                compression = 6
            if not isinstance(compression, bool) and isinstance(compression, (int, float)):
                with Dataset(self.db_file, 'r') as src, Dataset(self.db_file.replace('.nc', '_.nc'), 'w') as dst:
                    dst.description = src.description
                    for name, dimension in src.dimensions.items():
                        dst.createDimension(name, len(dimension))
                    for name, variable in src.variables.items():
                        v = dst.createVariable(name, variable.datatype, variable.dimensions,
                                               zlib=True, complevel=compression)
                        v.units = src.variables[name].units
                        dst.variables[name][:] = src.variables[name][:]
                remove(self.db_file)
                rename(self.db_file.replace('.nc', '_.nc'), self.db_file)
        elif self.out_format == 'csv':
            if compression is True:
                with open(self.db_file, 'rb') as f_in:
                    with gzip.open(self.db_file + '.gz', 'wb') as f_out:
                        shutil.copyfileobj(f_in, f_out)
                remove(self.db_file)

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

    # ------------------------------------------------------------------ reading a sample database back
    def _get_sampled_sets_from_file(self, file_location, param_names, obj_fn_names, decompression_csv):
        """-> (params float32 [N, 10], obj_fns float32 [N, k]) (montecarlo.py:233-262)."""
        if self.out_format == 'netcdf':
            if not Dataset:
                raise Exception(_NETCDF_OUT)
            with Dataset(file_location, 'r') as handle:
                params = np.asarray(handle.variables['Parameters'][:, :])
                obj_fns = np.asarray(handle.variables['ObjFunctions'][:, :])
            return np.array(params, dtype=np.float32), np.array(obj_fns, dtype=np.float32)
        opener = (lambda: gzip.open(file_location + '.gz', 'rt', encoding='utf8')) if decompression_csv else \
            (lambda: open(file_location, 'r', encoding='utf8'))
        obj_fns, params = list(), list()
        with opener() as handle:
            for row in DictReader(handle):
                obj_fns.append([row[obj_fn] for obj_fn in obj_fn_names])
                params.append([row[param] for param in param_names])
        return np.array(params, dtype=np.float32), np.array(obj_fns, dtype=np.float32)


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
