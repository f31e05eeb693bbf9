"""The ``SMART`` model object -- drop-in for the reference's ``smartpy/smart.py:29-277``
(same constructor, attributes, ``simulate`` / ``write_output_files`` /
``get_simulation_array`` / ``get_evaluation_array``), with the simulation itself routed to the
CUDA kernel (one member per ``simulate`` call; ``simulate_batch`` for many).

``simulate`` replaces ``structure.run`` (smartpy/structure.py:30-146) end to end: the initial
conditions guess, the warm-up run and the main run all happen inside one kernel launch.
There is no CPU path: without the built library or without a GPU it raises.
"""
from os import makedirs, sep

import numpy as np

from .timeframe import TimeFrame
from .parameters import Parameters
from .inout import get_forcing_series_simu, get_discharge_series, write_flow_file_from_nds
from .engine import BatchEngine, warm_up_length


class SMART(object):
    """SMART is the core object to set up and use to run an experiment."""

    def __init__(self, catchment, catchment_area_m2, start, end,
                 time_delta_simu, time_delta_save, warm_up_days,
                 in_format, out_format, root,
                 gauged_area_m2=None):
        # general information
        self.catchment = catchment
        self.area = catchment_area_m2
        # directory information
        self.in_fmt = in_format
        self.out_fmt = out_format
        self.root_f = root
        self.in_f = sep.join([self.root_f, 'in', self.catchment, sep])
        self.out_f = sep.join([self.root_f, 'out', self.catchment, sep])
        makedirs(self.out_f, exist_ok=True)      # (exist_ok: several ranks may get here together)
        # temporal information
        self.start = start
        self.end = end
        self.delta_simu = time_delta_simu
        self.delta_save = time_delta_save
        self.timeframe = TimeFrame(self.start, self.end, self.delta_simu, self.delta_save)
        self.timeseries = self.timeframe.get_series_simu()
        self.timeseries_report = self.timeframe.get_series_save()
        self.warm_up = warm_up_days
        # physical information as numpy arrays (on the simulation / reporting grids)
        ext = '.nc' if self.in_fmt == 'netcdf' else ''
        base = ''.join([self.in_f, self.catchment])
        self.nd_rain = get_forcing_series_simu(base + '.rain' + ext, self.in_fmt, 'rain',
                                               self.timeseries[1], self.timeseries[-1], self.delta_simu)
        self.nd_peva = get_forcing_series_simu(base + '.peva' + ext, self.in_fmt, 'peva',
                                               self.timeseries[1], self.timeseries[-1], self.delta_simu)
        self.nd_flow = get_discharge_series(base + '.flow' + ext, self.in_fmt,
                                            self.timeseries_report[1], self.timeseries_report[-1],
                                            catchment_area_m2, gauged_area_m2) if gauged_area_m2 else None
        self._dicts = {}
        # optional extra information for setting up initial levels in reservoirs
        self.extra = None
        # parameters
        #: Return the set of SMART model parameters as a `parameters.Parameters` object.
        self.parameters = Parameters()
        # model outputs
        self.outputs = None
        self.nd_discharge = None
        self.gw_contribution = None
        self._engines = {}

    # the reference keeps datetime-keyed dicts of the three series; build them on demand
    def _series_dict(self, name, stamps, values):
        if name not in self._dicts:
            self._dicts[name] = None if values is None else dict(zip(stamps, values))
        return self._dicts[name]

    @property
    def rain(self):
        return self._series_dict('rain', self.timeseries[1:], self.nd_rain)

    @property
    def peva(self):
        return self._series_dict('peva', self.timeseries[1:], self.nd_peva)

    @property
    def flow(self):
        return self._series_dict('flow', self.timeseries_report[1:], self.nd_flow)

    # ------------------------------------------------------------------ engine
    def get_engine(self, report='summary', gw_constraint=None, precision='f64', with_obs=True):
        """The device-resident batch engine for this catchment and reporting mode (cached)."""
        extra_key = None
        if self.extra:
            extra_key = (self.extra['aar'], self.extra['r-o_ratio'], tuple(self.extra['r-o_split']))
        key = (report, gw_constraint, precision, with_obs and self.nd_flow is not None, extra_key, self.warm_up)
        if key not in self._engines:
            simu_length = len(self.timeseries) - 1                                       # structure.py:73
            report_gap = (len(self.timeseries) - 1) // (len(self.timeseries_report) - 1)   # structure.py:75
            delta_sec = self.delta_simu.total_seconds()
            warm = warm_up_length(self.warm_up, delta_sec) if self.warm_up != 0 else 0   # structure.py:87-88
            if warm > simu_length:                                                       # structure.py:90-95
                raise Exception(
                    "The warm-up duration (i.e. {} days) cannot exceed the length of the simulation period "
                    "because the beginning of the simulation period is used as made-up warm-up data for the "
                    "sake of model states initialisation. Please specify another warm-up duration to comply "
                    "with this requirement, or consider using actual warm-up data at the beginning of the "
                    "simulation period and set the warm-up period to 0.".format(self.warm_up))
            self._engines[key] = BatchEngine(
                self.nd_rain[:simu_length], self.nd_peva[:simu_length], self.area, delta_sec, report_gap,
                obs=self.nd_flow if (with_obs and self.nd_flow is not None) else None,
                extra=self.extra, warm_up_steps=warm, report=report, gw_constraint=gw_constraint,
                precision=precision)
        return self._engines[key]

    # ------------------------------------------------------------------ simulation
    def simulate(self, param, report='summary'):
        """Run model simulation over period configuration at instantiation.

        param: dict of the ten SMART parameters (``parameters.names``);
        report: 'summary' (mean of the simulation steps in each reporting step) or 'raw'
        (last simulation step of each reporting step).  Returns ``(discharge, gw)``.
        """
        nd_parameters = np.array([param[name] for name in self.parameters.names], dtype=np.float64)
        engine = self.get_engine(report=report, with_obs=False)
        res = engine.run(nd_parameters[None, :], discharge=True, scores=False, gw=True)
        discharge = res['discharge'][:, 0].cpu().numpy()
        gw = float(res['gw'][0].item())
        self.outputs = (discharge, gw)
        self.nd_discharge = self.outputs[0]
        self.gw_contribution = self.outputs[1]
        return self.outputs

    def simulate_batch(self, params, report='summary', discharge=True, scores=None, gw_constraint=None,
                       precision='f64', best=None):
        """Run many parameter sets at once: params[N, 10] in ``parameters.names`` order.

        Returns a dict of CUDA tensors ('discharge' [n_report, N], 'scores' [N, 8], 'gw' [N]).
        """
        engine = self.get_engine(report=report, gw_constraint=gw_constraint, precision=precision)
        return engine.run(params, discharge=discharge, scores=scores, gw=True, best=best)

    # ------------------------------------------------------------------ outputs
    def write_output_files(self, which='both', parallel=False):
        """Record the discharge time series in output file(s): 'modelled', 'observed' or 'both'."""
        if (which == 'both') or (which == 'modelled'):
            if self.nd_discharge is not None:
                write_flow_file_from_nds(self.timeseries_report[1:], self.nd_discharge,
                                         ''.join([self.out_f, self.catchment, '.mod.flow']),
                                         out_file_format=self.out_fmt, parallel=parallel)
            else:
                raise Exception("The modelled flow output file cannot be written. Please make sure to call the "
                                "simulate method of your SMART instance before writing this output file.")

        if (which == 'both') or (which == 'observed'):
            if self.nd_flow is not None:
                write_flow_file_from_nds(self.timeseries_report[1:], self.nd_flow,
                                         ''.join([self.out_f, self.catchment, '.obs.flow']),
                                         out_file_format=self.out_fmt, parallel=parallel)
            else:
                raise Exception("The observed flow output file cannot be written. Please make sure that a value is "
                                "assigned to the gauged_area_m2 attribute of the SMART class instance.")

    def get_simulation_array(self):
        """Retrieve the simulated discharge time series as a `numpy.ndarray`."""
        if self.nd_discharge is not None:
            return self.nd_discharge
        raise Exception("The simulation array cannot be retrieved. Please make sure to call the simulate "
                        "method of your SMART instance before requesting this output array.")

    def get_evaluation_array(self):
        """Retrieve the observed discharge time series (missing = `numpy.nan`)."""
        if self.nd_flow is not None:
            return self.nd_flow
        raise Exception("The observation array does not exist. Please make sure that a value is assigned "
                        "to the gauged_area_m2 attribute of your SMART class instance.")
