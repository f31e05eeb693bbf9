"""Input/output files of a SMART experiment -- the public surface of the reference's
``smartpy/inout.py`` (:35-322) with the same file formats, checks and error messages.

Host-side and executed once per catchment, so outside the measured hot path; it exists so
that ``SMART(...)`` / ``montecarlo.LHS(...)`` are drop-ins.  Time series are parsed straight
into (epoch-second, value) arrays (``read_*_arrays``) and rescaled with the array cores of
``timeframe``; the reference's dict-returning functions are kept on top of those.
"""
from collections import OrderedDict
from csv import DictReader, writer
from datetime import datetime, timedelta
import argparse

import numpy as np

from .timeframe import (
    check_interval_in_seconds, get_required_resolution, rescale_regular_cumulative_grid,
    rescale_irregular_mean_grid, to_seconds, from_seconds)

# NetCDF: the reference reads/writes it through the optional package netCDF4 and raises these
# messages when the package is missing (inout.py:212-231, :299-310).  That package is outside this
# hot path's scope and not part of the image, so 'netcdf' always takes the
# reference's "package missing" exit here; 'csv' is the supported format.
_NETCDF_IN = ("The use of 'netcdf' as the input file format requires the package 'netCDF4', "
              "please install it and retry, or choose another file format.")
_NETCDF_OUT = ("The use of 'netcdf' as the output file format requires the package 'netCDF4', "
               "please install it and retry, or choose another file format.")


def _parse_stamps(texts):
    """'YYYY-MM-DD HH:MM:SS' strings -> int64 epoch seconds (strict, like strptime)."""
    try:
        for text in texts[:1] + texts[-1:]:
            datetime.strptime(text, "%Y-%m-%d %H:%M:%S")
        return np.array(texts, dtype='datetime64[s]').astype(np.int64)
    except ValueError:
        return np.array([to_seconds(datetime.strptime(t, "%Y-%m-%d %H:%M:%S")) for t in texts], dtype=np.int64)


# ------------------------------------------------------------------ readers (arrays)
def read_csv_series_arrays(csv_file, key_header, val_header):
    """(stamps int64[n], values float64[n]) of a regular CSV series; blank cells -> ValueError."""
    try:
        with open(csv_file, 'r', encoding='utf8') as handle:
            keys, vals = [], []
            try:
                for row in DictReader(handle):
                    keys.append(row[key_header])
                    vals.append(row[val_header])
            except KeyError:
                raise Exception('Field {} or {} does not exist in {}.'.format(key_header, val_header, csv_file))
    except IOError:
        raise Exception('File {} could not be found.'.format(csv_file))
    return _parse_stamps(keys), np.array([np.float64(v) for v in vals], dtype=np.float64)


def read_csv_flow_arrays(csv_file, key_header, val_header):
    """Available observations only: blank cells and the -99 flag are dropped (inout.py:240-246)."""
    try:
        with open(csv_file, 'r', encoding='utf8') as handle:
            keys, vals = [], []
            try:
                for row in DictReader(handle):
                    cell = row[val_header]
                    if cell == '':
                        continue
                    try:
                        value = np.float64(cell)
                    except ValueError:
                        raise Exception('Field {} in {} cannot be converted to float '
                                        'at {}.'.format(val_header, csv_file, row[key_header]))
                    if value != -99.0:
                        keys.append(row[key_header])
                        vals.append(value)
            except KeyError:
                raise Exception('Field {} or {} does not exist in {}.'.format(key_header, val_header, csv_file))
    except IOError:
        raise Exception('File {} could not be found.'.format(csv_file))
    return _parse_stamps(keys), np.array(vals, dtype=np.float64)


def _as_dict(stamps, values, ordered=False):
    out = OrderedDict() if ordered else dict()
    for s, v in zip(stamps.tolist(), values):
        out[from_seconds(s)] = v
    return out


# ------------------------------------------------------------------ reference-shaped readers (dicts)
def read_csv_time_series_with_delta_check(csv_file, key_header, val_header):
    stamps, values = read_csv_series_arrays(csv_file, key_header, val_header)
    first, last, step = check_interval_in_seconds(stamps, csv_file)
    return _as_dict(stamps, values), from_seconds(first), from_seconds(last), timedelta(seconds=step)


def read_csv_time_series_with_missing_check(csv_file, key_header, val_header):
    return _as_dict(*read_csv_flow_arrays(csv_file, key_header, val_header), ordered=True)


def _read_forcing_arrays(file_location, file_format, name):
    if file_format == 'netcdf':
        raise Exception(_NETCDF_IN)
    return read_csv_series_arrays(file_location, 'DateTime', name)


def read_rain_file(file_location, file_format):
    if file_format == 'netcdf':
        raise Exception(_NETCDF_IN)
    return read_csv_time_series_with_delta_check(file_location, key_header='DateTime', val_header='rain')


def read_peva_file(file_location, file_format):
    if file_format == 'netcdf':
        raise Exception(_NETCDF_IN)
    return read_csv_time_series_with_delta_check(file_location, key_header='DateTime', val_header='peva')


def read_flow_file(file_location, file_format):
    if file_format == 'netcdf':
        raise Exception(_NETCDF_IN)
    return read_csv_time_series_with_missing_check(file_location, key_header='DateTime', val_header='flow')


# ------------------------------------------------------------------ series on the simulation / report grids
def get_forcing_series_simu(file_location, file_format, name, start_simu, end_simu, time_delta_simu):
    """Array form of get_dict_{rain,peva}_series_simu: float64[n_steps] stamped start_simu..end_simu."""
    stamps, values = _read_forcing_arrays(file_location, file_format, name)
    first, last, step = check_interval_in_seconds(stamps, file_location)
    start_data, end_data, delta_data = from_seconds(first), from_seconds(last), timedelta(seconds=step)
    if (start_data - delta_data + time_delta_simu <= start_simu) and (end_simu <= end_data):
        res = get_required_resolution(start_data, start_simu, delta_data, time_delta_simu)
        return rescale_regular_cumulative_grid(
            values, first, step, int(res.total_seconds()),
            to_seconds(start_simu), to_seconds(end_simu), int(time_delta_simu.total_seconds()))
    raise Exception('{} data not sufficient for simulation.'.format({'rain': 'Rain', 'peva': 'PEva'}[name]))


def _grid_dict(start, step, series):
    return {start + k * step: series[k] for k in range(len(series))}


def get_dict_rain_series_simu(file_location, file_format, start_simu, end_simu, time_delta_simu):
    series = get_forcing_series_simu(file_location, file_format, 'rain', start_simu, end_simu, time_delta_simu)
    return _grid_dict(start_simu, time_delta_simu, series)


def get_dict_peva_series_simu(file_location, file_format, start_simu, end_simu, time_delta_simu):
    series = get_forcing_series_simu(file_location, file_format, 'peva', start_simu, end_simu, time_delta_simu)
    return _grid_dict(start_simu, time_delta_simu, series)


def get_discharge_series(file_location, file_format, start_report, end_report, catchment_area, gauged_area):
    """Array form of get_dict_discharge_series (inout.py:61-78): float64[n_report], NaN = missing."""
    if file_format == 'netcdf':
        raise Exception(_NETCDF_IN)
    stamps, values = read_csv_flow_arrays(file_location, 'DateTime', 'flow')
    scaling_factor = catchment_area / gauged_area
    # calendar-day window: two days before the first report stamp, one day after the last
    day = stamps // 86400
    first_day = to_seconds(start_report - timedelta(days=2)) // 86400
    last_day = to_seconds(end_report + timedelta(days=1)) // 86400
    keep = (first_day <= day) & (day <= last_day)
    return rescale_irregular_mean_grid(stamps[keep], values[keep] * scaling_factor,
                                       to_seconds(start_report), to_seconds(end_report), 86400, 3600)


def get_dict_discharge_series(file_location, file_format, start_report, end_report, catchment_area, gauged_area):
    series = get_discharge_series(file_location, file_format, start_report, end_report,
                                  catchment_area, gauged_area)
    out = OrderedDict()
    for k in range(series.size):
        out[start_report + k * timedelta(days=1)] = float(series[k])
    return out


# ------------------------------------------------------------------ settings
def read_simulation_settings_file(file_location):
    settings = dict()
    try:
        with open(file_location, 'r', encoding='utf8') as handle:
            for row in DictReader(handle):
                settings[row['ARGUMENT']] = row['VALUE']
    except KeyError:
        raise Exception("There is no 'ARGUMENT' or 'VALUE' column in {}.".format(file_location))
    except IOError:
        raise Exception("There is no simulation file at {}.".format(file_location))
    return settings


def _setting(settings, key, convert, label, what, required=True, default=None):
    try:
        return convert(settings[key])
    except KeyError:
        if required:
            raise Exception('Setting {} is missing from simulation file.'.format(label))
        return default
    except ValueError:
        raise Exception('Setting {} could not be converted to {}.'.format(label, what))


def get_dict_simulation_settings(file_location):
    """(.sttngs) -> c_area, g_area, start, end, delta_simu, delta_report, warm_up, gw_constraint
    (inout.py:81-140)."""
    s = read_simulation_settings_file(file_location)
    stamp = '%d/%m/%Y %H:%M:%S'
    c_area = _setting(s, "catchment_area_km2", lambda v: float(v) * 1e6, 'CATCHMENT AREA', 'a float')
    g_area = _setting(s, "gauged_area_km2", lambda v: float(v) * 1e6, 'GAUGED AREA', 'a float',
                      required=False, default=c_area)
    start = _setting(s, "start_datetime", lambda v: datetime.strptime(v, stamp), 'START',
                     'a datetime [format required: DD/MM/YYYY HH:MM:SS]')
    end = _setting(s, "end_datetime", lambda v: datetime.strptime(v, stamp), 'END',
                   'a datetime [format required: DD/MM/YYYY HH:MM:SS]')
    delta_simu = _setting(s, "simu_timedelta_min", lambda v: timedelta(minutes=int(v)), 'DELTA SIMU',
                          'an integer/timedelta')
    delta_report = _setting(s, "report_timedelta_min", lambda v: timedelta(minutes=int(v)), 'DELTA REPORT',
                            'an integer/timedelta')
    warm_up = _setting(s, "warm_up_days", int, 'WARM UP DURATION', 'an integer')
    gw_constraint = _setting(s, "gw_constraint", float, 'GROUNDWATER CONSTRAINT', 'a float',
                             required=False, default=None)
    return c_area, g_area, start, end, delta_simu, delta_report, warm_up, gw_constraint


# ------------------------------------------------------------------ writers
def write_flow_file_from_nds(series_report, discharge, the_file, out_file_format, parallel=False):
    if out_file_format == 'netcdf':
        raise Exception(_NETCDF_OUT)
    elif out_file_format == 'csv':
        write_flow_csv_file_from_nds(series_report, discharge, the_file)
    else:
        raise Exception("The output format type \'{}\' cannot be written by SMARTpy, "
                        "choose from: \'csv\', \'netcdf\'.".format(out_file_format))


def write_flow_csv_file_from_nds(series_report, discharge, csv_file):
    with open(csv_file, 'w', newline='', encoding='utf8') as handle:
        out = writer(handle, delimiter=',')
        out.writerow(['DateTime', 'flow'])
        out.writerows((dt, '%e' % val) for dt, val in zip(series_report, discharge))


def valid_file_format(fmt):
    if fmt.lower() == "netcdf":
        raise argparse.ArgumentTypeError("NetCDF4 module is not installed, please choose another file format.")
    elif fmt.lower() == "csv":
        return "csv"
    raise argparse.ArgumentTypeError("File format not recognised: '{0}'.".format(fmt))
