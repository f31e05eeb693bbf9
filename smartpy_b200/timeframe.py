"""Simulation/reporting time frames and the daily<->hourly rescaling of forcing and
observations -- the public surface of the reference's ``smartpy/timeframe.py`` (TimeFrame at
:26-127, helpers at :130-309), re-implemented on integer epoch-second arrays.

The reference walks dictionaries keyed by ``datetime`` one stamp at a time; here every series
is a regular grid described by (first stamp, step, values) and the rescaling is index
arithmetic on numpy arrays.  The dict-in/dict-out functions of the reference are kept as thin
wrappers.  Floating-point results are bit-identical to the reference's: values are divided
once (``value / divisor``, timeframe.py:181) and re-aggregated by sequential addition in the
same (backwards-in-time) order (timeframe.py:200-203, :291-294).
"""
from datetime import datetime, timedelta
import argparse
from collections import OrderedDict
from math import gcd

import numpy as np

_EPOCH = datetime(1970, 1, 1)


def to_seconds(dt):
    """Naive datetime -> integer seconds since 1970-01-01."""
    delta = dt - _EPOCH
    return delta.days * 86400 + delta.seconds


def from_seconds(sec):
    return _EPOCH + timedelta(seconds=int(sec))


def _whole_seconds(delta):
    return int(delta.total_seconds())


def _datetime_grid(first, last, step):
    """Inclusive list of datetimes first, first+step, ... <= last."""
    count = (to_seconds(last) - to_seconds(first)) // _whole_seconds(step) + 1 if first <= last else 0
    return [first + k * step for k in range(count)]


class TimeFrame(object):
    """Temporal attributes of a run: 'save' is the reporting grid the user asks for, 'simu'
    the model's own discretisation.  Both series carry one extra leading stamp for the
    initial conditions (timeframe.py:84-115).  The save gap must be a multiple of the simu
    gap (timeframe.py:86-87)."""

    def __init__(self, dt_save_start, dt_save_end, simu_increment, save_increment):
        self.save_start = dt_save_start
        self.save_gap = save_increment
        self.save_end = self._check_save_end(dt_save_end)

        self.simu_gap = simu_increment
        self.simu_start, self.simu_end = self._get_simu_start_end_given_save_start_end()

        self.save_series = self._get_list_save_dt_with_initial_conditions()
        self.simu_series = self._get_list_simu_dt_with_initial_conditions()

    def _check_save_end(self, save_end):
        if not self.save_start <= save_end:
            raise Exception("Save Start is greater than Save End.")
        span = int((save_end - self.save_start).total_seconds())
        gap = int(self.save_gap.total_seconds())
        whole, rest = divmod(span, gap)
        if rest != 0:
            latest = self.save_start + timedelta(seconds=self.save_gap.total_seconds()) * whole
            raise Exception("The combination of (start, end) datetimes and the saving time delta "
                            "are not compatible. For a start at {}, and a time delta of {}, the "
                            "latest end in the period is {}".format(self.save_start, self.save_gap, latest))
        return save_end

    def _get_simu_start_end_given_save_start_end(self):
        if not self.save_gap.total_seconds() % self.simu_gap.total_seconds() == 0:
            raise Exception("Save Gap is not greater and a multiple of Simulation Gap.")
        # the first reporting step needs all the simulation steps that it summarises
        return self.save_start - self.save_gap + self.simu_gap, self.save_end

    def _get_list_save_dt_with_initial_conditions(self):
        return _datetime_grid(self.save_start - self.save_gap, self.save_end, self.save_gap)

    def _get_list_simu_dt_with_initial_conditions(self):
        return _datetime_grid(self.simu_start - self.simu_gap, self.simu_end, self.simu_gap)

    def get_gap_simu(self):
        return self.simu_gap

    def get_gap_report(self):
        return self.save_gap

    def get_series_simu(self):
        return self.simu_series

    def get_series_save(self):
        return self.save_series

    # lengths the kernel needs (structure.py:73-75)
    def get_simu_length(self):
        return len(self.simu_series) - 1

    def get_report_gap(self):
        return (len(self.simu_series) - 1) // (len(self.save_series) - 1)


def valid_date(s):
    try:
        return datetime.strptime(s, "%d/%m/%Y_%H:%M:%S")
    except ValueError:
        raise argparse.ArgumentTypeError("Not a valid date: '{0}'.".format(s))


def valid_delta_min(n):
    try:
        return timedelta(minutes=int(n))
    except ValueError:
        raise argparse.ArgumentTypeError("Not a valid time delta: '{0}'.".format(n))


def check_interval_in_seconds(stamps, source):
    """Array form of check_interval_in_list: (first, last, step) in seconds, or raise."""
    stamps = np.asarray(stamps, dtype=np.int64)
    steps = np.unique(np.diff(stamps))
    if steps.size == 1:
        if stamps[0] + steps[0] * (stamps.size - 1) == stamps[-1]:
            return int(stamps[0]), int(stamps[-1]), int(steps[0])
        raise Exception('Missing Data: {} is missing at least one datetime in period.'.format(source))
    raise Exception('Inconsistent Interval: {} does not feature a single time interval.'.format(source))


def check_interval_in_list(list_of_dt, csv_file):
    first, last, step = check_interval_in_seconds([to_seconds(dt) for dt in list_of_dt], csv_file)
    return from_seconds(first), from_seconds(last), timedelta(seconds=step)


def get_required_resolution(start_data, start_simu, delta_data, delta_simu):
    """Coarsest resolution that matches both the data/simu time deltas and their start shift."""
    shift = int((start_data - start_simu).total_seconds())
    return timedelta(seconds=gcd(shift, gcd(_whole_seconds(delta_data), _whole_seconds(delta_simu))))


def _divisor(delta_lo, delta_hi, what):
    whole, rest = divmod(int(delta_lo), int(delta_hi))
    if rest != 0:
        raise Exception("{} Resolution: Time Deltas are not multiples of each other.".format(what))
    if whole < 1:
        raise Exception("{} Resolution: Low resolution lower than higher resolution "
                        "{} < {}.".format(what, timedelta(seconds=int(delta_lo)), timedelta(seconds=int(delta_hi))))
    return whole


# ----------------------------------------------------------------------------------------------
# cumulative data (rain, PET): regular grids
# ----------------------------------------------------------------------------------------------
def rescale_regular_cumulative_grid(values, data_first, data_step, res, simu_first, simu_last, simu_step):
    """Array core of rescale_time_resolution_of_regular_cumulative_data (timeframe.py:211-233).

    values[i] is the total over (data_first + (i-1)*data_step, data_first + i*data_step].
    Returns the totals over each simulation step stamped simu_first .. simu_last.
    """
    values = np.asarray(values, dtype=np.float64)
    up = _divisor(data_step, res, "Increase") if data_step > res else 1
    down = _divisor(simu_step, res, "Decrease")
    portion = values / up if up > 1 else values            # timeframe.py:180-181
    # fine grid: stamp x belongs to the data step that ends at or after it
    fine_first = data_first - (up - 1) * res
    n_fine = values.size * up
    n_simu = (simu_last - simu_first) // simu_step + 1
    stamps = simu_first + simu_step * np.arange(n_simu, dtype=np.int64)
    total = np.zeros(n_simu, dtype=np.float64)
    for k in range(down):                                   # timeframe.py:200-203, same order
        x = stamps - k * res
        if np.any((x - fine_first) % res != 0):
            raise KeyError(from_seconds(x[0]))
        j = (x - fine_first) // res
        if j.min() < 0 or j.max() >= n_fine:
            bad = x[(j < 0) | (j >= n_fine)][0]
            raise KeyError(from_seconds(bad))
        total = total + portion[j // up]
    return total


def _dict_to_grid(dict_info, start, end, step):
    count = (to_seconds(end) - to_seconds(start)) // _whole_seconds(step) + 1
    return np.array([dict_info[start + k * step] for k in range(count)], dtype=np.float64)


def increase_time_resolution_of_regular_cumulative_data(dict_info, start_lo, end_lo,
                                                        time_delta_lo, time_delta_hi):
    """ Use the low resolution to create the high resolution """
    up = _divisor(_whole_seconds(time_delta_lo), _whole_seconds(time_delta_hi), "Increase")
    out = dict()
    stamp = start_lo
    while (start_lo <= stamp) and (stamp <= end_lo):
        portion = dict_info[stamp] / up
        for k in range(up):
            out[stamp - k * time_delta_hi] = portion
        stamp += time_delta_lo
    return out


def decrease_time_resolution_of_regular_cumulative_data(dict_info, start_lo, end_lo,
                                                        time_delta_lo, time_delta_hi):
    """ Use the high resolution to create the low resolution """
    down = _divisor(_whole_seconds(time_delta_lo), _whole_seconds(time_delta_hi), "Decrease")
    out = dict()
    stamp = start_lo
    while (start_lo <= stamp) and (stamp <= end_lo):
        total = 0.0
        for k in range(down):
            total += dict_info[stamp - k * time_delta_hi]
        out[stamp] = total
        stamp += time_delta_lo
    return out


def rescale_time_resolution_of_regular_cumulative_data(dict_data,
                                                       start_data, end_data, time_delta_data,
                                                       time_delta_res,
                                                       start_simu, end_simu, time_delta_simu):
    values = _dict_to_grid(dict_data, start_data, end_data, time_delta_data)
    series = rescale_regular_cumulative_grid(
        values, to_seconds(start_data), _whole_seconds(time_delta_data), _whole_seconds(time_delta_res),
        to_seconds(start_simu), to_seconds(end_simu), _whole_seconds(time_delta_simu))
    return {start_simu + k * time_delta_simu: series[k] for k in range(series.size)}


# ----------------------------------------------------------------------------------------------
# mean data (observed discharge): irregular stamps with gaps
# ----------------------------------------------------------------------------------------------
def rescale_irregular_mean_grid(stamps, values, start, end, delta_lo, delta_hi):
    """Array core of rescale_time_resolution_of_irregular_mean_data (timeframe.py:236-309).

    stamps/values: the available observations (missing ones simply absent), seconds / float.
    Each value is replicated backwards over the high-resolution steps since the previous
    stamp (or over delta_lo when the gap is >= 1.5 delta_lo), then the report grid
    start..end (step delta_lo) takes the mean of its delta_lo/delta_hi high-resolution
    steps, NaN when any of them is not covered.
    """
    stamps = np.asarray(stamps, dtype=np.int64)
    values = np.asarray(values, dtype=np.float64)
    n_out = (end - start) // delta_lo + 1
    if stamps.size == 0:
        return np.full(n_out, np.nan)
    prev = np.concatenate([[stamps[0] - delta_lo], stamps[:-1]])
    span = stamps - prev
    span = np.where(span >= 1.5 * delta_lo, delta_lo, span)          # timeframe.py:250-251
    if np.any(span % delta_hi != 0):
        raise Exception("Increase Resolution: Time Deltas are not multiples of each other.")
    reps = span // delta_hi
    if np.any(reps < 1):
        raise Exception("Increase Resolution: Low resolution lower than higher resolution.")
    # only stamps on the report grid's high-resolution lattice can ever be looked up
    on_lattice = (stamps - start) % delta_hi == 0
    stamps, values, reps = stamps[on_lattice], values[on_lattice], reps[on_lattice]
    if stamps.size == 0:
        return np.full(n_out, np.nan)
    lo_stamp = int(min((stamps - (reps - 1) * delta_hi).min(), start - (delta_lo - delta_hi)))
    hi_stamp = int(max(stamps.max(), end))
    n_hi = (hi_stamp - lo_stamp) // delta_hi + 1
    fine = np.full(n_hi, np.nan)
    covered = np.zeros(n_hi, dtype=bool)
    # consecutive spans cannot overlap (each reaches back at most to the previous stamp), so
    # the reference's "overwriting" guard (timeframe.py:266-267) can never fire here
    owner = np.repeat(np.arange(stamps.size), reps)
    back = np.arange(int(reps.sum())) - np.repeat(np.cumsum(reps) - reps, reps)
    j = (stamps[owner] - back * delta_hi - lo_stamp) // delta_hi
    fine[j] = values[owner]
    covered[j] = True

    down = _divisor(delta_lo, delta_hi, "Decrease")
    out_stamps = start + delta_lo * np.arange(n_out, dtype=np.int64)
    total = np.zeros(n_out, dtype=np.float64)
    ok = np.ones(n_out, dtype=bool)
    for k in range(down):                                             # timeframe.py:291-294
        jj = (out_stamps - k * delta_hi - lo_stamp) // delta_hi
        inside = (jj >= 0) & (jj < n_hi)
        jj_c = np.clip(jj, 0, n_hi - 1)
        ok &= inside & covered[jj_c]
        total = total + np.where(ok, fine[jj_c], 0.0)
    return np.where(ok, total / down, np.nan)


def increase_time_resolution_of_irregular_mean_data(dict_info, time_delta_lo, time_delta_hi):
    """
    Create high resolution mean data from lower resolution mean data
    using backwards replication.
    """
    out = dict()
    stamps = list(dict_info)
    previous = stamps[0] - time_delta_lo
    for stamp in stamps:
        span = stamp - previous
        if span >= timedelta(seconds=1.5 * time_delta_lo.total_seconds()):
            span = time_delta_lo
        reps = _divisor(_whole_seconds(span), _whole_seconds(time_delta_hi), "Increase")
        try:
            value = float(dict_info[stamp])
        except ValueError:
            value = float('nan')
        for k in range(reps):
            if out.get(stamp - k * time_delta_hi):
                raise Exception("Increase Resolution: Overwriting already existing data for datetime.")
            out[stamp - k * time_delta_hi] = value
        previous = stamp
    return out


def decrease_time_resolution_of_irregular_mean_data(dict_info, dt_start, dt_end, time_delta_hi, time_delta_lo):
    """ Creates low resolution cumulative data from high resolution cumulative data
    using arithmetic mean. """
    down = _divisor(_whole_seconds(time_delta_lo), _whole_seconds(time_delta_hi), "Decrease")
    out = OrderedDict()
    stamp = dt_start
    while (dt_start <= stamp) and (stamp <= dt_end):
        try:
            total = 0.0
            for k in range(down):
                total += dict_info[stamp - k * time_delta_hi]
            out[stamp] = total / down
        except (KeyError, TypeError):
            out[stamp] = float('nan')
        stamp += time_delta_lo
    return out


def rescale_time_resolution_of_irregular_mean_data(dict_data, start_data, end_data, time_delta_lo, time_delta_hi):
    stamps = np.array([to_seconds(dt) for dt in dict_data], dtype=np.int64)
    values = np.array([float(v) for v in dict_data.values()], dtype=np.float64)
    series = rescale_irregular_mean_grid(stamps, values, to_seconds(start_data), to_seconds(end_data),
                                         _whole_seconds(time_delta_lo), _whole_seconds(time_delta_hi))
    out = OrderedDict()
    for k in range(series.size):
        out[start_data + k * time_delta_lo] = float(series[k])
    return out
