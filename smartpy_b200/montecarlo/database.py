"""The sample database of a Monte Carlo run -- the on-disk format of the reference
(``smartpy/montecarlo/montecarlo.py:90-127`` header, ``:211-231`` rows, ``:155-177`` compression,
``:233-262`` reader), written and read in bulk.

Format ('csv'): one header line ``obj_fn_names + param_names + report stamps`` and one line per
sample with every value printed as ``'%.6e'`` of its float32 rounding.  The reference appends one
line per simulation from inside spotpy's loop and reads the file back through ``csv.DictReader``
row by row; here whole blocks of rows are formatted and parsed by the library on all host cores
(``smart_csv_format_f32`` / ``smart_csv_parse_f32``, include/smart_b200_io.h: the same bytes as
``numpy.savetxt(fmt='%.6e')``, 30-50 times faster -- at 1e7 samples the text is the step right after
the hot path) and the reader pulls the wanted columns (looked up BY NAME in the header, as the
reference does, so older files with extra columns still load) straight into float32 arrays.

'netcdf' needs the optional package netCDF4 in the reference and raises when it is missing; that
package is outside this path's scope, so 'netcdf' always takes that exit here (see inout.py).
"""
import gzip
import io
import os
import shutil

import numpy as np

from .. import _native

_BLOCK_ROWS = 1 << 18          # rows formatted / parsed per library call (bounds the text buffer)

NETCDF_MESSAGE = ("The use of 'netcdf' as the output file format requires the package 'netCDF4', "
                  "please install it and retry, or choose another file format.")


def database_path(out_folder, catchment, func, out_format):
    """``<out>/<catchment>.SMART.<func>`` (+ ``.nc`` for netcdf), montecarlo.py:79-82."""
    stem = '{}{}.SMART.{}'.format(out_folder, catchment, func)
    return stem + '.nc' if out_format == 'netcdf' else stem


class SampleDatabase(object):
    """Writer.  ``open()`` -> header; ``write_rows()`` any number of times, rows in sample order;
    ``close(compression)`` -> optional gzip (``<file>.gz`` replaces the file, as the reference)."""

    def __init__(self, path, out_format, columns, stamps=()):
        if out_format == 'netcdf':
            raise Exception(NETCDF_MESSAGE)
        if out_format != 'csv':
            raise Exception("The output format type '{}' cannot be written by SMARTpy, "
                            "choose from: 'csv', 'netcdf'.".format(out_format))
        self.path = path
        self.header = list(columns) + [dt.strftime('%Y-%m-%d %H:%M:%S') for dt in stamps]
        self.n_series = len(stamps)
        self._handle = None
        self.rows_written = 0

    def open(self):
        self._handle = io.open(self.path, 'wb')
        self._handle.write((','.join(self.header) + '\n').encode('utf8'))
        return self

    def write_rows(self, obj_fns, parameters, simulations=None):
        """obj_fns [n, k], parameters [n, 10] (+ simulations [n, n_report] when the run keeps them)."""
        blocks = [np.asarray(obj_fns), np.asarray(parameters)]
        if self.n_series:
            if simulations is None or np.asarray(simulations).shape[1] != self.n_series:
                raise ValueError("the database keeps the simulated series: simulations [n, {}] required".format(
                    self.n_series))
            blocks.append(np.asarray(simulations))
        write_formatted(self._handle, blocks)      # '%.6e' of the float32 value, montecarlo.py:226-231
        self.rows_written += blocks[0].shape[0]

    def close(self, compression=None):
        if self._handle is not None:
            self._handle.close()
            self._handle = None
        if compression is True:
            with io.open(self.path, 'rb') as plain, gzip.open(self.path + '.gz', 'wb') as packed:
                shutil.copyfileobj(plain, packed)
            os.remove(self.path)


def write_formatted(handle, column_blocks):
    """Write the rows [block_0 | block_1 | ...] (arrays with the same number of rows, any float type)
    to the binary handle as numpy.savetxt(fmt='%.6e', delimiter=',') of their float32 rounding would.
    Blocks of _BLOCK_ROWS rows at a time: the float32 row buffer and the text buffer are allocated
    once and reused, and the text goes to the handle without a copy."""
    lib = _native.load()
    column_blocks = [np.asarray(b) for b in column_blocks]
    n = column_blocks[0].shape[0]
    k = sum(b.shape[1] for b in column_blocks)
    if any(b.ndim != 2 or b.shape[0] != n for b in column_blocks):
        raise ValueError("column blocks must be 2-D with the same number of rows")
    if n == 0:
        return
    step = min(n, _BLOCK_ROWS)
    rows = np.empty((step, k), dtype=np.float32)
    text = np.empty(int(lib.smart_csv_bound(step, k)), dtype=np.uint8)
    view = memoryview(text)
    for first in range(0, n, step):
        m = min(step, n - first)
        col = 0
        for b in column_blocks:                      # (the cast to float32 happens in this copy)
            rows[:m, col:col + b.shape[1]] = b[first:first + m]
            col += b.shape[1]
        written = lib.smart_csv_format_f32(rows.ctypes.data, m, k, k, text.ctypes.data, text.size, 0)
        if written < 0:
            _native.check(int(written))
        handle.write(view[:written])


def format_rows(table):
    """float [n, k] -> the bytes numpy.savetxt(fmt='%.6e', delimiter=',') would write for its float32
    rounding."""
    sink = io.BytesIO()
    write_formatted(sink, [np.asarray(table)])
    return sink.getvalue()


def parse_rows(text, n_columns, wanted):
    """bytes of comma-separated lines (no header) -> float32 [n, len(wanted)]: column wanted[k] of every
    line, text -> binary64 -> float32."""
    lib = _native.load()
    wanted = np.ascontiguousarray(wanted, dtype=np.int32)
    buf = np.frombuffer(text, dtype=np.uint8)
    max_rows = text.count(b'\n') + 1
    out = np.empty((max_rows, wanted.size), dtype=np.float32)
    n = lib.smart_csv_parse_f32(buf.ctypes.data, buf.size, int(n_columns), wanted.ctypes.data, wanted.size,
                                out.ctypes.data, max_rows, 0)
    if n < 0:
        raise ValueError(_native.last_error())
    return out[:n]


def read_sample_database(path, out_format, param_names, obj_fn_names, gzipped=False):
    """-> (parameters float32 [N, 10], objective functions float32 [N, k]) of a database written
    by a previous run (montecarlo.py:233-262): columns are found by name in the header."""
    if out_format == 'netcdf':
        raise Exception(NETCDF_MESSAGE)
    opener = (lambda: gzip.open(path + '.gz', 'rb')) if gzipped else (lambda: io.open(path, 'rb'))
    with opener() as handle:
        header = handle.readline().decode('utf8').rstrip('\r\n').split(',')
        try:
            wanted = [header.index(name) for name in list(param_names) + list(obj_fn_names)]
        except ValueError as missing:
            raise KeyError(str(missing))                       # DictReader's row[name] raises KeyError
        # text -> binary64 -> float32, the conversion np.array(list_of_strings, dtype=float32) makes
        table = parse_rows(handle.read(), len(header), wanted)
    n_par = len(param_names)
    return np.ascontiguousarray(table[:, :n_par]), np.ascontiguousarray(table[:, n_par:])
