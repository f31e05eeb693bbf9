"""ctypes binding of libsmart_b200.so (the C ABI declared in include/smart_b200.h).

There is deliberately NO fallback here: if the shared library is missing or no CUDA device
is visible, every compute entry point raises.  The library is built in-tree by
``smartpy_b200._build.build()`` (``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SMART_B200_LIB lets kernel-tuning experiments load an alternative build of the same ABI
LIB_PATH = os.environ.get("SMART_B200_LIB") or os.path.join(_HERE, "libsmart_b200.so")

N_PARAMS = 10
N_VARS = 19
N_SCORES = 8
OBS_STATS = 6
REPORT_SUMMARY = 1
REPORT_RAW = 2

OK = 0
ERR_BAD_ARG = -1
ERR_WARMUP_TOO_LONG = -2
ERR_GAP = -3
ERR_CUDA = -4
ERR_NO_DEVICE = -5

FLAG_FORCE_GENERAL = 0x1
FLAG_NO_TMA = 0x2


def flag_relay_segs(n):
    """SMART_FLAG_RELAY_SEGS(n): ask for a relay of about n segments whatever the batch size."""
    return (int(n) & 0xff) << 8

# every symbol include/smart_b200.h declares
SYMBOLS = (
    "smart_version", "smart_last_error", "smart_launch_count", "smart_batch_n_report", "smart_batch_workspace_bytes",
    "smart_obs_stats", "smart_batch_run_f64", "smart_batch_run_f32", "smart_score_discharge",
    "smart_disaggregate", "smart_expand", "smart_stamp", "smart_batch_run_host",
    "smart_allsteps_host", "smart_host_arena_release", "smart_fma_peak_probe",
    "smart_member_order_len", "smart_member_order_workspace_bytes", "smart_member_order", "smart_fold_blocks",
    "smart_condition_workspace_bytes", "smart_condition_rows", "smart_best_rows",
    "smart_lhs_rows",
)
# ... and include/smart_b200_io.h (host-side bulk text I/O of the sample database)
IO_SYMBOLS = ("smart_csv_bound", "smart_csv_format_f32", "smart_csv_parse_f32")

MAX_CONDITIONS = 8
COND_KINDS = {'equal': 0, 'min': 1, 'max': 2, 'inside': 3, 'outside': 4}


class Condition(ctypes.Structure):
    """Mirror of ``smart_condition`` (include/smart_b200.h)."""
    _fields_ = [("column", ctypes.c_int32), ("kind", ctypes.c_int32), ("lo", ctypes.c_double),
                ("hi", ctypes.c_double)]


class BatchDesc(ctypes.Structure):
    """Mirror of ``smart_batch_desc`` (include/smart_b200.h), field for field."""
    _fields_ = [
        ("n_members", ctypes.c_int64),
        ("n_steps", ctypes.c_int64),
        ("n_warmup", ctypes.c_int64),
        ("n_catchments", ctypes.c_int32),
        ("members_per_catchment", ctypes.c_int32),
        ("report_gap", ctypes.c_int32),
        ("report_type", ctypes.c_int32),
        ("flags", ctypes.c_uint32),
        ("forcing_repeat", ctypes.c_int32),
        ("dt_sec", ctypes.c_double),
        ("params", ctypes.c_void_p),
        ("rain", ctypes.c_void_p),
        ("peva", ctypes.c_void_p),
        ("area_m2", ctypes.c_void_p),
        ("obs", ctypes.c_void_p),
        ("obs_stats", ctypes.c_void_p),
        ("initial_state", ctypes.c_void_p),
        ("has_extra", ctypes.c_int32),
        ("reserved1", ctypes.c_int32),
        ("aar", ctypes.c_double),
        ("ro_ratio", ctypes.c_double),
        ("ro_split", ctypes.c_double * 5),
        ("gw_constraint", ctypes.c_double),
        ("discharge", ctypes.c_void_p),
        ("ld_discharge", ctypes.c_int64),
        ("scores", ctypes.c_void_p),
        ("gw", ctypes.c_void_p),
        ("last_state", ctypes.c_void_p),
        ("best_column", ctypes.c_int32),
        ("best_sign", ctypes.c_int32),
        ("best_score", ctypes.c_void_p),
        ("best_index", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p),
        ("member_order", ctypes.c_void_p),
        ("ld_scores", ctypes.c_int64),
        ("ld_gw", ctypes.c_int64),
        ("member_order_len", ctypes.c_int64),
        ("workspace_bytes", ctypes.c_int64),
    ]


class NativeLibraryMissing(RuntimeError):
    pass


_lib = None


def load():
    """Load libsmart_b200.so; raise loudly if it was not built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            "{} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). smartpy_b200 has no CPU fallback.".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    pdesc = ctypes.POINTER(BatchDesc)
    lib.smart_version.restype = ctypes.c_int
    lib.smart_version.argtypes = []
    lib.smart_last_error.restype = ctypes.c_char_p
    lib.smart_last_error.argtypes = []
    lib.smart_launch_count.restype = ctypes.c_int64
    lib.smart_launch_count.argtypes = []
    lib.smart_host_arena_release.restype = ctypes.c_int
    lib.smart_host_arena_release.argtypes = []
    lib.smart_member_order_len.restype = ctypes.c_int64
    lib.smart_member_order_len.argtypes = [ctypes.c_int64]
    lib.smart_member_order_workspace_bytes.restype = ctypes.c_size_t
    lib.smart_member_order_workspace_bytes.argtypes = [ctypes.c_int64]
    lib.smart_member_order.restype = ctypes.c_int
    lib.smart_member_order.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]
    lib.smart_fold_blocks.restype = ctypes.c_int
    lib.smart_fold_blocks.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p]
    lib.smart_batch_n_report.restype = ctypes.c_int64
    lib.smart_batch_n_report.argtypes = [pdesc]
    lib.smart_batch_workspace_bytes.restype = ctypes.c_size_t
    lib.smart_batch_workspace_bytes.argtypes = [pdesc]
    lib.smart_obs_stats.restype = ctypes.c_int
    lib.smart_obs_stats.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p,
                                    ctypes.c_void_p]
    for name in ("smart_batch_run_f64", "smart_batch_run_f32"):
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [pdesc, ctypes.c_void_p]
    lib.smart_score_discharge.restype = ctypes.c_int
    lib.smart_score_discharge.argtypes = [
        ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_int32, ctypes.c_int32, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.smart_disaggregate.restype = ctypes.c_int
    lib.smart_disaggregate.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_void_p, ctypes.c_void_p]
    lib.smart_expand.restype = ctypes.c_int
    lib.smart_expand.argtypes = lib.smart_disaggregate.argtypes
    lib.smart_stamp.restype = ctypes.c_int
    lib.smart_stamp.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                                ctypes.c_void_p, ctypes.c_void_p]
    lib.smart_batch_run_host.restype = ctypes.c_int
    lib.smart_batch_run_host.argtypes = [pdesc, ctypes.c_int, ctypes.c_int]
    lib.smart_allsteps_host.restype = ctypes.c_int
    lib.smart_allsteps_host.argtypes = [
        ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.smart_fma_peak_probe.restype = ctypes.c_int
    lib.smart_fma_peak_probe.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_void_p]
    lib.smart_condition_workspace_bytes.restype = ctypes.c_size_t
    lib.smart_condition_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64]
    pcond = ctypes.POINTER(Condition)
    lib.smart_condition_rows.restype = ctypes.c_int
    lib.smart_condition_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, pcond, ctypes.c_int32,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.smart_best_rows.restype = ctypes.c_int
    lib.smart_best_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, pcond,
                                    ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p]
    lib.smart_lhs_rows.restype = ctypes.c_int
    lib.smart_lhs_rows.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.smart_csv_bound.restype = ctypes.c_int64
    lib.smart_csv_bound.argtypes = [ctypes.c_int64, ctypes.c_int32]
    lib.smart_csv_format_f32.restype = ctypes.c_int64
    lib.smart_csv_format_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32]
    lib.smart_csv_parse_f32.restype = ctypes.c_int64
    lib.smart_csv_parse_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p,
                                        ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32]
    _lib = lib
    return lib


def last_error():
    return load().smart_last_error().decode("utf8", "replace")


def check(rc):
    """Translate a negative return code into the exception the reference raises there."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_GAP:
        raise ValueError(msg)          # numpy's reshape error at structure.py:190
    if rc == ERR_NO_DEVICE:
        raise RuntimeError(msg)
    raise Exception(msg)               # the reference raises bare Exception (structure.py:70, :91-95)
