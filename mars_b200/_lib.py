"""ctypes binding of libmarsgb.so (declared in include/mgb.h).

There is no fallback: if the shared library is missing, cannot be built, or CUDA is not usable, every
compute call raises ``MarsGBError`` (SURVEY.md section 8b "Errors": executors raise, the dispatcher wraps
into ExecutionError -- mars/services/subtask/worker/processor.py:210-214).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

MGB_MAX_KEYS = 2
MGB_MAX_VALS = 64
MGB_MAX_ACCS = 64

F64, I64 = 0, 1
SUM, SUMSQ, COUNT, MIN, MAX = 0, 1, 2, 3, 4
OP_NAMES = {SUM: "sum", SUMSQ: "sumsq", COUNT: "count", MIN: "min", MAX: "max"}

_STATUS = {
    1: "MGB_ERR_INVALID",
    2: "MGB_ERR_CUDA",
    3: "MGB_ERR_UNSUPPORTED",
    4: "MGB_ERR_NCCL",
    5: "MGB_ERR_OVERFLOW",
    6: "MGB_ERR_STATE",
}


class MarsGBError(RuntimeError):
    """Raised for every non-OK status of the C ABI and for a missing / unloadable CUDA library."""

    def __init__(self, status: int, message: str):
        self.status = status
        super().__init__(f"{_STATUS.get(status, status)}: {message}")


class AggSpec(C.Structure):
    _fields_ = [
        ("n_keys", C.c_int32),
        ("n_vals", C.c_int32),
        ("n_accs", C.c_int32),
        ("val_dtype", C.c_int32 * MGB_MAX_VALS),
        ("acc_col", C.c_int32 * MGB_MAX_ACCS),
        ("acc_op", C.c_int32 * MGB_MAX_ACCS),
    ]


class AffineFactor(C.Structure):
    _fields_ = [("col", C.c_int32), ("alpha", C.c_double), ("beta", C.c_double)]


class RowProgram(C.Structure):
    _fields_ = [("n_phys", C.c_int32), ("n_factors", C.c_int32 * 8), ("factors", (AffineFactor * 3) * 8),
                ("filter_col", C.c_int32), ("filter_op", C.c_int32), ("filter_dtype", C.c_int32),
                ("filter_i64", C.c_int64), ("filter_f64", C.c_double)]


class PostStep(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a0", C.c_int32), ("a1", C.c_int32), ("a2", C.c_int32), ("ddof", C.c_int32)]


POST_IDENTITY, POST_MEAN, POST_VAR = 0, 1, 2
MGB_MAX_POST = 128

_vp = C.c_void_p
_pp = C.POINTER(C.c_void_p)
_i64 = C.c_int64
_i32 = C.c_int32

# name -> (restype, argtypes); must list every symbol of include/mgb.h (tests/test_abi.py checks this)
SIGNATURES = {
    "mgb_last_error": (C.c_char_p, []),
    "mgb_version": (C.c_int, []),
    "mgb_device_info": (C.c_int, [C.POINTER(_i32), C.POINTER(_i32)]),
    "mgb_hash_keys": (C.c_int, [_pp, _i32, _i64, _vp, _vp]),
    "mgb_partition_ids": (C.c_int, [_pp, _i32, _i64, _i64, _vp, _vp]),
    "mgb_partition_ids_frame": (C.c_int, [_pp, _i32, _i64, _i64, _vp, _vp]),
    "mgb_range_partition_ids": (C.c_int, [_pp, _i32, _i64, _pp, _i32, _vp, _vp]),
    "mgb_partition_scatter": (C.c_int, [_vp, _i64, _i32, _pp, _pp, _i32, _vp, _vp]),
    "mgb_hash_partition": (C.c_int, [_pp, _i32, _i64, _i32, _pp, _pp, _i32, _vp, _vp]),
    "mgb_split_tiles": (C.c_int, [_i64, _i32, _i32, C.POINTER(_i64)]),
    "mgb_split_plan": (C.c_int, [_pp, _i32, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "mgb_split_scatter": (C.c_int, [_pp, _i32, _i32, _vp, _i64, _i32, _vp, _pp, _i32, _pp, _vp, _vp]),
    "mgb_agg_create": (C.c_int, [C.POINTER(_vp), C.POINTER(AggSpec), _vp]),
    "mgb_agg_set_hint": (C.c_int, [_vp, _i32]),
    "mgb_agg_update": (C.c_int, [_vp, _pp, _pp, _i64]),
    "mgb_agg_finalize": (C.c_int, [_vp, C.POINTER(_i64)]),
    "mgb_agg_emit": (C.c_int, [_vp, _pp, _pp, _i32]),
    "mgb_agg_destroy": (C.c_int, [_vp]),
    "mgb_agg_stats": (C.c_int, [_vp, C.POINTER(_i64)]),
    "mgb_estimate_distinct": (C.c_int, [_pp, _i32, _i64, C.POINTER(_i64), C.POINTER(_i64), _vp]),
    "mgb_dense_capacity": (C.c_int, [C.POINTER(_i64)]),
    "mgb_preagg_dense": (C.c_int, [C.POINTER(AggSpec), _pp, _pp, _i64, _pp, _pp, _i64, C.POINTER(_i64), _vp]),
    "mgb_merge_small": (C.c_int, [C.POINTER(AggSpec), _i32, C.POINTER(_i64), C.POINTER(_pp), C.POINTER(_pp), _pp, _pp,
                                  _i64, _i32, C.POINTER(_i64), _vp]),
    "mgb_preagg_dense_async": (C.c_int, [C.POINTER(AggSpec), _pp, _pp, _i64, _pp, _pp, _i64, _vp, _vp]),
    "mgb_merge_small_async": (C.c_int, [C.POINTER(AggSpec), _i32, C.POINTER(_i64), _pp, C.POINTER(_pp), C.POINTER(_pp),
                                        _pp, _pp, _i64, _i32, _vp, _vp]),
    "mgb_dense_batch_max": (C.c_int, [C.POINTER(_i32)]),
    "mgb_dense_last_launches": (C.c_int, [C.POINTER(_i32)]),
    "mgb_preagg_dense_batch_async": (C.c_int, [C.POINTER(AggSpec), _i32, C.POINTER(_i64), C.POINTER(_pp), C.POINTER(_pp),
                                               _pp, _pp, _i64, _vp, _vp]),
    "mgb_preagg_dense_batch_fused_async": (C.c_int, [C.POINTER(AggSpec), C.POINTER(RowProgram), _i32, C.POINTER(_i64),
                                                     C.POINTER(_pp), C.POINTER(_pp), _pp, _pp, _i64, _vp, _vp]),
    "mgb_eval_product": (C.c_int, [_pp, C.POINTER(AffineFactor), _i32, _vp, _i64, _vp]),
    "mgb_filter_rows": (C.c_int, [_vp, _i32, _i32, _i64, C.c_double, _pp, _pp, _i32, _i64, C.POINTER(_i64), _vp]),
    "mgb_mid_capacity": (C.c_int, [C.POINTER(_i64)]),
    "mgb_preagg_mid_batch_async": (C.c_int, [C.POINTER(AggSpec), _i32, C.POINTER(_i64), C.POINTER(_pp), C.POINTER(_pp),
                                             _pp, _pp, _i64, _vp, _vp]),
    "mgb_merge_small_post_async": (C.c_int, [C.POINTER(AggSpec), _i32, C.POINTER(_i64), _pp, C.POINTER(_pp),
                                             C.POINTER(_pp), _pp, _pp, _i64, _i32, C.POINTER(PostStep), _i32, _pp,
                                             _vp, _vp]),
    "mgb_groupby_agg_host": (
        C.c_int,
        [C.POINTER(AggSpec), _pp, _pp, _i64, _i64, _pp, _pp, _i64, _i32, C.POINTER(_i64), _vp],
    ),
    "mgb_row_partial": (C.c_int, [_vp, _i32, _i32, _vp, _i64, _vp]),
    "mgb_host_narrow_i64": (C.c_int, [_vp, _i64, _i32, _vp, C.POINTER(_i32)]),
    "mgb_widen_i64": (C.c_int, [_vp, _i32, _i64, _vp, _vp]),
    "mgb_host_narrow_f64": (C.c_int, [_vp, _i64, _i32, _vp, C.POINTER(_i32)]),
    "mgb_widen_f64": (C.c_int, [_vp, _i32, _i64, _vp, _vp]),
    "mgb_host_narrow_isa": (C.c_int, [_i32]),
    "mgb_post_mean": (C.c_int, [_vp, _i32, _vp, _vp, _i64, _vp]),
    "mgb_post_var": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _i64, _vp]),
    "mgb_copy_rows": (C.c_int, [_vp, _i64, _vp, _i64, _vp]),
    "mgb_copy_segments": (C.c_int, [_pp, _pp, C.POINTER(_i64), _i32, _vp]),
    "mgb_argsort_keys": (C.c_int, [_pp, _i32, _i64, _vp, _vp]),
    "mgb_gather_rows": (C.c_int, [_pp, _pp, _i32, _vp, _i64, _vp]),
    "mgb_nccl_load": (C.c_int, [C.c_char_p]),
    "mgb_comm_unique_id": (C.c_int, [_vp]),
    "mgb_comm_init": (C.c_int, [C.POINTER(_vp), _vp, _i32, _i32]),
    "mgb_comm_alltoallv": (C.c_int, [_vp, _pp, C.POINTER(_i64), _pp, C.POINTER(_i64), _i32, _vp]),
    "mgb_comm_allgather_i64": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "mgb_comm_destroy": (C.c_int, [_vp]),
    "mgb_ipc_arena": (C.c_int, [_i64, C.POINTER(_vp), _vp]),
    "mgb_ipc_open": (C.c_int, [_i32, _vp, C.POINTER(_vp)]),
    "mgb_ipc_pull": (C.c_int, [_vp, _vp, _i64, _vp]),
    "mgb_ipc_arena_slot": (C.c_int, [_i32, _i64, C.POINTER(_vp), _vp, C.POINTER(_i64), C.POINTER(_i32)]),
    "mgb_ipc_open_slot": (C.c_int, [_i32, _i32, _vp, C.POINTER(_vp)]),
    "mgb_ipc_close_slot": (C.c_int, [_i32, _i32]),
    "mgb_prof_enable": (C.c_int, [_i32]),
    "mgb_prof_reset": (C.c_int, []),
    "mgb_prof_read": (C.c_int, [_i32, C.POINTER(C.c_double), C.POINTER(_i64)]),
}

_lock = threading.Lock()
_lib = None


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmarsgb.so")


def load(build_if_missing: bool = True):
    """Load (building in-tree first if needed) and return the ctypes handle.  Raises MarsGBError."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            if not build_if_missing:
                raise MarsGBError(2, f"{path} is missing; run `python -m mars_b200._build`")
            from . import _build

            try:
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise MarsGBError(2, f"libmarsgb.so is missing and could not be built: {e}") from e
        try:
            lib = C.CDLL(path)
        except OSError as e:
            raise MarsGBError(2, f"cannot load {path}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise MarsGBError(2, f"{path} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        if lib.mgb_version() != 100:
            raise MarsGBError(2, f"libmarsgb.so version {lib.mgb_version()} != binding version 100")
        _lib = lib
    return _lib


def check(status: int):
    if status != 0:
        msg = load().mgb_last_error()
        raise MarsGBError(status, msg.decode("utf-8", "replace") if msg else "")


def ptr_array(ptrs):
    """Build a `void*[]` from integer addresses."""
    arr = (C.c_void_p * max(len(ptrs), 1))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def make_spec(n_keys, val_dtypes, accs) -> AggSpec:
    """accs: list of (value_column_index, op)."""
    if not (1 <= n_keys <= MGB_MAX_KEYS):
        raise MarsGBError(3, f"{n_keys} group keys: only 1..{MGB_MAX_KEYS} int64 keys are supported")
    if not (1 <= len(val_dtypes) <= MGB_MAX_VALS):
        raise MarsGBError(3, f"{len(val_dtypes)} value columns outside [1, {MGB_MAX_VALS}]")
    if not (1 <= len(accs) <= MGB_MAX_ACCS):
        raise MarsGBError(3, f"{len(accs)} accumulators outside [1, {MGB_MAX_ACCS}]")
    s = AggSpec()
    s.n_keys, s.n_vals, s.n_accs = n_keys, len(val_dtypes), len(accs)
    for i, d in enumerate(val_dtypes):
        s.val_dtype[i] = d
    for i, (c, op) in enumerate(accs):
        s.acc_col[i] = c
        s.acc_op[i] = op
    return s
