"""Tensor-level wrappers over the C ABI (include/mgb.h).  PyTorch is used for allocation and streams only.

Every function takes CUDA tensors (int64 keys; float64 / int64 values), passes raw device pointers and the
current CUDA stream to libmarsgb.so and returns freshly allocated CUDA tensors.  Nothing here computes on
the CPU: a non-CUDA tensor raises ``MarsGBError`` (no fallback path exists).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import F64, I64, SUM, SUMSQ, COUNT, MIN, MAX, MarsGBError, check  # noqa: F401
from ._lib import POST_IDENTITY, POST_MEAN, POST_VAR, MGB_MAX_POST as MAX_POST  # noqa: F401

# number of libmarsgb kernel launches issued through this module (bench.py reports it as gpu_launches)
launch_counter = 0


def _count(n):
    global launch_counter
    launch_counter += n


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the stream every kernel is launched on)."""
    if not torch.cuda.is_available():
        raise MarsGBError(2, "no CUDA device: the groupby path runs on the GPU only (no CPU fallback)")
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, what: str, dtypes=(torch.int64, torch.float64)) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise MarsGBError(1, f"{what}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise MarsGBError(2, f"{what}: tensor lives on {t.device}; the groupby path runs on CUDA only "
                             f"(no CPU fallback)")
    if t.dtype not in dtypes:
        raise MarsGBError(3, f"{what}: dtype {t.dtype} not supported (need one of {dtypes})")
    if t.dim() != 1:
        raise MarsGBError(1, f"{what}: expected a 1-D column, got shape {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def dtype_code(t: torch.Tensor) -> int:
    return F64 if t.dtype == torch.float64 else I64


def _ptrs(ts: Sequence[torch.Tensor]):
    return _lib.ptr_array([t.data_ptr() for t in ts])


def device_info() -> Tuple[int, int]:
    lib = _lib.load()
    a, b = C.c_int32(), C.c_int32()
    check(lib.mgb_device_info(C.byref(a), C.byref(b)))
    return a.value, b.value


# ---------------------------------------------------------------------------------------------- hashing
def hash_keys(keys: Sequence[torch.Tensor]) -> torch.Tensor:
    """uint64 pandas hash of each key tuple, returned as the int64 bit pattern (mgb_hash_keys)."""
    lib = _lib.load()
    keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
    n = keys[0].numel()
    out = torch.empty(n, dtype=torch.int64, device=keys[0].device)
    check(lib.mgb_hash_keys(_ptrs(keys), len(keys), n, out.data_ptr(), _stream()))
    _count(1 if n else 0)
    return out


def partition_ids(keys: Sequence[torch.Tensor], n_partitions: int) -> torch.Tensor:
    """int32 reducer index `hash % n_partitions` per row (mgb_partition_ids)."""
    lib = _lib.load()
    keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
    n = keys[0].numel()
    out = torch.empty(n, dtype=torch.int32, device=keys[0].device)
    check(lib.mgb_partition_ids(_ptrs(keys), len(keys), n, int(n_partitions), out.data_ptr(), _stream()))
    _count(1 if n else 0)
    return out


def partition_ids_frame(keys: Sequence[torch.Tensor], n_partitions: int) -> torch.Tensor:
    """int32 reducer index of every ROW of the given key columns, pandas DataFrame hashing semantics
    (`hash_dataframe_on(df, on=[...], size)`: mgb_partition_ids_frame)."""
    lib = _lib.load()
    keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
    n = keys[0].numel()
    out = torch.empty(n, dtype=torch.int32, device=keys[0].device)
    check(lib.mgb_partition_ids_frame(_ptrs(keys), len(keys), n, int(n_partitions), out.data_ptr(), _stream()))
    _count(1 if n else 0)
    return out


def range_partition_ids(keys: Sequence[torch.Tensor], pivots: Sequence[torch.Tensor]) -> torch.Tensor:
    """int32 range partition of every key tuple: number of pivot tuples strictly smaller than it
    (mgb_range_partition_ids; PSRS shuffle of groupby(sort=True))."""
    lib = _lib.load()
    keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
    pivots = [_req(p, f"pivot {i}", (torch.int64,)) for i, p in enumerate(pivots)]
    n = keys[0].numel()
    out = torch.empty(n, dtype=torch.int32, device=keys[0].device)
    check(lib.mgb_range_partition_ids(_ptrs(keys), len(keys), n, _ptrs(pivots), pivots[0].numel() if pivots else 0,
                                      out.data_ptr(), _stream()))
    _count(1 if n else 0)
    return out


def partition_scatter(pid: torch.Tensor, n_partitions: int, cols: Sequence[torch.Tensor]):
    """Stable split of rows into per-partition blocks.  Returns (out_cols, offsets) with offsets a host
    list of n_partitions+1 ints (one small device->host copy)."""
    lib = _lib.load()
    pid = _req(pid, "pid", (torch.int32,))
    cols = [_req(c, f"column {i}") for i, c in enumerate(cols)]
    n = pid.numel()
    dev = pid.device
    outs = [torch.empty_like(c) for c in cols]
    offsets = torch.empty(n_partitions + 1, dtype=torch.int64, device=dev)
    check(lib.mgb_partition_scatter(pid.data_ptr(), n, int(n_partitions), _ptrs(cols), _ptrs(outs), len(cols),
                                    offsets.data_ptr(), _stream()))
    _count(7 if n else 1)
    return outs, offsets.cpu().tolist()


def hash_partition(keys: Sequence[torch.Tensor], n_partitions: int, cols: Sequence[torch.Tensor]):
    """a5 in one call (mgb_hash_partition): reducer index = pandas hash of the key tuple % n_partitions, stable
    split of ``cols`` into per-reducer blocks.  Returns (out_cols, offsets) with offsets a host list of
    n_partitions + 1 ints -- the only host wait of the call."""
    lib = _lib.load()
    keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
    cols = [_req(c, f"column {i}") for i, c in enumerate(cols)]
    n = keys[0].numel()
    outs = [torch.empty_like(c) for c in cols]
    offsets = torch.empty(n_partitions + 1, dtype=torch.int64, device=keys[0].device)
    check(lib.mgb_hash_partition(_ptrs(keys), len(keys), n, int(n_partitions), _ptrs(cols), _ptrs(outs), len(cols),
                                 offsets.data_ptr(), _stream()))
    _count(6 if n else 1)     # histogram, scan (3), offsets, scatter
    return outs, offsets.cpu().tolist()


SPLIT_LONG_RUNS = 4     # mode bit: the rows will be stored into peer GPUs (8192-row tiles: longer runs over NVLink)


def split_plan(keys: Sequence[torch.Tensor], n_partitions: int, mode: int = 0):
    """pass 1 of the partition kernel (mgb_split_plan): -> (tile_hist, offsets), both on the device; offsets holds
    the n_partitions + 1 block boundaries.  No host wait.  ``mode``: 0 Index hash, 1 DataFrame-row hash, optionally
    | SPLIT_LONG_RUNS -- the same mode must be given to split_scatter."""
    lib = _lib.load()
    keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
    n = keys[0].numel()
    tiles = C.c_int64()
    check(lib.mgb_split_tiles(n, int(n_partitions), int(mode), C.byref(tiles)))
    dev = keys[0].device
    hist = torch.empty(max(1, tiles.value * n_partitions), dtype=torch.int64, device=dev)
    offsets = torch.empty(n_partitions + 1, dtype=torch.int64, device=dev)
    check(lib.mgb_split_plan(_ptrs(keys), len(keys), int(mode), None, n, int(n_partitions), hist.data_ptr(),
                             offsets.data_ptr(), _stream()))
    _count(5 if n else 1)
    return hist, offsets


def split_scatter(keys: Sequence[torch.Tensor], n_partitions: int, hist: torch.Tensor, cols: Sequence[torch.Tensor],
                  outs: Optional[Sequence[torch.Tensor]] = None, table: Optional[torch.Tensor] = None, mode: int = 0):
    """pass 2 (mgb_split_scatter): every row of ``cols`` written once to its block -- into ``outs`` (block p of column
    c at outs[c][offsets[p]:offsets[p + 1]]) or, with ``table`` (device int64 [n_cols * n_partitions] of addresses,
    possibly in IPC-mapped peer memory), to (uint64*)table[c * n_partitions + p] + row-in-block."""
    lib = _lib.load()
    keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
    cols = [_req(c, f"column {i}") for i, c in enumerate(cols)]
    n = keys[0].numel()
    if table is not None:
        if not table.is_cuda or table.dtype != torch.int64 or table.numel() != len(cols) * n_partitions:
            raise MarsGBError(1, "destination table must be a CUDA int64 tensor of n_cols * n_partitions addresses")
    elif outs is None or len(outs) != len(cols):
        raise MarsGBError(1, "split_scatter needs output columns or a destination table")
    check(lib.mgb_split_scatter(_ptrs(keys), len(keys), int(mode), None, n, int(n_partitions), hist.data_ptr(),
                                _ptrs(cols), len(cols), _ptrs(outs) if table is None else None,
                                table.data_ptr() if table is not None else None, _stream()))
    _count(-(-len(cols) // 80) if n and cols else 0)


# ------------------------------------------------------------------------------------------ aggregation
class Aggregator:
    """Streaming hash aggregation: update() any number of row batches, then result()."""

    def __init__(self, n_keys: int, val_dtypes: Sequence[int], accs: Sequence[Tuple[int, int]],
                 expect_distinct: bool = False):
        self._lib = _lib.load()
        self.n_keys = n_keys
        self.val_dtypes = list(val_dtypes)
        self.accs = list(accs)
        self._spec = _lib.make_spec(n_keys, self.val_dtypes, self.accs)
        self._h = C.c_void_p()
        self._keep = []  # batches must stay alive until finalize
        self._device = None
        self.stats = None
        check(self._lib.mgb_agg_create(C.byref(self._h), C.byref(self._spec), _stream()))
        if expect_distinct:   # skip the shared-memory pre-aggregation pass (mgb_agg_set_hint)
            check(self._lib.mgb_agg_set_hint(self._h, 1))

    def update(self, keys: Sequence[torch.Tensor], vals: Sequence[torch.Tensor]):
        keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
        vals = [_req(v, f"value {i}") for i, v in enumerate(vals)]
        if len(keys) != self.n_keys or len(vals) != len(self.val_dtypes):
            raise MarsGBError(1, "column count does not match the aggregator spec")
        for i, v in enumerate(vals):
            if dtype_code(v) != self.val_dtypes[i]:
                raise MarsGBError(1, f"value column {i} dtype {v.dtype} does not match the spec")
            if v.numel() != keys[0].numel():
                raise MarsGBError(1, "ragged batch: columns differ in length")
        self._device = keys[0].device
        n = keys[0].numel()
        # one update() handles < 2^31 rows; larger batches are split here
        step = (1 << 31) - 1024
        for r0 in range(0, n, step):
            ks = [k[r0:r0 + step] for k in keys]
            vs = [v[r0:r0 + step] for v in vals]
            self._keep.append((ks, vs))
            check(self._lib.mgb_agg_update(self._h, _ptrs(ks), _ptrs(vs), ks[0].numel()))
        return self

    def update_ptrs(self, key_ptrs: Sequence[int], val_ptrs: Sequence[int], n: int, keep, device=None):
        """``update`` for callers that hold validated columns and already know their device pointers (row
        ranges of scattered columns: no tensor views are created).  ``keep`` is whatever keeps the memory
        alive until finalize."""
        if len(key_ptrs) != self.n_keys or len(val_ptrs) != len(self.val_dtypes):
            raise MarsGBError(1, "column count does not match the aggregator spec")
        if n >= (1 << 31) - 1024:
            raise MarsGBError(4, "a pointer-level batch must hold fewer than 2^31 rows")
        if device is not None:
            self._device = device
        if n > 0:
            self._keep.append(keep)
            check(self._lib.mgb_agg_update(self._h, _lib.ptr_array(key_ptrs), _lib.ptr_array(val_ptrs), n))
        return self

    def result(self, sort: bool = False, device=None):
        """-> (key tensors, accumulator tensors)."""
        g = C.c_int64()
        check(self._lib.mgb_agg_finalize(self._h, C.byref(g)))
        self._keep.clear()
        G = g.value
        dev = self._device or device or torch.device("cuda", torch.cuda.current_device())
        keys = [torch.empty(G, dtype=torch.int64, device=dev) for _ in range(self.n_keys)]
        accs = []
        for col, op in self.accs:
            dt = torch.int64 if (op == COUNT or self.val_dtypes[col] == I64) else torch.float64
            accs.append(torch.empty(G, dtype=dt, device=dev))
        check(self._lib.mgb_agg_emit(self._h, _ptrs(keys), _ptrs(accs), 1 if sort else 0))
        st = (C.c_int64 * 6)()
        check(self._lib.mgb_agg_stats(self._h, st))
        self.stats = dict(bucket_fallback=st[0], rows_smem=st[1], rows_spilled=st[2], partial_rows=st[3],
                          table_capacity=st[4], launches=st[5])
        _count(st[5])
        return keys, accs

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mgb_agg_destroy(self._h)
            self._h = C.c_void_p()
        self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def groupby_aggregate(keys, vals, accs, sort=False, expect_distinct=False):
    """One-shot: (key tensors, accumulator tensors, stats)."""
    agg = Aggregator(len(keys), [dtype_code(v) for v in vals], accs, expect_distinct=expect_distinct)
    try:
        agg.update(keys, vals)
        k, a = agg.result(sort=sort, device=keys[0].device)
        return k, a, agg.stats
    finally:
        agg.close()


def estimate_distinct(keys: Sequence[torch.Tensor]) -> Tuple[int, int]:
    """(rows sampled, distinct key hashes among them) of a strided sample of the chunk (mgb_estimate_distinct)"""
    lib = _lib.load()
    keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
    a, b = C.c_int64(), C.c_int64()
    check(lib.mgb_estimate_distinct(_ptrs(keys), len(keys), keys[0].numel(), C.byref(a), C.byref(b), _stream()))
    _count(1 if keys[0].numel() else 0)
    return a.value, b.value


# ------------------------------------------------------------------------------ low-cardinality fast paths
_spec_cache = {}
_dense_cap = None


def _spec(n_keys, val_dtypes, accs):
    key = (n_keys, tuple(val_dtypes), tuple(accs))
    sp = _spec_cache.get(key)
    if sp is None:
        sp = _spec_cache[key] = _lib.make_spec(n_keys, list(val_dtypes), list(accs))
    return sp


def _out_dtype(val_dtypes, col, op):
    return torch.int64 if (op == COUNT or val_dtypes[col] == I64) else torch.float64


def _alloc_block(n_keys, val_dtypes, accs, cap, device):
    """One [n_keys + n_accs, cap] buffer; returns (block, key views, accumulator views)."""
    block = torch.empty((n_keys + len(accs), max(cap, 1)), dtype=torch.int64, device=device)
    keys = [block[j] for j in range(n_keys)]
    outs = []
    for a, (col, op) in enumerate(accs):
        row = block[n_keys + a]
        outs.append(row if _out_dtype(val_dtypes, col, op) == torch.int64 else row.view(torch.float64))
    return block, keys, outs


class _FastCall:
    """Pre-built ctypes argument arrays of one call shape: per call only raw pointers are filled in."""

    __slots__ = ("spec", "nk", "nv", "na", "kp", "vp", "okp", "oap", "is_float", "n_int", "n_float", "int_row",
                 "float_row", "g", "col_dtypes", "acc_dts")

    def __init__(self, nk, dts, accs):
        self.spec = _lib.make_spec(nk, list(dts), list(accs))
        self.nk, self.nv, self.na = nk, len(dts), len(accs)
        self.kp = (C.c_void_p * nk)()
        self.vp = (C.c_void_p * max(self.nv, 1))()
        self.okp = (C.c_void_p * nk)()
        self.oap = (C.c_void_p * max(self.na, 1))()
        self.is_float = [_out_dtype(dts, col, op) == torch.float64 for col, op in accs]
        self.int_row, self.float_row = [], []
        ni, nf = nk, 0
        for f in self.is_float:
            if f:
                self.float_row.append(nf)
                self.int_row.append(-1)
                nf += 1
            else:
                self.int_row.append(ni)
                self.float_row.append(-1)
                ni += 1
        self.n_int, self.n_float = ni, nf
        self.g = C.c_int64()
        # torch dtypes of the output columns (keys, then accumulators) and dtype codes of the accumulators
        self.col_dtypes = [torch.int64] * nk + [torch.float64 if f else torch.int64 for f in self.is_float]
        self.acc_dts = [F64 if f else I64 for f in self.is_float]

    def alloc(self, cap, device):
        """int64 block [n_keys + int accs + 1, cap] and float64 block [float accs, cap]; output pointers set.
        The last int64 row is scratch for the device-side group count of the asynchronous entry points."""
        cap = max(int(cap), 1)
        # ONE allocation: the float64 block is a dtype view of the rows behind the int64 block, so a consumer
        # can ship / read back the whole result (count row included) as a single tensor (``bi._base``)
        ni = self.n_int + 1
        blk = torch.empty((ni + max(self.n_float, 1), cap), dtype=torch.int64, device=device)
        bi = blk[:ni]
        bf = blk[ni:].view(torch.float64)
        pi, pf, stride = bi.data_ptr(), bf.data_ptr(), cap * 8
        for j in range(self.nk):
            self.okp[j] = pi + j * stride
        for a in range(self.na):
            self.oap[a] = pf + self.float_row[a] * stride if self.is_float[a] else pi + self.int_row[a] * stride
        return bi, bf

    def count_ptr(self, bi) -> int:
        return bi.data_ptr() + self.n_int * bi.shape[1] * 8

    def views(self, bi, bf, G=None):
        """(key tensors, accumulator tensors) of the first G rows (all rows when G is None) -- two slices +
        two unbinds in total."""
        ri = (bi if G is None else bi[:, :G]).unbind(0)
        rf = (bf if G is None else bf[:, :G]).unbind(0)
        keys = list(ri[:self.nk])
        accs = [rf[self.float_row[a]] if self.is_float[a] else ri[self.int_row[a]] for a in range(self.na)]
        return keys, accs


_fast_cache = {}


_fast_by_id = {}


def _fast(nk, dts, accs) -> _FastCall:
    # callers pass the same (cached) dtype-code list and accumulator tuple for every chunk of a computation
    ik = (nk, id(dts), id(accs))
    hit = _fast_by_id.get(ik)
    if hit is not None and hit[0] is dts and hit[1] is accs:
        return hit[2]
    key = (nk, tuple(dts), tuple(accs))
    fc = _fast_cache.get(key)
    if fc is None:
        fc = _fast_cache[key] = _FastCall(nk, dts, accs)
    if len(_fast_by_id) > 1024:
        _fast_by_id.clear()
    _fast_by_id[ik] = (dts, accs, fc)
    return fc


def _check_col(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dim() == 1 and t.is_contiguous()
            and t.dtype in (torch.int64, torch.float64)):
        _req(t, what)   # raises the precise error
        raise MarsGBError(1, f"{what}: expected a contiguous 1-D CUDA column")


def preagg_dense(keys: Sequence[torch.Tensor], vals: Sequence[torch.Tensor], accs: Sequence[Tuple[int, int]]):
    """Map-side pre-aggregation of one chunk on the low-cardinality path (mgb_preagg_dense): one launch.
    Returns (key tensors, accumulator tensors) or None when the chunk is not a dense shape (too many groups
    for the on-chip accumulators) -- the caller then uses ``groupby_aggregate``."""
    global _dense_cap
    lib = _lib.load()
    if len(vals) > _lib.MGB_MAX_VALS or len(accs) > _lib.MGB_MAX_ACCS:
        return None
    for i, k in enumerate(keys):
        _check_col(k, "key column")
        if k.dtype != torch.int64:
            raise MarsGBError(3, f"key {i}: dtype {k.dtype} not supported (int64 keys)")
    for v in vals:
        _check_col(v, "value column")
    n = keys[0].numel()
    fc = _fast(len(keys), [F64 if v.dtype == torch.float64 else I64 for v in vals], accs)
    if _dense_cap is None:
        c = C.c_int64()
        check(lib.mgb_dense_capacity(C.byref(c)))
        _dense_cap = c.value
    for j, k in enumerate(keys):
        fc.kp[j] = k.data_ptr()
    for j, v in enumerate(vals):
        fc.vp[j] = v.data_ptr()
    bi, bf = fc.alloc(_dense_cap, keys[0].device)
    check(lib.mgb_preagg_dense(C.byref(fc.spec), fc.kp, fc.vp, n, fc.okp, fc.oap, _dense_cap, C.byref(fc.g), _stream()))
    G = fc.g.value
    if G < 0:
        return None
    _count(1 if n else 0)
    return fc.views(bi, bf, G)


def merge_small(pieces_keys: Sequence[Sequence[torch.Tensor]], pieces_accs: Sequence[Sequence[torch.Tensor]],
                acc_ops: Sequence[int], sort: bool = False):
    """Re-aggregation of small partial blocks without concatenating them (mgb_merge_small): one launch.
    pieces_keys[i] / pieces_accs[i]: key / accumulator columns of partial block i; acc_ops[a]: merge op
    (SUM for sums and counts, MIN, MAX).  Returns (keys, accs) or None if the input is not small."""
    lib = _lib.load()
    n_pieces = len(pieces_keys)
    rows = [int(pk[0].numel()) for pk in pieces_keys]
    total = sum(rows)
    na = len(acc_ops)
    if total > 2048 or n_pieces > 64 or na > _lib.MGB_MAX_ACCS or na > _lib.MGB_MAX_VALS:
        return None
    nk = len(pieces_keys[0])
    first = pieces_accs[0]
    dts = [F64 if t.dtype == torch.float64 else I64 for t in first]
    fc = _fast(nk, dts, tuple((a, op) for a, op in enumerate(acc_ops)))
    kt = (_lib._pp * max(n_pieces, 1))()
    at = (_lib._pp * max(n_pieces, 1))()
    keep = []
    for i, (pk, pa) in enumerate(zip(pieces_keys, pieces_accs)):
        karr = (C.c_void_p * nk)()
        aarr = (C.c_void_p * max(na, 1))()
        for j, t in enumerate(pk):
            _check_col(t, "piece key")
            karr[j] = t.data_ptr()
        for a, t in enumerate(pa):
            _check_col(t, "piece accumulator")
            if (t.dtype == torch.float64) != (dts[a] == F64):
                raise MarsGBError(1, "partial blocks differ in accumulator dtype")
            aarr[a] = t.data_ptr()
        keep.append((karr, aarr))
        kt[i] = C.cast(karr, _lib._pp)
        at[i] = C.cast(aarr, _lib._pp)
    bi, bf = fc.alloc(total, pieces_keys[0][0].device)
    rows_arr = (C.c_int64 * max(n_pieces, 1))(*rows)
    check(lib.mgb_merge_small(C.byref(fc.spec), n_pieces, rows_arr, kt, at, fc.okp, fc.oap, total, 1 if sort else 0,
                              C.byref(fc.g), _stream()))
    G = fc.g.value
    if G < 0:
        return None
    _count(1 if total else 0)
    return fc.views(bi, bf, G)


def preagg_dense_async(keys: Sequence[torch.Tensor], vals: Sequence[Optional[torch.Tensor]],
                       accs: Sequence[Tuple[int, int]], checked: bool = False, val_dtypes=None):
    """Like ``preagg_dense`` but without host synchronisation (mgb_preagg_dense_async): returns
    (int block, float block, fast-call) -- columns at capacity ``dense_capacity()``, the group count (or -1) in
    the device int64 at ``fc.count_ptr(bi)``.  None when the shape can never be dense (too many columns)."""
    global _dense_cap
    lib = _lib.load()
    if len(vals) > _lib.MGB_MAX_VALS or len(accs) > _lib.MGB_MAX_ACCS:
        return None
    if not checked:   # columns of frames built by this package were validated when the frame was made
        for i, k in enumerate(keys):
            _check_col(k, "key column")
            if k.dtype != torch.int64:
                raise MarsGBError(3, f"key {i}: dtype {k.dtype} not supported (int64 keys)")
        for v in vals:
            if v is not None:
                _check_col(v, "value column")
    n = keys[0].numel()
    # a value column may be None: an int64 column that is only counted is never read by the kernel (its dtype
    # then comes from val_dtypes); passing a null pointer for any other column is rejected by the library
    dts = val_dtypes if val_dtypes is not None else [F64 if v.dtype == torch.float64 else I64 for v in vals]
    fc = _fast(len(keys), dts, accs)
    if _dense_cap is None:
        c = C.c_int64()
        check(lib.mgb_dense_capacity(C.byref(c)))
        _dense_cap = c.value
    for j, k in enumerate(keys):
        fc.kp[j] = k.data_ptr()
    for j, v in enumerate(vals):
        fc.vp[j] = v.data_ptr() if v is not None else None
    bi, bf = fc.alloc(_dense_cap, keys[0].device)
    check(lib.mgb_preagg_dense_async(C.byref(fc.spec), fc.kp, fc.vp, n, fc.okp, fc.oap, _dense_cap, fc.count_ptr(bi),
                                     _stream()))
    _count(1 if n else 0)
    return bi, bf, fc


_batch_max = None


def dense_batch_max() -> int:
    global _batch_max
    if _batch_max is None:
        c = C.c_int32()
        check(_lib.load().mgb_dense_batch_max(C.byref(c)))
        _batch_max = c.value
    return _batch_max


_band_tables: dict = {}


def _preagg_batch(entry: str, cap: int, chunks, accs, val_dtypes, n_keys: int, device):
    lib = _lib.load()
    nv, n = len(val_dtypes), len(chunks)
    if nv > _lib.MGB_MAX_VALS or len(accs) > _lib.MGB_MAX_ACCS:
        return None
    if n > dense_batch_max():
        raise MarsGBError(1, f"{n} chunks in one batch (at most {dense_batch_max()})")
    fc = _fast(n_keys, val_dtypes, accs)
    # the chunk table of a band of RESIDENT chunks (cached entries, see executors._launch_band) is the same object list
    # call after call: its ctypes form is kept (bounded), keyed on the identity of the entries it was built from
    tkey = tuple(map(id, chunks)) if all(isinstance(e, list) and len(e) > 3 for e in chunks) else None
    hit = _band_tables.get(tkey) if tkey is not None else None
    if hit is not None and hit[0] == n_keys and hit[1] == nv:
        _, _, rows, kt, vt, keep, live, _ = hit
    else:
        rows = (C.c_int64 * max(n, 1))()
        kt = (_lib._pp * max(n, 1))()
        vt = (_lib._pp * max(n, 1))()
        keep = []
        live = 0
        for i, ent in enumerate(chunks):
            kp, vp, m = ent[0], ent[1], ent[2]
            if len(ent) > 3:                   # ctypes arrays built by an earlier call over the same resident chunk
                karr, varr = ent[3]
            else:
                karr = (C.c_void_p * n_keys)(*kp)
                varr = (C.c_void_p * max(nv, 1))(*vp)
                if isinstance(ent, list):
                    ent.append((karr, varr))
            keep.append((karr, varr))
            kt[i] = C.cast(karr, _lib._pp)
            vt[i] = C.cast(varr, _lib._pp)
            rows[i] = int(m)
            live += 1 if m else 0
        if tkey is not None:
            if len(_band_tables) >= 16:
                _band_tables.clear()
            _band_tables[tkey] = (n_keys, nv, rows, kt, vt, keep, live, list(chunks))   # the entries stay alive: ids stay valid
    bi, bf = fc.alloc(cap, device)
    check(getattr(lib, entry)(C.byref(fc.spec), n, rows, kt, vt, fc.okp, fc.oap, cap, fc.count_ptr(bi), _stream()))
    return bi, bf, fc, live


def preagg_dense_batch_async(chunks, accs: Sequence[Tuple[int, int]], val_dtypes, n_keys: int, device):
    """Band-level batch of the dense pre-aggregation (mgb_preagg_dense_batch_async): ONE launch streams all
    row batches and yields the partial rows of the whole band -- what n stage=map operands plus the combine
    tree yield.  ``chunks``: list of (key pointers, value pointers (None for a count-only int64 column), rows);
    at most ``dense_batch_max()`` of them.  Returns (int block, float block, fast-call) like
    ``preagg_dense_async``; None when the shape can never be dense."""
    global _dense_cap
    if _dense_cap is None:
        c = C.c_int64()
        check(_lib.load().mgb_dense_capacity(C.byref(c)))
        _dense_cap = c.value
    res = _preagg_batch("mgb_preagg_dense_batch_async", _dense_cap, chunks, accs, val_dtypes, n_keys, device)
    if res is None:
        return None
    if res[3]:
        n = C.c_int32()
        check(_lib.load().mgb_dense_last_launches(C.byref(n)))
        _count(n.value)     # 2 when the register-accumulator kernel (mgb_tiny.cu) went first, the dense one behind it
    return res[:3]


def make_row_program(n_phys: int, logical, row_filter=None):
    """ctypes row program (include/mgb.h: mgb_row_program).  ``logical``: per loaded logical column either None (the
    column is physical column of the same index) or a list of up to 3 factors (physical column, alpha, beta);
    ``row_filter``: None or (physical column, op code 0..5, dtype code, value)."""
    prog = _lib.RowProgram()
    prog.n_phys = n_phys
    for c, fac in enumerate(logical):
        prog.n_factors[c] = 0 if fac is None else len(fac)
        for f, (col, alpha, beta) in enumerate(fac or []):
            prog.factors[c][f].col, prog.factors[c][f].alpha, prog.factors[c][f].beta = col, float(alpha), float(beta)
    if row_filter is None:
        prog.filter_col = -1
    else:
        col, op, dt, val = row_filter
        prog.filter_col, prog.filter_op, prog.filter_dtype = col, op, dt
        prog.filter_i64 = int(val) if dt == I64 else 0
        prog.filter_f64 = float(val)
    return prog


def preagg_dense_batch_fused_async(chunks, accs, val_dtypes, n_keys: int, device, prog):
    """Band-level batch with the fused pre-path (mgb_preagg_dense_batch_fused_async): ``chunks`` carry the
    PHYSICAL value columns of the row program ``prog``; the spec (accs / val_dtypes) describes the logical ones."""
    global _dense_cap
    if _dense_cap is None:
        c = C.c_int64()
        check(_lib.load().mgb_dense_capacity(C.byref(c)))
        _dense_cap = c.value
    lib = _lib.load()
    n = len(chunks)
    if n > dense_batch_max():
        raise MarsGBError(1, f"{n} chunks in one batch (at most {dense_batch_max()})")
    fc = _fast(n_keys, val_dtypes, accs)
    rows = (C.c_int64 * max(n, 1))()
    kt = (_lib._pp * max(n, 1))()
    vt = (_lib._pp * max(n, 1))()
    keep = []
    for i, (kp, vp, m) in enumerate(chunks):
        karr = (C.c_void_p * n_keys)(*kp)
        varr = (C.c_void_p * max(len(vp), 1))(*vp)
        keep.append((karr, varr))
        kt[i] = C.cast(karr, _lib._pp)
        vt[i] = C.cast(varr, _lib._pp)
        rows[i] = int(m)
    bi, bf = fc.alloc(_dense_cap, device)
    check(lib.mgb_preagg_dense_batch_fused_async(C.byref(fc.spec), C.byref(prog), n, rows, kt, vt, fc.okp, fc.oap,
                                                 _dense_cap, fc.count_ptr(bi), _stream()))
    if any(m for _, _, m in chunks):
        nl = C.c_int32()
        check(lib.mgb_dense_last_launches(C.byref(nl)))
        _count(nl.value)
    return bi, bf, fc


def eval_product(phys: Sequence[torch.Tensor], factors) -> torch.Tensor:
    """materialises one derived column: product of (alpha + beta * phys[col]) terms (mgb_eval_product)"""
    lib = _lib.load()
    phys = [_req(t, f"column {i}", (torch.float64,)) for i, t in enumerate(phys)]
    n = phys[0].numel()
    out = torch.empty(n, dtype=torch.float64, device=phys[0].device)
    fa = (_lib.AffineFactor * max(len(factors), 1))()
    for f, (col, alpha, beta) in enumerate(factors):
        fa[f].col, fa[f].alpha, fa[f].beta = col, float(alpha), float(beta)
    check(lib.mgb_eval_product(_ptrs(phys), fa, len(factors), out.data_ptr(), n, _stream()))
    _count(1 if n else 0)
    return out


def filter_rows(filter_col: torch.Tensor, op: int, value, cols: Sequence[torch.Tensor]):
    """stable compaction of ``cols`` to the rows where ``filter_col <op> value`` (mgb_filter_rows; one host wait)"""
    lib = _lib.load()
    fcol = _req(filter_col, "filter column")
    cols = [_req(c, f"column {i}") for i, c in enumerate(cols)]
    n = fcol.numel()
    outs = [torch.empty_like(c) for c in cols]
    k = C.c_int64()
    dt = dtype_code(fcol)
    check(lib.mgb_filter_rows(fcol.data_ptr(), dt, int(op), int(value) if dt == I64 else 0, float(value), _ptrs(cols),
                              _ptrs(outs), len(cols), n, C.byref(k), _stream()))
    _count(8 if n else 0)
    return [o[:k.value] for o in outs], k.value


_mid_cap = None


def preagg_mid_batch_async(chunks, accs: Sequence[Tuple[int, int]], val_dtypes, n_keys: int, device):
    """The same for the mid-cardinality sum / mean / count shapes (mgb_preagg_mid_batch_async): up to
    mgb_mid_capacity() groups in the band; the device count is -1 for anything the path does not take (other
    functions, too many groups, NaN values)."""
    global _mid_cap
    if _mid_cap is None:
        c = C.c_int64()
        check(_lib.load().mgb_mid_capacity(C.byref(c)))
        _mid_cap = c.value
    res = _preagg_batch("mgb_preagg_mid_batch_async", _mid_cap, chunks, accs, val_dtypes, n_keys, device)
    if res is None:
        return None
    _count(2 if res[3] else 0)
    return res[:3]


SMALL_MAX_ROWS = 2048
SMALL_MAX_PIECES = 16


def merge_small_async(pieces, acc_ops: Sequence[int], sort: bool = False, checked: bool = False):
    """Re-aggregation of small partial blocks without host synchronisation (mgb_merge_small_async).
    ``pieces``: list of (key tensors, accumulator tensors, n_rows or None, device count pointer or 0); a piece
    with n_rows None has its size in the device int64 at the given pointer.  Returns (int block, float block,
    fast-call) at capacity SMALL_MAX_ROWS with the group count (or -1) at ``fc.count_ptr(bi)``; None when the
    host can already tell that the input is not small."""
    lib = _lib.load()
    n_pieces = len(pieces)
    na = len(acc_ops)
    known = sum(p[2] for p in pieces if p[2] is not None)
    live = sum(1 for p in pieces if p[2] is None or p[2] > 0)
    if known > SMALL_MAX_ROWS or live > SMALL_MAX_PIECES or na > _lib.MGB_MAX_ACCS or na > _lib.MGB_MAX_VALS:
        return None
    nk = len(pieces[0][0])
    by_ptr = len(pieces[0]) == 6        # (key pointers, accumulator pointers, n, count pointer, dtype codes, device)
    if by_ptr:
        dts = pieces[0][4]
        out_device = pieces[0][5]
    else:
        dts = [F64 if t.dtype == torch.float64 else I64 for t in pieces[0][1]]
        out_device = pieces[0][0][0].device
    fc = _fast(nk, dts, tuple((a, op) for a, op in enumerate(acc_ops)))
    kt = (_lib._pp * max(n_pieces, 1))()
    at = (_lib._pp * max(n_pieces, 1))()
    rows_arr = (C.c_int64 * max(n_pieces, 1))()
    rows_dev = (C.c_void_p * max(n_pieces, 1))()
    keep = []
    for i, piece in enumerate(pieces):
        pk, pa, n, cptr = piece[0], piece[1], piece[2], piece[3]
        if by_ptr:
            if piece[4] != dts and list(piece[4]) != list(dts):
                raise MarsGBError(1, "partial blocks differ in accumulator dtype")
            karr = (C.c_void_p * nk)(*pk)
            aarr = (C.c_void_p * max(na, 1))(*pa)
            keep.append((karr, aarr))
            kt[i] = C.cast(karr, _lib._pp)
            at[i] = C.cast(aarr, _lib._pp)
            rows_arr[i] = -1 if n is None else int(n)
            rows_dev[i] = cptr if n is None else None
            continue
        karr = (C.c_void_p * nk)()
        aarr = (C.c_void_p * max(na, 1))()
        for j, t in enumerate(pk):
            if not checked:
                _check_col(t, "piece key")
            karr[j] = t.data_ptr()
        for a, t in enumerate(pa):
            if not checked:
                _check_col(t, "piece accumulator")
            if (t.dtype == torch.float64) != (dts[a] == F64):
                raise MarsGBError(1, "partial blocks differ in accumulator dtype")
            aarr[a] = t.data_ptr()
        keep.append((karr, aarr))
        kt[i] = C.cast(karr, _lib._pp)
        at[i] = C.cast(aarr, _lib._pp)
        rows_arr[i] = -1 if n is None else int(n)
        rows_dev[i] = cptr if n is None else None
    bi, bf = fc.alloc(SMALL_MAX_ROWS, out_device)
    check(lib.mgb_merge_small_async(C.byref(fc.spec), n_pieces, rows_arr, rows_dev, kt, at, fc.okp, fc.oap,
                                    SMALL_MAX_ROWS, 1 if sort else 0, fc.count_ptr(bi), _stream()))
    _count(1)
    return bi, bf, fc


class PostPlan:
    """ctypes side of the post steps of one stage=agg call shape (cached by the executors per operand shape)"""
    __slots__ = ("steps", "n", "is_float")

    def __init__(self, steps, is_float):
        """steps: list of (kind, a0, a1, a2, ddof); is_float[j]: result column j is float64 (else int64)"""
        self.n = len(steps)
        if self.n > _lib.MGB_MAX_POST:
            raise MarsGBError(3, f"{self.n} result columns (at most {_lib.MGB_MAX_POST} in one launch)")
        self.steps = (_lib.PostStep * max(self.n, 1))()
        for j, (kind, a0, a1, a2, ddof) in enumerate(steps):
            st = self.steps[j]
            st.kind, st.a0, st.a1, st.a2, st.ddof = kind, a0, a1, a2, ddof
        self.is_float = list(is_float)


def merge_small_post_async(pieces, acc_ops: Sequence[int], post: PostPlan, sort: bool = False):
    """stage=agg in one launch (mgb_merge_small_post_async): merge of small partial blocks + the post functions.
    ``pieces`` as for ``merge_small_async`` in pointer form: (key pointers, accumulator pointers, n or None,
    count pointer, dtype codes, device).  Returns (result block, n_keys, n_post) -- an int64 tensor
    [1 + n_keys + n_post (+ scratch), SMALL_MAX_ROWS]: row 0 holds the row count (or -1) in its first element,
    rows 1..n_keys the group keys, then the result columns (float64 bit patterns where ``post.is_float``) --
    or None when the host can already tell that the input is not small."""
    lib = _lib.load()
    n_pieces = len(pieces)
    na = len(acc_ops)
    known = sum(p[2] for p in pieces if p[2] is not None)
    live = sum(1 for p in pieces if p[2] is None or p[2] > 0)
    if known > SMALL_MAX_ROWS or live > SMALL_MAX_PIECES or na > _lib.MGB_MAX_ACCS or na > _lib.MGB_MAX_VALS:
        return None
    nk = len(pieces[0][0])
    dts = pieces[0][4]
    device = pieces[0][5]
    fc = _fast(nk, dts, tuple((a, op) for a, op in enumerate(acc_ops)))
    kt = (_lib._pp * max(n_pieces, 1))()
    at = (_lib._pp * max(n_pieces, 1))()
    rows_arr = (C.c_int64 * max(n_pieces, 1))()
    rows_dev = (C.c_void_p * max(n_pieces, 1))()
    keep = []
    for i, piece in enumerate(pieces):
        pk, pa, n, cptr = piece[0], piece[1], piece[2], piece[3]
        if piece[4] != dts and list(piece[4]) != list(dts):
            raise MarsGBError(1, "partial blocks differ in accumulator dtype")
        karr = (C.c_void_p * nk)(*pk)
        aarr = (C.c_void_p * max(na, 1))(*pa)
        keep.append((karr, aarr))
        kt[i] = C.cast(karr, _lib._pp)
        at[i] = C.cast(aarr, _lib._pp)
        rows_arr[i] = -1 if n is None else int(n)
        rows_dev[i] = cptr if n is None else None
    cap = SMALL_MAX_ROWS
    n_res = 1 + nk + post.n
    blk = torch.empty((n_res + na, cap), dtype=torch.int64, device=device)
    base, stride = blk.data_ptr(), cap * 8
    okp = (C.c_void_p * nk)(*[base + (1 + j) * stride for j in range(nk)])
    opp = (C.c_void_p * max(post.n, 1))(*[base + (1 + nk + j) * stride for j in range(post.n)])
    oap = (C.c_void_p * max(na, 1))(*[base + (n_res + a) * stride for a in range(na)])
    check(lib.mgb_merge_small_post_async(C.byref(fc.spec), n_pieces, rows_arr, rows_dev, kt, at, okp, oap, cap,
                                         1 if sort else 0, post.steps, post.n, opp, base, _stream()))
    _count(1)
    return blk, nk, post.n


_pinned = {}


def _pinned_block(nbytes: int) -> torch.Tensor:
    """re-used pinned staging block of the current thread (results are copied out of it before it is re-used)"""
    import threading

    key = threading.get_ident()
    buf = _pinned.get(key)
    if buf is None or buf.numel() * 8 < nbytes:
        words = max(1 << 15, 1 << (max(nbytes - 1, 1).bit_length() - 3 + 1))
        buf = _pinned[key] = torch.empty(words, dtype=torch.int64, pin_memory=True)
    return buf


def result_block_to_host(blk: torch.Tensor, nk: int, n_post: int, is_float: Sequence[bool]):
    """One device -> host copy of the result part of a ``merge_small_post_async`` block (count row, keys, result
    columns) into a re-used pinned block, one synchronisation.  Returns (n, key arrays, column arrays) cut to
    n rows -- n < 0: the speculative chain failed somewhere (arrays are None)."""
    import numpy as np

    cap = blk.shape[1]
    rows = 1 + nk + n_post
    host = _pinned_block(rows * cap * 8)[:rows * cap].view(rows, cap)
    host.copy_(blk[:rows], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    n = int(host[0, 0])
    if n < 0:
        return n, None, None
    hn = host.numpy()
    keys = [hn[1 + j, :n].copy() for j in range(nk)]
    cols = [hn[1 + nk + j, :n].view(np.float64).copy() if is_float[j] else hn[1 + nk + j, :n].copy()
            for j in range(n_post)]
    return n, keys, cols


def to_host_counted(columns: Sequence[torch.Tensor], count_block: torch.Tensor, count_row: int):
    """Device -> host read of capacity-length columns TOGETHER with their device-side row count: all copies
    and the count go into one pinned block, one synchronisation.  Returns (n, numpy arrays cut to n rows) --
    n < 0 means the speculative result is invalid (arrays are then None)."""
    import numpy as np

    cap = int(columns[0].numel()) if columns else 1
    # columns that are rows of one 2-D block (the output blocks of the merge) travel as ONE copy of the block
    groups = {}      # id(base) -> (base, host row offset)
    where = []       # per column: (host row, is_float)
    rows = 0
    for c in columns:
        base = c._base
        if base is not None and base.dim() == 2 and base.is_contiguous() and base.shape[1] == cap and c.stride(0) == 1:
            ent = groups.get(id(base))
            if ent is None:
                ent = groups[id(base)] = (base, rows)
                rows += base.shape[0]
            where.append(ent[1] + (c.data_ptr() - base.data_ptr()) // (cap * 8))
        else:
            where.append(rows)
            groups[id(c)] = (c, rows)
            rows += 1
    crow = rows
    host = torch.empty((rows + 1, max(cap, 1)), dtype=torch.int64, pin_memory=True)
    for t, r0 in groups.values():
        if t.dim() == 2:
            host[r0:r0 + t.shape[0], :cap].view(t.dtype).copy_(t, non_blocking=True)
        else:
            host[r0, :cap].view(t.dtype).copy_(t, non_blocking=True)
    host[crow, :1].copy_(count_block[count_row, :1], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    n = int(host[crow, 0])
    if n < 0:
        return n, None
    hn = host.numpy()
    out = []
    for r, c in zip(where, columns):
        arr = hn[r, :n]
        out.append(arr.view(np.float64).copy() if c.dtype == torch.float64 else arr.copy())
    return n, out


# Result frames wrap the arrays that came off the device without a copy, so they keep the PINNED block alive for as long
# as the user holds the frame.  That is the point for the usual (small to medium) result; a block above this many bytes
# is copied into pageable memory instead, so that a multi-GB result does not page-lock that much host RAM indefinitely.
PINNED_RESULT_MAX = int(os.environ.get("MGB_PINNED_RESULT_MAX", str(1 << 30)))


def _unpin_large(arrays, host_block: torch.Tensor):
    if host_block.numel() * host_block.element_size() <= PINNED_RESULT_MAX:
        return arrays
    return [a.copy() for a in arrays]


def to_host(columns: Sequence[torch.Tensor]):
    """Device -> host read of several equally long 8-byte columns with one synchronisation: asynchronous
    copies into one pinned block, then a single stream sync.  Returns numpy arrays (copies)."""
    import numpy as np

    if not columns:
        return []
    n = int(columns[0].numel())
    host = torch.empty((len(columns), max(n, 1)), dtype=torch.int64, pin_memory=True)
    for j, c in enumerate(columns):
        c = _req(c, f"column {j}")
        host[j, :n].view(c.dtype).copy_(c, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    # the arrays are views of the pinned block (which they keep alive): no second pass over the result
    hn = host.numpy()
    return _unpin_large([hn[j, :n].view(np.float64) if c.dtype == torch.float64 else hn[j, :n] for j, c in enumerate(columns)],
                        host)


def to_host_concat(parts: Sequence[Sequence[torch.Tensor]]):
    """Device -> host read of several row blocks that share their columns, concatenated row-wise on the way:
    parts[i][j] is column j of block i.  One pinned block [columns, total rows], one asynchronous copy per
    (block, column) straight to its final place, one synchronisation.  Returns one numpy array per column
    (views of the pinned block)."""
    import numpy as np

    ncol = len(parts[0])
    sizes = [int(p[0].numel()) if ncol else 0 for p in parts]
    total = sum(sizes)
    host = torch.empty((max(ncol, 1), max(total, 1)), dtype=torch.int64, pin_memory=True)
    off = 0
    for p, n in zip(parts, sizes):
        for j, c in enumerate(p):
            c = _req(c, f"column {j}")
            if c.dtype != parts[0][j].dtype or int(c.numel()) != n:
                raise MarsGBError(1, "result blocks differ in column dtype or length")
            host[j, off:off + n].view(c.dtype).copy_(c, non_blocking=True)
        off += n
    torch.cuda.current_stream().synchronize()
    hn = host.numpy()
    out = [hn[j, :total].view(np.float64) if c.dtype == torch.float64 else hn[j, :total] for j, c in enumerate(parts[0])]
    return _unpin_large(out, host)


# ------------------------------------------------------------------------------------------------- post
def row_partials(vals: Sequence[torch.Tensor], accs: Sequence[Tuple[int, int]]) -> List[torch.Tensor]:
    """per-row partials of a chunk whose keys are (nearly) all distinct (mgb_row_partial): SUM / MIN / MAX alias the
    value column, COUNT and SUMSQ are written (one elementwise launch each)."""
    lib = _lib.load()
    out = []
    for col, op in accs:
        v = _req(vals[col], f"value column {col}")
        if op in (SUM, MIN, MAX):
            out.append(v)
            continue
        o = torch.empty(v.numel(), dtype=torch.int64 if op == COUNT else v.dtype, device=v.device)
        check(lib.mgb_row_partial(v.data_ptr(), dtype_code(v), int(op), o.data_ptr(), v.numel(), _stream()))
        _count(1 if v.numel() else 0)
        out.append(o)
    return out


def post_mean(s: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    s, c = _req(s, "sum"), _req(c, "count", (torch.int64,))
    out = torch.empty(s.numel(), dtype=torch.float64, device=s.device)
    check(lib.mgb_post_mean(s.data_ptr(), dtype_code(s), c.data_ptr(), out.data_ptr(), s.numel(), _stream()))
    _count(1 if s.numel() else 0)
    return out


def post_var(s: torch.Tensor, q: torch.Tensor, c: torch.Tensor, ddof: int = 1) -> torch.Tensor:
    lib = _lib.load()
    s, q, c = _req(s, "sum"), _req(q, "sum of squares"), _req(c, "count", (torch.int64,))
    if s.dtype != q.dtype:
        raise MarsGBError(1, "sum and sum-of-squares must share a dtype")
    out = torch.empty(s.numel(), dtype=torch.float64, device=s.device)
    check(lib.mgb_post_var(s.data_ptr(), q.data_ptr(), dtype_code(s), c.data_ptr(), int(ddof), out.data_ptr(),
                           s.numel(), _stream()))
    _count(1 if s.numel() else 0)
    return out


# ----------------------------------------------------------------------------------------- concat / sort
def concat_columns(pieces: Sequence[torch.Tensor]) -> torch.Tensor:
    """Row-wise concatenation of one column's pieces with device-to-device copies (mgb_copy_rows)."""
    lib = _lib.load()
    pieces = [_req(p, "piece") for p in pieces]
    if len(pieces) == 1:
        return pieces[0]
    total = sum(p.numel() for p in pieces)
    out = torch.empty(total, dtype=pieces[0].dtype, device=pieces[0].device)
    off = 0
    for p in pieces:
        if p.dtype != out.dtype:
            raise MarsGBError(1, "concat: pieces differ in dtype")
        check(lib.mgb_copy_rows(out.data_ptr(), off, p.data_ptr(), p.numel(), _stream()))
        off += p.numel()
    return out


def copy_segments(segments) -> None:
    """segments: list of (src tensor, dst) of 8-byte elements, dst a tensor of equal length or a raw device
    address: all copied by one launch per 96 segments (mgb_copy_segments)."""
    lib = _lib.load()
    segs = [(a, b) for a, b in segments if a.numel()]
    if not segs:
        return
    n = len(segs)
    src = (C.c_void_p * n)(*[a.data_ptr() for a, _ in segs])
    dst = (C.c_void_p * n)(*[b if isinstance(b, int) else b.data_ptr() for _, b in segs])
    lens = (C.c_int64 * n)(*[a.numel() for a, _ in segs])
    check(lib.mgb_copy_segments(src, dst, lens, n, _stream()))
    _count((n + 95) // 96)


def argsort_keys(keys: Sequence[torch.Tensor]) -> torch.Tensor:
    lib = _lib.load()
    keys = [_req(k, f"key {i}", (torch.int64,)) for i, k in enumerate(keys)]
    n = keys[0].numel()
    perm = torch.empty(n, dtype=torch.int32, device=keys[0].device)
    check(lib.mgb_argsort_keys(_ptrs(keys), len(keys), n, perm.data_ptr(), _stream()))
    _count(4 * 8 * len(keys) if n else 0)
    return perm


def gather_rows(cols: Sequence[torch.Tensor], perm: torch.Tensor) -> List[torch.Tensor]:
    lib = _lib.load()
    cols = [_req(c, f"column {i}") for i, c in enumerate(cols)]
    perm = _req(perm, "perm", (torch.int32,))
    n = perm.numel()
    outs = [torch.empty(n, dtype=c.dtype, device=c.device) for c in cols]
    check(lib.mgb_gather_rows(_ptrs(cols), _ptrs(outs), len(cols), perm.data_ptr(), n, _stream()))
    _count(1 if n and cols else 0)
    return outs
