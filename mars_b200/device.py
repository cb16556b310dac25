"""Device-resident frames: what flows through ``ctx`` between the GPU executors.

The reference moves pandas (or cudf) DataFrames between operands; on a GPU band the partial results of the
groupby path never leave HBM, so the values placed in ``ctx`` are ``DeviceFrame`` objects: a list of int64
index (group key) columns plus named float64/int64 data columns, each a 1-D CUDA tensor.  A tuple of K
DeviceFrames sharing their index tensors replaces the reference's tuple of K DataFrames
(mars/dataframe/groupby/aggregation.py:1128).  Frames are freed by Python reference counting, exactly like
the processor frees ctx entries (processor.py:276-282).
"""
from __future__ import annotations

import contextlib
import os
import threading
import weakref
from typing import Any, List, Optional, Sequence

import numpy as np
import pandas as pd
import torch


h2d_bytes = 0   # bytes copied host -> device by this module (bench.py reports them per step)


class HostColumn:
    """A column of a chunk that is still in host memory: uploaded by DeviceFrame.col(i) on first use.  A column
    no kernel reads (e.g. an int64 column that is only counted) never crosses PCIe."""
    __slots__ = ("array",)

    def __init__(self, array: np.ndarray):
        self.array = array

    @property
    def dtype(self):
        return torch.float64 if self.array.dtype == np.float64 else torch.int64

    def numel(self):
        return int(self.array.shape[0])

    def element_size(self):
        return 8


class DerivedColumn:
    """A column of a chunk that is DEFINED, not stored: the left-to-right product of up to three affine factors
    ``alpha + beta * <data column>`` of the same frame (``price * (1 - discount) * (1 + tax)``).  The band-level
    batch kernel evaluates it inside the pre-aggregation pass (kernels.make_row_program); any other consumer gets
    it materialised by ``DeviceFrame.col`` (one elementwise kernel)."""
    __slots__ = ("factors",)
    dtype = torch.float64

    def __init__(self, factors):
        self.factors = [(c, float(a), float(b)) for c, a, b in factors]

    def element_size(self):
        return 8

    def numel(self):
        return 0        # not stored


FILTER_OPS = {"<=": 0, "<": 1, ">=": 2, ">": 3, "==": 4, "!=": 5}

# set by mars_b200.executors (the kernels module it dispatches to): materialisation of lazy parts
_lazy_kernels = None


def _usable_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))         # what nproc reports (the container's share), not the machine's cores
    except (AttributeError, OSError):
        return os.cpu_count() or 8


class _StagingRing:
    """Pinned staging for uploads from PAGEABLE host memory (what a Mars chunk store hands over: pandas blocks).
    cudaMemcpy from pageable memory is staged by the driver on one thread (~10 GB/s measured); here worker threads copy
    slab-sized pieces of the column into a ring of pinned slabs (numpy releases the GIL for the copy) while the copy
    engine moves the slabs already filled -- the H2D of one piece overlaps the host copy of the next ones, and the
    call returns as soon as the last piece is enqueued."""

    SLAB_BYTES = int(os.environ.get("MGB_STAGE_SLAB_MB", "8")) << 20
    SLABS = int(os.environ.get("MGB_STAGE_SLABS", "8"))
    THREADS = int(os.environ.get("MGB_STAGE_THREADS", "4"))
    # host threads of the narrowing pass: the ranks of a node share its cores
    NARROW_THREADS = int(os.environ.get("MGB_NARROW_THREADS", "0")) or max(
        2, min(8, _usable_cores() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))))

    def __init__(self):
        self.slabs = None
        self.events = None
        self.pool = None
        self.next = 0

    def _init(self):
        from concurrent.futures import ThreadPoolExecutor

        self.slabs = [torch.empty(self.SLAB_BYTES // 8, dtype=torch.int64).pin_memory() for _ in range(self.SLABS)]
        self.views = [t.numpy() for t in self.slabs]
        self.events = [None] * self.SLABS
        self.pool = ThreadPoolExecutor(self.THREADS, thread_name_prefix="mgb-stage")
        self.npool = ThreadPoolExecutor(self.NARROW_THREADS, thread_name_prefix="mgb-narrow")

    def upload(self, arr: np.ndarray, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.slabs is None:
            self._init()
        n = arr.shape[0]
        src = arr.view(np.int64)
        if out is None:
            out = torch.empty(n, dtype=torch.float64 if arr.dtype == np.float64 else torch.int64, device=device)
        out_i = out.view(torch.int64)
        per = self.SLAB_BYTES // 8
        pending = []          # (future, slab index, start, end) in submission order

        def drain(limit):
            while len(pending) > limit:
                fut, si, a, b = pending.pop(0)
                fut.result()
                out_i[a:b].copy_(self.slabs[si][:b - a], non_blocking=True)
                ev = self.events[si] or torch.cuda.Event()
                ev.record()
                self.events[si] = ev

        for a in range(0, n, per):
            b = min(a + per, n)
            si = self.next
            self.next = (self.next + 1) % self.SLABS
            if any(p[1] == si for p in pending):       # the ring went round: that slab's piece must leave first
                drain(0)
            ev = self.events[si]
            if ev is not None:
                ev.synchronize()                        # its previous H2D has read the slab
            pending.append((self.pool.submit(np.copyto, self.views[si][:b - a], src[a:b]), si, a, b))
            drain(self.SLABS - 2)
        drain(0)
        return out

    # ---- int64 columns whose values fit 1 / 2 / 4 bytes cross PCIe narrow (mgb_host_narrow_i64 / mgb_widen_i64)
    def upload_narrow(self, arr: np.ndarray, device, width: int) -> Optional[torch.Tensor]:
        """Worker threads rewrite slab-sized pieces of the int64 (or integer-valued float64) column as ``width``-byte
        integers straight into the pinned slabs (every value checked: lossless or refused), the copy engine moves the
        narrow bytes, one kernel restores the 8-byte column in HBM.  Works from pinned and pageable sources alike (the
        host pass reads the source once and writes 1/8 .. 1/2 of it).  None when a value does not fit: nothing useful
        was enqueued, the caller uploads the 8-byte column."""
        import ctypes as C

        from . import _lib

        if self.slabs is None:
            self._init()
        lib = _lib.load()
        n = int(arr.shape[0])
        narrow = torch.empty(n * width, dtype=torch.uint8, device=device)
        per = (self.SLAB_BYTES // width) & ~15                  # elements per slab
        T = max(1, min(self.NARROW_THREADS, arr.nbytes >> 21))     # >= 2 MB of source per thread
        src0 = arr.ctypes.data

        is_float = arr.dtype == np.float64
        host_narrow = lib.mgb_host_narrow_f64 if is_float else lib.mgb_host_narrow_i64

        def job(a, b, dst):
            fits = C.c_int32(0)
            _lib.check(host_narrow(src0 + a * 8, b - a, width, dst, C.byref(fits)))
            return fits.value

        for a in range(0, n, per):
            b = min(a + per, n)
            si = self.next
            self.next = (self.next + 1) % self.SLABS
            ev = self.events[si]
            if ev is not None:
                ev.synchronize()                                # its previous H2D has read the slab
            base = self.slabs[si].data_ptr()
            cuts = [a + (b - a) * t // T for t in range(T + 1)]
            futs = [self.npool.submit(job, cuts[t], cuts[t + 1], base + (cuts[t] - a) * width)
                    for t in range(1, T) if cuts[t + 1] > cuts[t]]
            ok = job(cuts[0], cuts[1], base)                    # the calling thread takes the first part
            if not all([f.result() for f in futs]) or not ok:
                return None
            narrow[a * width:b * width].copy_(self.slabs[si].view(torch.uint8)[:(b - a) * width], non_blocking=True)
            ev = ev or torch.cuda.Event()
            ev.record()
            self.events[si] = ev
        out = torch.empty(n, dtype=torch.float64 if is_float else torch.int64, device=device)
        widen = lib.mgb_widen_f64 if is_float else lib.mgb_widen_i64
        _lib.check(widen(narrow.data_ptr(), width, n, out.data_ptr(), torch.cuda.current_stream(device).cuda_stream))
        return out


    def submit_narrow(self, arr: np.ndarray, device, width: int, label) -> Optional[torch.Tensor]:
        """Deferred form of ``upload_narrow`` (inside ``deferred_uploads()``): ONE worker narrows the whole column
        into a pinned buffer of its own while the caller goes on enqueueing the copies of the other columns; the
        returned device column is filled when the block ends (``_flush_deferred``: narrow H2D + widen, or the 8-byte
        copy if a value did not fit).  None when the pinned pool is exhausted (the caller narrows synchronously)."""
        import ctypes as C

        from . import _lib

        if self.slabs is None:
            self._init()
        n = int(arr.shape[0])
        buf = _pinned.acquire(n * width)
        if buf is None:
            return None
        lib = _lib.load()
        is_float = arr.dtype == np.float64
        host_narrow = lib.mgb_host_narrow_f64 if is_float else lib.mgb_host_narrow_i64
        src, dst = arr.ctypes.data, buf.data_ptr()

        def job():
            fits = C.c_int32(0)
            _lib.check(host_narrow(src, n, width, dst, C.byref(fits)))
            return fits.value

        out = torch.empty(n, dtype=torch.float64 if is_float else torch.int64, device=device)
        _tls.pending.append((self.npool.submit(job), buf, out, arr, width, label, device))
        return out


class _PinnedPool:
    """Pinned buffers of the deferred narrowing pass, by power-of-two size class; a buffer returns with the event
    behind its last H2D and is handed out again only after that event."""
    CAP = int(os.environ.get("MGB_NARROW_PINNED_MB", "1024")) << 20

    def __init__(self):
        self.free = {}
        self.total = 0
        self.lock = threading.Lock()

    def acquire(self, nbytes: int) -> Optional[torch.Tensor]:
        cls = max(1 << 20, 1 << (max(nbytes, 1) - 1).bit_length())
        with self.lock:
            lst = self.free.get(cls)
            hit = lst.pop(0) if lst else None
            if hit is None:
                if self.total + cls > self.CAP:
                    return None
                self.total += cls
        if hit is None:
            return torch.empty(cls, dtype=torch.uint8).pin_memory()
        t, ev = hit
        if ev is not None:
            ev.synchronize()
        return t

    def release(self, t: torch.Tensor, ev) -> None:
        with self.lock:
            self.free.setdefault(int(t.numel()), []).append((t, ev))


_pinned = _PinnedPool()
_tls = threading.local()      # .pending: (future, pinned buffer, device column, ...) of the deferred_uploads() block this
                              # thread is in -- another thread's uploads must not land in it (its consumer would not wait)


@contextlib.contextmanager
def deferred_uploads():
    """Uploads requested inside the block may complete when it ends: columns that cross PCIe narrow are narrowed by
    worker threads in the background while the caller enqueues the copies of the other columns.  ONLY for blocks that
    collect columns (pointers) and launch nothing that reads them -- the consumer is enqueued after the block."""
    if getattr(_tls, "pending", None) is not None or not (NARROW and NARROW_DEFER):      # nested: the outer block flushes
        yield
        return
    _tls.pending = []
    try:
        yield
    finally:
        pend, _tls.pending = _tls.pending, None
        _flush_deferred(pend)


def _upload_into(out: torch.Tensor, arr: np.ndarray, device) -> None:
    """the 8-byte column into an existing device column (pinned: one asynchronous copy; pageable: the slab ring)"""
    t = _host_tensor(arr)
    if arr.nbytes >= STAGE_MIN_BYTES and not t.is_pinned():
        _staging.upload(t.numpy(), device, out=out)
    else:
        out.copy_(t, non_blocking=True)


def _flush_deferred(pend) -> None:
    """End of a deferred block: per column, in submission order, the narrow copy + widen kernel -- or the 8-byte copy
    when a value did not fit.  The copy engine must not wait for the host: when the stream has run dry and a column's
    pass has not even started (``NARROW_OPPORTUNISTIC``), the job is cancelled and the column crosses at full width,
    so the share of narrow columns follows what the host cores and its memory can keep up with (8 ranks of a node
    share them)."""
    global h2d_bytes, narrowed_columns, deferred_columns, skipped_columns
    import concurrent.futures as cf

    from . import _lib

    if not pend:
        return
    lib = _lib.load()
    used, err = [], None
    for fut, buf, out, arr, width, label, device in pend:
        used.append(buf)
        n = int(out.numel())
        fits = None
        try:
            while True:
                try:
                    fits = fut.result(timeout=None if not NARROW_OPPORTUNISTIC else 2e-4)
                    break
                except cf.TimeoutError:
                    if torch.cuda.current_stream(device).query() and fut.cancel():
                        break
        except Exception as e:  # noqa: BLE001  (every future is waited for: the workers read ``arr`` and write ``buf``)
            err = err or e
            continue
        if fits is None:                      # cancelled before it started
            _upload_into(out, arr, device)
            h2d_bytes += n * 8
            skipped_columns += 1
        elif fits:
            narrow = torch.empty(n * width, dtype=torch.uint8, device=device)
            narrow.copy_(buf[:n * width], non_blocking=True)
            widen = lib.mgb_widen_f64 if out.dtype == torch.float64 else lib.mgb_widen_i64
            _lib.check(widen(narrow.data_ptr(), width, n, out.data_ptr(), torch.cuda.current_stream(device).cuda_stream))
            h2d_bytes += n * width
            narrowed_columns += 1
            deferred_columns += 1
            if label is not None:
                _narrow_hint[label] = width
        else:
            _upload_into(out, arr, device)
            h2d_bytes += n * 8
            _narrow_failed(label, width)
    ev = torch.cuda.Event()
    ev.record()
    for buf in used:
        _pinned.release(buf, ev)
    if err is not None:
        raise err


_staging = _StagingRing()
STAGE_MIN_BYTES = 1 << 20     # smaller pageable columns go through the driver's own staging
NARROW = os.environ.get("MGB_NARROW", "1") != "0"
NARROW_MIN_BYTES = 4 << 20    # below this the host pass and its thread hand-off cost more than the PCIe bytes saved
NARROW_FLOATS = os.environ.get("MGB_NARROW_FLOATS", "1") != "0"   # float64 columns that hold small integers, too
NARROW_DEFER = os.environ.get("MGB_NARROW_DEFER", "1") != "0"     # band launches narrow in the background (deferred_uploads)
NARROW_OPPORTUNISTIC = os.environ.get("MGB_NARROW_OPPORTUNISTIC", "1") != "0"   # never let the copy engine wait for a host pass
_no_narrow: set = set()       # column labels whose values did not fit the width their sample suggested
_narrow_hint: dict = {}       # column label -> width its last chunk crossed with (saves the sample of the next chunk)
narrowed_columns = 0          # columns that crossed PCIe narrow (tests, bench records)
deferred_columns = 0          # ... of which the host pass ran in the background of a band's other uploads
skipped_columns = 0           # columns sent at full width because their host pass had not started when PCIe ran dry


def _narrow_width(arr: np.ndarray) -> int:
    """bytes per element a strided sample of the int64 (or integer-valued float64) column needs (8: leave it alone);
    the full column is checked by the narrowing pass itself"""
    n = arr.shape[0]
    s = np.concatenate([arr[:256], arr[::max(1, n // 256)]])       # a few cache lines: the column is cold
    if arr.dtype == np.float64:
        with np.errstate(invalid="ignore"):
            if not bool(np.all(s == np.rint(s))) or not bool(np.all(np.abs(s) <= 2.0 ** 31)):   # fractions, NaN, inf
                return 8
    lo, hi = int(s.min()), int(s.max())
    for width, bound in ((1, 1 << 7), (2, 1 << 15), (4, 1 << 31)):
        if -bound <= lo and hi < bound:
            return width
    return 8


def _narrow_failed(label, width) -> None:
    """a value did not fit: with a width remembered from an earlier chunk the next chunk samples again, with a freshly
    sampled width the label is not tried any more"""
    if label is None:
        return
    if _narrow_hint.pop(label, None) != width:
        _no_narrow.add(label)


def _host_tensor(arr: np.ndarray) -> torch.Tensor:
    if not arr.flags.writeable:
        # pandas hands out read-only views of the column; the copy only reads, so re-flag the view instead of
        # copying the chunk on the host
        try:
            arr = arr.view()
            arr.flags.writeable = True
        except ValueError:
            arr = arr.copy()
    return torch.from_numpy(arr)


def _to_device(arr: np.ndarray, device, label=None) -> torch.Tensor:
    """H2D copy of one 8-byte column (pandas hands out read-only views; torch wants writable memory).  Pinned
    memory is copied asynchronously as it is; pageable memory goes through the pinned staging ring; int64 columns
    (and float64 columns holding integers) whose values fit a narrower integer cross the link narrow and are restored
    in HBM.  A label whose column did not pass is not tried again (most float columns: the sample already says no)."""
    global h2d_bytes, narrowed_columns
    arr = np.ascontiguousarray(arr)
    on_gpu = getattr(device, "type", "cuda") == "cuda" and torch.cuda.is_available()
    if (NARROW and on_gpu and (arr.dtype == np.int64 or (NARROW_FLOATS and arr.dtype == np.float64)) and arr.ndim == 1
            and arr.nbytes >= NARROW_MIN_BYTES and label not in _no_narrow):
        width = _narrow_hint.get(label) or _narrow_width(arr)
        if width < 8:
            if getattr(_tls, "pending", None) is not None:
                out = _staging.submit_narrow(arr, device, width, label)
                if out is not None:
                    return out                        # accounted for when the block ends
            out = _staging.upload_narrow(arr, device, width)
            if out is not None:
                h2d_bytes += arr.size * width
                narrowed_columns += 1
                if label is not None:
                    _narrow_hint[label] = width
                return out
            _narrow_failed(label, width)
        elif label is not None:
            _no_narrow.add(label)
    h2d_bytes += arr.size * arr.itemsize
    t = _host_tensor(arr)
    arr = t.numpy()
    if (arr.nbytes >= STAGE_MIN_BYTES and arr.dtype.itemsize == 8 and arr.ndim == 1 and getattr(device, "type", "cuda") == "cuda"
            and torch.cuda.is_available() and not t.is_pinned()):
        return _staging.upload(arr, device)
    return t.to(device, non_blocking=True)


class Pending:
    """Row count of a partial tuple that only the device knows so far.

    The low-cardinality kernels can be enqueued without the host waiting for their result size: the output
    columns are allocated at capacity, the group count lands in a device int64, and the frames built on the
    columns carry this object instead of a length.  The first consumer that needs host-side sizes calls
    ``resolve()``: one 8-byte read.  A count of -1 means the speculation failed (the chunk was not a dense
    shape / a merge input was too large): ``replay`` then recomputes the columns on the general path from the
    inputs it kept alive, and every frame of the tuple is re-pointed at the new columns."""

    __slots__ = ("block", "row", "cap", "_frames", "replay", "value", "fblock", "layout")

    def __init__(self, block: torch.Tensor, row: int, cap: int, replay, fblock=None, layout=None):
        self.block, self.row, self.cap = block, row, cap
        # float64 block and column layout of the same result (kernels._FastCall): lets a consumer ship the
        # whole speculative result (both blocks, count row included) without looking at it
        self.fblock, self.layout = fblock, layout
        # weak: frames own their Pending (and through its replay closure the inputs), never the other way
        # round -- no reference cycle, so dropping the frames frees the inputs immediately
        self._frames: List[Any] = []
        self.replay = replay
        self.value: Optional[int] = None

    @property
    def frames(self):
        return [f for f in (r() for r in self._frames) if f is not None]

    def attach(self, frame):
        self._frames.append(weakref.ref(frame))

    def count_ptr(self) -> int:
        return self.block.data_ptr() + self.row * self.block.shape[1] * 8

    def settle(self, n: int):
        """the count became known (>= 0) by other means (e.g. it travelled with the result block)"""
        if self.value is None:
            self.value = n
            for f in self.frames:
                f._finalize(n)
            self.replay = None

    def resolve(self) -> int:
        if self.value is None:
            n = int(self.block[self.row, 0].item())          # 8-byte D2H + sync
            if n >= 0:
                self.settle(n)
            else:
                keys, accs = self.replay()                    # general path, eager
                n = int(keys[0].numel())
                self.value = n
                for f in self.frames:
                    f._repoint(keys, accs, n)
                self.replay = None
        return self.value


_dtypes_cache: dict = {}


class ColumnSet:
    """Equally long device columns that many frames share -- the output of one partition scatter, or the
    int64 / float64 block pair of one aggregation result -- with their base pointers known without creating a
    tensor view per column.  ``cols`` (the views) are only built when a consumer asks for tensors."""
    __slots__ = ("_cols", "ptrs", "_block", "dtypes", "device")

    def __init__(self, cols: Sequence[torch.Tensor]):
        self._cols = list(cols)
        self.ptrs = [t.data_ptr() for t in self._cols]
        self._block = None
        self.dtypes = [t.dtype for t in self._cols]
        self.device = self._cols[0].device if self._cols else None

    @classmethod
    def from_block(cls, bi: torch.Tensor, bf: torch.Tensor, layout) -> "ColumnSet":
        """columns of a result block pair (kernels._FastCall layout): keys first, then the accumulators in spec
        order; pointers by arithmetic on the two base pointers"""
        s = cls.__new__(cls)
        s._cols = None
        s._block = (bi, bf, layout)
        pi, pf, stride = bi.data_ptr(), bf.data_ptr(), bi.shape[1] * 8
        s.ptrs = [pi + j * stride for j in range(layout.nk)] + [
            (pf + layout.float_row[a] * stride) if layout.is_float[a] else (pi + layout.int_row[a] * stride)
            for a in range(layout.na)]
        s.dtypes = layout.col_dtypes
        s.device = bi.device
        return s

    @property
    def cols(self):
        if self._cols is None:
            bi, bf, layout = self._block
            k, a = layout.views(bi, bf)
            self._cols = list(k) + list(a)
        return self._cols


class DeviceFrame:
    __slots__ = ("index_names", "_index", "columns", "_data", "_n", "_pending", "_slot", "_home", "_src",
                 "_ptrs_cache", "range_start", "_filter", "__weakref__")

    def __init__(self, index_names: Sequence[Any], index: Optional[Sequence[torch.Tensor]],
                 columns: Sequence[Any], data: Sequence[torch.Tensor], n_rows: Optional[int] = None,
                 pending: Optional[Pending] = None, slot=None):
        self.index_names = list(index_names)
        self._index = None if index is None else list(index)   # None == RangeIndex (raw input chunk)
        self.columns = list(columns)
        self._data = list(data)
        self._pending = pending
        self._slot = slot
        self._home = None            # device of lazily uploaded columns (set by from_pandas(lazy=True))
        self._src = None             # row range of scattered columns that has not been sliced yet (from_range)
        self._ptrs_cache = None      # (column selection, pointers) of the last band-level batch over this chunk
        self.range_start = None      # raw chunk (index None): first label of its RangeIndex, when known
        self._filter = None          # pending row filter (column, op code, value): rows are dropped when first needed
        if pending is not None:
            pending.attach(self)
            self._n = None
            return
        if n_rows is None:
            if self._data:
                n_rows = int(self._data[0].numel())
            elif self._index:
                n_rows = int(self._index[0].numel())
            else:
                n_rows = 0
        self._n = n_rows

    # ---- row range of shared columns (shuffle blocks) ---------------------------------------------------
    @classmethod
    def from_range(cls, index_names, columns, src: "ColumnSet", k0: int, k1: int, d0: int, d1: int, a: int,
                   b: int, pending: Optional[Pending] = None, slot=None) -> "DeviceFrame":
        """Rows [a, b) of ``src.cols[k0:k1]`` (group keys) and ``src.cols[d0:d1]`` (data): a shuffle block.  A
        mapper emits R x K of these per chunk; nothing is sliced until a consumer asks for tensors, and the
        aggregation kernels never do -- they take ``ptrs()``."""
        f = cls.__new__(cls)
        f.index_names = index_names
        f._index = None
        f.columns = columns
        f._data = None
        f._n = b - a
        f._pending = pending
        f._slot = slot
        f._home = None
        f._ptrs_cache = None
        f.range_start = None
        f._filter = None
        f._src = (src, k0, k1, d0, d1, a, b)
        if pending is not None:      # rows [0, capacity) of a speculative result: the size is still on the device
            pending.attach(f)
            f._n = None
        return f

    def _mat(self):
        src, k0, k1, d0, d1, a, b = self._src
        cols = src.cols
        self._index = [t[a:b] for t in cols[k0:k1]]
        self._data = [t[a:b] for t in cols[d0:d1]]
        self._src = None

    def ptrs_raw(self):
        """like ``ptrs`` but without resolving a pending size (capacity-length columns; the consumer reads the
        row count on the device)"""
        if self._src is not None:
            src, k0, k1, d0, d1, a, b = self._src
            off, p = a * 8, src.ptrs
            return [x + off for x in p[k0:k1]], [x + off for x in p[d0:d1]]
        idx = self._index
        return ([t.data_ptr() for t in idx] if idx else []), [t.data_ptr() for t in self._data]

    def ptrs(self):
        """(key column pointers, data column pointers) as integers"""
        if self._src is not None:
            src, k0, k1, d0, d1, a, b = self._src
            off, p = a * 8, src.ptrs
            return [x + off for x in p[k0:k1]], [x + off for x in p[d0:d1]]
        idx = self.index
        return ([t.data_ptr() for t in idx] if idx else []), [t.data_ptr() for t in self.data]

    # ---- lazy row filter / derived columns (fused pre-path) ---------------------------------------------
    @property
    def is_lazy(self) -> bool:
        return self._filter is not None or (self._data is not None and any(type(t) is DerivedColumn for t in self._data))

    @property
    def physical_rows(self) -> int:
        """rows stored (before a pending row filter is applied)"""
        return self._n

    def lazy_spec(self):
        """hashable description of the pending filter and the derived columns (equal for the chunks of one frame)"""
        der = tuple((self.columns[i], tuple(t.factors)) for i, t in enumerate(self._data) if type(t) is DerivedColumn)
        return self._filter, der

    def with_filter(self, column, op: str, value) -> "DeviceFrame":
        """shallow copy whose rows are (lazily) restricted to ``column <op> value``"""
        if self._filter is not None:
            self._apply_lazy()
        out = DeviceFrame(self.index_names, self._index, self.columns, list(self.raw_data), self._n)
        out._home, out.range_start = self._home, self.range_start
        out._filter = (column, FILTER_OPS[op], value)
        return out

    def with_column(self, name, factors) -> "DeviceFrame":
        """shallow copy with column ``name`` set to the product of the given affine factors of data columns"""
        cols, data = list(self.columns), list(self.raw_data)
        for c, _, _ in factors:
            j = cols.index(c)
            if type(data[j]) is DerivedColumn:
                raise NotImplementedError("derived columns of derived columns are materialised by the caller")
            if data[j].dtype != torch.float64:
                raise NotImplementedError(f"column {c!r} has dtype {data[j].dtype}: float64 operands only")
        d = DerivedColumn(factors)
        if name in cols:
            data[cols.index(name)] = d
        else:
            cols.append(name)
            data.append(d)
        out = DeviceFrame(self.index_names, self._index, cols, data, self._n)
        out._home, out.range_start, out._filter = self._home, self.range_start, self._filter
        return out

    def physical(self, name) -> torch.Tensor:
        """a stored data column on the device, WITHOUT applying a pending row filter (fused pre-path only)"""
        i = self.columns.index(name)
        t = self._data[i]
        if isinstance(t, HostColumn):
            t = self._data[i] = _to_device(t.array, self._home, self.columns[i])
        if type(t) is DerivedColumn:
            raise KeyError(name)
        return t

    def _materialise_derived(self, i):
        t = self._data[i]
        names = []
        for c, _, _ in t.factors:
            if c not in names:
                names.append(c)
        phys = [self.physical(c) for c in names]
        self._data[i] = _lazy_kernels.eval_product(phys, [(names.index(c), a, b) for c, a, b in t.factors])
        return self._data[i]

    def _apply_lazy(self):
        """materialise what is pending: derived columns by elementwise kernels, the row filter by a stable
        compaction of every column (what the unfused operator chain of the reference does, pass by pass)"""
        if self._data is None:
            return
        for i, t in enumerate(self._data):
            if type(t) is DerivedColumn:
                self._materialise_derived(i)
        if self._filter is not None:
            col, op, value = self._filter
            self._filter = None
            for i, t in enumerate(self._data):
                if isinstance(t, HostColumn):
                    self._data[i] = _to_device(t.array, self._home, self.columns[i])
            self._home = None
            n = self._n
            idx = list(self._index) if self._index is not None else (
                [torch.arange(self.range_start, self.range_start + n, dtype=torch.int64, device=self._data[0].device)]
                if self.range_start is not None else [])
            cols = idx + list(self._data)
            outs, kept = _lazy_kernels.filter_rows(self._data[self.columns.index(col)], op, value, cols)
            if idx:
                self._index, self.index_names = outs[:len(idx)], (self.index_names or [None] * len(idx))
            self._data = outs[len(idx):]
            self._n = kept
            self._ptrs_cache = None

    # ---- deferred size -------------------------------------------------------------------------------
    @property
    def is_pending(self) -> bool:
        return self._pending is not None and self._pending.value is None

    def _resolve(self):
        if self._filter is not None:
            self._apply_lazy()
        if self._pending is not None and self._pending.value is None:
            self._pending.resolve()

    def _finalize(self, n: int):
        """the speculative columns are valid: cut the capacity-length views down to n rows"""
        if self._src is not None:
            src, k0, k1, d0, d1, a, _ = self._src
            self._src = (src, k0, k1, d0, d1, a, a + n)
            self._n = n
            return
        if self._index is not None:
            self._index = [t[:n] for t in self._index]
        self._data = [t[:n] for t in self._data]
        self._n = n

    def _repoint(self, keys, accs, n: int):
        self._src = None
        a0, a1 = self._slot
        self._index = list(keys)
        self._data = list(accs[a0:a1])
        self._n = n

    @property
    def index(self):
        if self._src is not None:
            self._mat()
        self._resolve()
        return self._index

    @index.setter
    def index(self, value):
        self._index = value

    @property
    def data(self):
        if self._src is not None:
            self._mat()
        self._resolve()
        if self._home is not None:
            for i, t in enumerate(self._data):
                if isinstance(t, HostColumn):
                    self._data[i] = _to_device(t.array, self._home, self.columns[i])
            self._home = None
        for i, t in enumerate(self._data):
            if type(t) is DerivedColumn:
                self._materialise_derived(i)
        return self._data

    def col(self, i: int) -> torch.Tensor:
        """data column i on the device (uploads it now if it is still in host memory)"""
        if self._src is not None:
            self._mat()
        if self._filter is not None:
            self._apply_lazy()
        t = self._data[i]
        if isinstance(t, HostColumn):
            t = self._data[i] = _to_device(t.array, self._home, self.columns[i])
        elif type(t) is DerivedColumn:
            t = self._materialise_derived(i)
        elif self._pending is not None:
            return self.data[i]
        return t

    def col_dtype(self, i: int):
        if self._src is not None:
            return self._src[0].dtypes[self._src[3] + i]
        return self._data[i].dtype

    @data.setter
    def data(self, value):
        self._data = value

    @property
    def raw_index(self):
        """index columns as they are (capacity-length while the size is pending)"""
        if self._src is not None:
            self._mat()
        return self._index

    @property
    def raw_data(self):
        if self._src is not None:
            self._mat()
        return self._data

    # ---- pandas-like surface used by metas / size accounting (subtask/utils.py:62-68)
    def __len__(self):
        self._resolve()
        return self._n

    @property
    def shape(self):
        return (len(self), len(self.columns))

    @property
    def ndim(self):
        return 2

    @property
    def dtypes(self):
        if self._src is not None:
            self._mat()
        # one Series object per schema: graph building looks at it for every chunk and caches on its identity
        try:
            key = (tuple(self.columns), tuple(t.dtype for t in self._data))
            hit = _dtypes_cache.get(key)
        except TypeError:
            key, hit = None, None
        if hit is None:
            hit = pd.Series([np.dtype(str(t.dtype).replace("torch.", "")) for t in self._data], index=self.columns)
            if key is not None:
                if len(_dtypes_cache) > 256:
                    _dtypes_cache.clear()
                _dtypes_cache[key] = hit
        return hit

    @property
    def nbytes(self) -> int:
        if self._src is not None:
            _, k0, k1, d0, d1, a, b = self._src
            return (b - a) * 8 * (k1 - k0 + d1 - d0)
        total = sum(t.numel() * t.element_size() for t in self._data)
        if self._index:
            total += sum(t.numel() * t.element_size() for t in self._index)
        return total

    def __sizeof__(self):
        return 64 + self.nbytes

    @property
    def device(self):
        if self._src is not None:
            return self._src[0].device
        if self._home is not None:
            return self._home
        ts = self._data or self._index or []
        return ts[0].device if ts else torch.device("cpu")

    def column(self, name) -> torch.Tensor:
        return self.data[self.columns.index(name)]

    def slice_rows(self, start: int, stop: int) -> "DeviceFrame":
        return DeviceFrame(self.index_names, None if self.index is None else [t[start:stop] for t in self.index],
                           self.columns, [t[start:stop] for t in self.data], stop - start)

    def same_index(self, other: "DeviceFrame") -> bool:
        """True when both frames share their index tensors (reference: item.index.equals(index),
        groupby/core.py:381-391 -- here a pointer comparison, no data pass)."""
        sa, sb = self._src, other._src
        if sa is not None and sb is not None:   # two unsliced row ranges: same columns, same rows
            return sa[0] is sb[0] and sa[1] == sb[1] and sa[2] == sb[2] and sa[5] == sb[5] and sa[6] == sb[6]
        if sa is not None:
            self._mat()
        if sb is not None:
            other._mat()
        a, b = self._index, other._index
        if a is None or b is None:
            return a is b and len(self) == len(other)
        if len(a) != len(b):
            return False
        if self._pending is not other._pending and (self.is_pending or other.is_pending):
            return False
        return all(x.data_ptr() == y.data_ptr() and x.numel() == y.numel() for x, y in zip(a, b))

    # ---- host <-> device (reference: DataFrameToGPU / DataFrameToCPU, dataframe/base/to_gpu.py:28-35)
    @classmethod
    def from_pandas(cls, df: pd.DataFrame, device="cuda", index_as_keys: bool = False, lazy: bool = False) -> "DeviceFrame":
        """H2D copy of a pandas chunk.  Numeric columns only: float64, int64 (narrower ints / bool are widened
        to int64, which is what pandas' own group_sum does); anything else raises NotImplementedError so
        that the caller can fall back to the reference executor.  lazy: data columns stay on the host until a
        consumer asks for them (``col(i)`` / ``data``)."""
        cols, data = [], []
        for name in df.columns:
            s = df[name]
            arr = s.to_numpy()
            if arr.dtype == np.float64 or arr.dtype == np.int64:
                pass
            elif arr.dtype.kind in "iub" and arr.dtype.itemsize <= 8 and arr.dtype != np.uint64:
                arr = arr.astype(np.int64)
            else:
                raise NotImplementedError(f"column {name!r} has dtype {arr.dtype}: the GPU groupby path handles "
                                          f"float64 and int64 columns")
            data.append(HostColumn(arr) if lazy else _to_device(arr, device, name))
            cols.append(name)
        index = None
        names: List[Any] = []
        if index_as_keys:
            names = list(df.index.names)
            if isinstance(df.index, pd.MultiIndex):
                levels = [df.index.get_level_values(i).to_numpy() for i in range(df.index.nlevels)]
            else:
                levels = [df.index.to_numpy()]
            index = []
            for lv in levels:
                if lv.dtype != np.int64:
                    raise NotImplementedError(f"group key dtype {lv.dtype}: int64 keys only")
                index.append(_to_device(lv, device))
        out = cls(names, index, cols, data, len(df))
        if index is None and isinstance(df.index, pd.RangeIndex) and df.index.step == 1:
            out.range_start = int(df.index.start)
        if lazy:
            out._home = torch.device(device) if not isinstance(device, torch.device) else device
        return out

    def to_pandas(self) -> pd.DataFrame:
        data = {i: t.cpu().numpy() for i, t in enumerate(self.data)}
        df = pd.DataFrame(data)
        df.columns = pd.Index(self.columns) if not (self.columns and isinstance(self.columns[0], tuple)) \
            else pd.MultiIndex.from_tuples(self.columns)
        if self.index is None and self.range_start is not None:
            df.index = pd.RangeIndex(self.range_start, self.range_start + len(df))
        if self.index is not None:
            arrays = [t.cpu().numpy() for t in self.index]
            if len(arrays) == 1:
                df.index = pd.Index(arrays[0], name=self.index_names[0] if self.index_names else None)
            else:
                df.index = pd.MultiIndex.from_arrays(arrays, names=self.index_names)
        return df

    def __repr__(self):
        rows = "pending" if self.is_pending else self._n
        has_index = self._index is not None or (self._src is not None and self._src[2] > self._src[1])
        return (f"DeviceFrame(rows={rows}, index={self.index_names if has_index else 'range'}, "
                f"columns={self.columns}, device={self.device})")


class PiecewiseFrame:
    """Row-wise concatenation of DeviceFrames that is NOT materialised: what the GPU executors return for
    ``pd.concat(axis=0)`` of partial frames (DataFrameConcat tuple branch, merge/concat.py:339-348, and the
    reducer's gather, groupby/core.py:457).  The consumer -- a combine / agg stage -- reads the pieces in
    place, so the concat costs no kernel and no copy.  ``materialize()`` performs the real concatenation for
    consumers that need one contiguous frame."""

    __slots__ = ("pieces",)

    def __init__(self, pieces: Sequence[DeviceFrame]):
        flat: List[DeviceFrame] = []
        for p in pieces:
            flat.extend(p.pieces if isinstance(p, PiecewiseFrame) else [p])
        self.pieces = flat

    def __len__(self):
        return sum(len(p) for p in self.pieces)

    @property
    def columns(self):
        return self.pieces[0].columns

    @property
    def index_names(self):
        return self.pieces[0].index_names

    @property
    def shape(self):
        return (len(self), len(self.columns))

    @property
    def ndim(self):
        return 2

    @property
    def dtypes(self):
        return self.pieces[0].dtypes

    @property
    def nbytes(self):
        return sum(p.nbytes for p in self.pieces)

    def __sizeof__(self):
        return 64 + self.nbytes

    def materialize(self, concat_columns, shared_keys=None) -> DeviceFrame:
        first = self.pieces[0]
        if len(self.pieces) == 1:
            return first
        keys = shared_keys
        if keys is None and first.index is not None:
            keys = [concat_columns([f.index[j] for f in self.pieces]) for j in range(len(first.index))]
        data = [concat_columns([f.data[j] for f in self.pieces]) for j in range(len(first.columns))]
        return DeviceFrame(first.index_names, keys, first.columns, data, len(self))

    def to_pandas(self) -> pd.DataFrame:
        return pd.concat([p.to_pandas() for p in self.pieces], axis=0)

    def __repr__(self):
        return f"PiecewiseFrame(rows={len(self)}, pieces={len(self.pieces)}, columns={self.columns})"
