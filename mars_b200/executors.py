"""GPU executors for the groupby-aggregation operands: ``fn(ctx, op) -> None``.

Each function replaces one ``execute`` classmethod of the reference and is registered through
``Operand.register_executor`` (mars/core/operand/core.py:458-464).  They are written against the operand
*protocol* only (``op.stage``, ``op.inputs/outputs``, ``op.groupby_params``, ``op.pre_funcs/agg_funcs/
post_funcs``, ``op.shuffle_size``, ``op.iter_mapper_data(ctx)``, ``ctx.get_current_chunk()``), so the same
code serves the host mirror in ``mars_b200.dataframe`` and the real reference classes (``mars_b200.plugin``).

  execute_groupby_agg      DataFrameGroupByAgg.execute        aggregation.py:1269-1284 (map/combine/agg)
  execute_groupby_operand  DataFrameGroupByOperand.execute    groupby/core.py:353-508 (shuffle map/reduce)
  execute_concat           DataFrameConcat.execute (tuples)   merge/concat.py:339-348
  execute_to_gpu           DataFrameToGPU.execute             base/to_gpu.py:28-35
  execute_datasource       (host chunk provider of the mirror)

Values in ``ctx``: raw chunks and partials are ``DeviceFrame`` objects (HBM resident); the final stage=agg
output is a ``pandas.DataFrame`` because chunk metas are computed from it (subtask/utils.py:62-68).
All arithmetic happens in libmarsgb.so (mars_b200.kernels); there is no CPU fallback.
"""
from __future__ import annotations

import contextlib
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd
import torch

from . import _lib as _lib_mod
from . import kernels as K
from ._lib import MarsGBError
from . import device as _device_mod
from .device import ColumnSet, DerivedColumn, DeviceFrame, Pending, PiecewiseFrame
from .reduction import _DECOMP

_OPCODE = {"sum": K.SUM, "count": K.COUNT, "min": K.MIN, "max": K.MAX}
_DECOMP_EXT = dict(_DECOMP, std=_DECOMP["var"])

# the kernels module is looked up through this indirection so that host-logic tests can substitute a
# recording double; the product never rebinds it.
_kernels = K
_device_mod._lazy_kernels = K     # materialisation of lazy row filters / derived columns (device.DeviceFrame)
_dense_rejected = set()
_mid_rejected = set()      # shapes the mid-cardinality batch kernel declined (functions, groups, NaNs)
_probed = set()            # shapes whose first chunk was probed for its cardinality
_distinct_shapes = set()   # shapes whose chunks spilled most rows: skip the shared-memory pre-aggregation
# MGB_EAGER=1 disables the speculative asynchronous chain (every operand then waits for its result size)
_ASYNC = os.environ.get("MGB_EAGER", "0") in ("", "0")
_LAZY_H2D = os.environ.get("MGB_EAGER_H2D", "0") in ("", "0")


def _kernels_limits():
    """(max keys, max value columns, max accumulators) of the C ABI (include/mgb.h)"""
    return _lib_mod.MGB_MAX_KEYS, _lib_mod.MGB_MAX_VALS, _lib_mod.MGB_MAX_ACCS


def _check_key_count(n: int):
    if n > _kernels_limits()[0]:
        raise NotImplementedError(f"{n} group keys: the GPU path handles up to {_kernels_limits()[0]} int64 keys")


def _stage(op) -> Optional[str]:
    st = getattr(op, "stage", None)
    return None if st is None else getattr(st, "name", str(st))


def _device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device())


def _as_device_frame(value, index_as_keys=False, lazy=False) -> DeviceFrame:
    if isinstance(value, DeviceFrame):
        return value
    if isinstance(value, pd.DataFrame):
        if index_as_keys:
            _check_key_count(value.index.nlevels)      # before any byte is copied: the caller falls back
        return DeviceFrame.from_pandas(value, device=_device(), index_as_keys=index_as_keys, lazy=lazy)
    raise NotImplementedError(f"cannot move {type(value).__name__} to the GPU groupby path")


# ----------------------------------------------------------------------------------------------------
# step planning
# ----------------------------------------------------------------------------------------------------
def classify_pre(func) -> str:
    """'identity' or 'square'.  Mirror steps carry ``kind``; the reference's generated callables are probed
    on a two-element series (their source is ``return invar0`` or ``invar0.pow(2)``, Appendix A.3)."""
    kind = getattr(func, "kind", None)
    if kind is not None:
        return kind
    probe = pd.Series([3.0, -2.0])
    try:
        out = func(probe, gpu=False)
    except TypeError:
        out = func(probe)
    out = np.asarray(out, dtype="float64")
    if np.array_equal(out, [3.0, -2.0]):
        return "identity"
    if np.array_equal(out, [9.0, 4.0]):
        return "square"
    raise NotImplementedError("pre-aggregation function is neither identity nor x**2")


def classify_post(func, n_inputs: int) -> Optional[Tuple[str, int]]:
    """(kind, ddof) for identity / mean / var post functions, None if unrecognised (then the callable itself
    is evaluated on host copies of the merged partials)."""
    kind = getattr(func, "kind", None)
    if kind is not None:
        return kind, getattr(func, "ddof", 1)

    def call(*a):
        try:
            return np.asarray(func(*a, gpu=False), dtype="float64")
        except TypeError:
            return np.asarray(func(*a), dtype="float64")

    try:
        if n_inputs == 1:
            x = pd.Series([3.0, 5.0])
            return ("identity", 1) if np.array_equal(call(x), [3.0, 5.0]) else None
        if n_inputs == 2:
            s, c = pd.Series([6.0, 9.0]), pd.Series([3, 2])
            return ("mean", 1) if np.array_equal(call(s, c), [2.0, 4.5]) else None
        if n_inputs == 3:
            q, s, c = pd.Series([14.0, 41.0]), pd.Series([6.0, 9.0]), pd.Series([3, 2])
            got = call(q, s, c)
            for ddof in (1, 0):
                if ddof == 0:
                    exp = (q / c - (s / c) ** 2).to_numpy()
                else:
                    exp = ((q - s ** 2 / c) / (c - ddof)).to_numpy()
                if np.allclose(got, exp, rtol=1e-15, atol=0):
                    return "var", ddof
    except Exception:  # noqa: BLE001
        return None
    return None


class _Step:
    __slots__ = ("pre_kind", "map_func", "agg_func", "output_key", "cols")

    def __init__(self, pre_kind, map_func, agg_func, output_key, cols):
        self.pre_kind, self.map_func, self.agg_func = pre_kind, map_func, agg_func
        self.output_key, self.cols = output_key, cols


def _needed_columns(op, value_cols: Sequence[Any], pre_kind: str, map_func: str, pre_cols) -> List[Any]:
    """Value columns on which the atomic step (pre_kind, map_func) is actually consumed by a user function.
    The reference evaluates every step on every referred column (its map frames carry all of them); the
    columns no post function reads are dead and are not computed here."""
    func = getattr(op, "func", None)
    base = [c for c in value_cols if pre_cols is None or c in pre_cols]
    if func is None:
        return base
    if isinstance(func, (list, tuple)):
        per_col = {c: list(func) for c in base}
    else:
        per_col = {c: list(fs) if isinstance(fs, (list, tuple)) else [fs] for c, fs in func.items()}
    out = []
    for c in base:
        for f in per_col.get(c, []):
            dec = _DECOMP_EXT.get(f) if isinstance(f, str) else None
            if dec is None:
                raise NotImplementedError(f"aggregation function {f!r} is outside the GPU groupby path")
            if any(k == pre_kind and m == map_func for k, m, _ in dec):
                out.append(c)
                break
    return out


def _plan_steps(op, value_cols) -> List[_Step]:
    pre_by_key = {}
    for p in (getattr(op, "pre_funcs", None) or []):
        pre_by_key[p.output_key] = (classify_pre(p.func), p.columns)
    steps = []
    for a in op.agg_funcs:
        if a.custom_reduction is not None or a.map_func_name == "custom_reduction":
            raise NotImplementedError("custom reductions run on the reference executor")
        if a.map_func_name not in _OPCODE or a.agg_func_name not in _OPCODE:
            raise NotImplementedError(f"atomic step {a.map_func_name}->{a.agg_func_name} is outside the GPU path")
        if a.kwds and not a.kwds.get("skipna", True):
            raise NotImplementedError("skipna=False is outside the GPU path")
        pre_kind, pre_cols = pre_by_key.get(a.input_key, ("identity", None))
        cols = _needed_columns(op, value_cols, pre_kind, a.map_func_name, pre_cols) if value_cols is not None else None
        steps.append(_Step(pre_kind, a.map_func_name, a.agg_func_name, a.output_key, cols))
    return steps


# ----------------------------------------------------------------------------------------------------
# DataFrameGroupByAgg
# ----------------------------------------------------------------------------------------------------
class _MapPlan:
    """Everything of a stage=map call that only depends on the operand's functions and the chunk schema."""
    __slots__ = ("by", "key_pos", "used", "val_pos", "accs", "slots", "shape_key", "_by_dtypes")

    def for_dtypes(self, frame):
        """per column-dtype signature of the chunk: (keys are int64?, positions (in `used`) of int64 value
        columns whose only accumulators are COUNTs, dtype codes of the value columns)"""
        sig = tuple(t.dtype for t in frame.raw_data)
        hit = self._by_dtypes.get(sig)
        if hit is None:
            keys_ok = all(sig[i] == torch.int64 for i in self.key_pos)
            skip = tuple(j for j, i in enumerate(self.val_pos)
                         if sig[i] == torch.int64 and all(op == K.COUNT for c, op in self.accs if c == j))
            dts = [K.F64 if sig[i] == torch.float64 else K.I64 for i in self.val_pos]
            hit = self._by_dtypes[sig] = (keys_ok, skip, dts)
        return hit


_map_plans: Dict[Any, _MapPlan] = {}


_map_plans_fast: Dict[Any, Any] = {}
_DEVICE_RESULTS = os.environ.get("MGB_DEVICE_RESULTS", "1") not in ("0", "")
_DEVICE_RESULT_ROWS = int(os.environ.get("MGB_DEVICE_RESULT_ROWS", "2049"))   # smaller results go straight to pandas
_PREAGG_GROUPS = 2048     # chunks with at most this many groups are worth the shared-memory pre-aggregation
_PROBE_MIN_ROWS = 65536   # smaller chunks are not worth the cardinality probe
# map stage of chunks whose sampled keys are >= 90 % distinct hands the rows on as their own partials (MGB_MAP_PASSTHROUGH=0:
# always group the chunk, like the reference's map stage)
_PASSTHROUGH = os.environ.get("MGB_MAP_PASSTHROUGH", "1") not in ("0", "")
passthrough_chunks = 0    # map operands that took that path (tests, bench records)
_pt_streak: Dict[Any, int] = {}       # shape -> consecutive chunks that took the pass-through
_group_ratio: Dict[Any, float] = {}   # shape -> groups / rows of the last chunk of that shape that was really grouped


def _map_plan(op, frame: DeviceFrame) -> _MapPlan:
    # fast path: the mirror's map operands of one tileable share their agg_funcs list object
    fast_key = (id(op.agg_funcs), id(op.groupby_params.get("by")), len(frame.columns))
    hit = _map_plans_fast.get(fast_key)
    if hit is not None and hit[0] is op.agg_funcs and hit[1] == frame.columns:
        return hit[2]
    plan = _map_plan_slow(op, frame)
    if len(_map_plans_fast) > 256:
        _map_plans_fast.clear()
    _map_plans_fast[fast_key] = (op.agg_funcs, list(frame.columns), plan)
    return plan


def _map_plan_slow(op, frame: DeviceFrame) -> _MapPlan:
    params = op.groupby_params
    by = params.get("by")
    selection = params.get("selection")
    try:
        sig = (tuple(frame.columns), tuple(by) if isinstance(by, (list, tuple)) else by,
               None if selection is None else tuple(selection), repr(getattr(op, "func", None)),
               tuple((a.input_key, a.map_func_name, a.agg_func_name, a.output_key) for a in op.agg_funcs),
               tuple((p.input_key, p.output_key, None if p.columns is None else tuple(p.columns))
                     for p in (op.pre_funcs or [])))
        plan = _map_plans.get(sig)
        if plan is not None:
            return plan
    except TypeError:
        sig = None
    if not isinstance(by, (list, tuple)) or not by or any(not isinstance(b, (str, int, tuple)) for b in by):
        raise NotImplementedError("groupby keys must be column labels on the GPU path")
    by = list(by)
    if len(by) > _kernels_limits()[0]:
        # library envelope (include/mgb.h: MGB_MAX_KEYS); raised BEFORE any data is touched so that the plugin
        # hands the operand to the reference executor (plugin._with_reference_fallback)
        raise NotImplementedError(f"{len(by)} group keys: the GPU path handles up to {_kernels_limits()[0]} int64 keys")
    value_cols = [c for c in frame.columns if c not in by and (selection is None or c in selection)]
    steps = _plan_steps(op, value_cols)
    used: List[Any] = []
    for s in steps:
        for c in s.cols:
            if c not in used:
                used.append(c)
    if not used:
        raise NotImplementedError("no value column to aggregate")
    accs, slots = [], []
    for s in steps:
        code = _OPCODE[s.map_func]
        if s.pre_kind == "square":
            if s.map_func != "sum":
                raise NotImplementedError("x**2 is only aggregated with sum")
            code = K.SUMSQ
        slots.append((list(s.cols), len(accs), len(accs) + len(s.cols)))
        for c in s.cols:
            accs.append((used.index(c), code))
    if len(used) > _kernels_limits()[1] or len(accs) > _kernels_limits()[2]:
        raise NotImplementedError(f"{len(used)} value columns / {len(accs)} map-side accumulators exceed the library "
                                  f"envelope ({_kernels_limits()[1]} / {_kernels_limits()[2]})")
    plan = _MapPlan()
    plan.by = by
    plan.key_pos = [frame.columns.index(b) for b in by]
    plan.used = used
    plan.val_pos = [frame.columns.index(c) for c in used]
    plan.accs = tuple(accs)
    plan.slots = slots
    plan.shape_key = (tuple(by), tuple(used), plan.accs)
    plan._by_dtypes = {}
    if sig is not None:
        if len(_map_plans) > 1024:
            _map_plans.clear()
        _map_plans[sig] = plan
    return plan


def _execute_agg_map(ctx, op):
    """stage=map: per-chunk group-by + K atomic reductions (reference: _execute_map, aggregation.py:1043-1128)"""
    frame = _as_device_frame(ctx[op.inputs[0].key])
    if frame.raw_index is not None:
        raise NotImplementedError("map stage expects a raw chunk (RangeIndex)")
    plan = _map_plan(op, frame)
    keys_ok, count_only, dts = plan.for_dtypes(frame)
    if not keys_ok:
        for b, i in zip(plan.by, plan.key_pos):
            if frame.col_dtype(i) != torch.int64:
                raise NotImplementedError(f"group key {b!r} has dtype {frame.col_dtype(i)}: int64 keys only")
    if frame.is_lazy and _ASYNC and hasattr(_kernels, "preagg_dense_batch_async"):
        # pending row filter / derived columns: the batch kernels evaluate them inside the pre-aggregation pass
        # when the shape allows (fused pre-path), else they are materialised on the way
        per_slot = _launch_band(plan, (keys_ok, count_only, dts), [frame])
        if per_slot is not None:
            frames = [p[0] for p in per_slot]
            rec = getattr(op, "size_recorder_name", None)
            if rec is not None:
                pend0 = frames[0]._pending
                bound = sum(pend0.cap * 8 * (len(f.columns) + len(plan.by)) for f in frames)
                _record_sizes(ctx, rec, frame.nbytes,
                              lambda: sum(len(f) * 8 * (len(f.columns) + len(plan.by)) for f in frames), bound)
            ctx[op.outputs[0].key] = tuple(frames)
            return
    keys = [frame.col(i) for i in plan.key_pos]
    if _ASYNC and hasattr(_kernels, "preagg_dense_batch_async"):
        which = _band_kernel(plan, (keys_ok, count_only, dts), [frame])
    else:
        which = "dense" if plan.shape_key not in _dense_rejected else None
    dense_ok = which == "dense"
    # the dense kernel does not read int64 columns that are only counted (count == rows of the group): those
    # stay where they are (possibly in host memory) unless the general path turns out to be needed
    skip = count_only if dense_ok and _ASYNC and hasattr(_kernels, "preagg_dense_async") else ()
    vals = [None if j in skip else frame.col(i) for j, i in enumerate(plan.val_pos)]
    # low-cardinality chunks take the one-launch dense path, enqueued WITHOUT waiting for its result size
    # (the frames carry a Pending count).  A shape that turned out not to fit is remembered, so the following
    # chunks of the same computation go straight to the general hash-table path.
    rec = getattr(op, "size_recorder_name", None)
    frames = None

    def all_vals():
        return [frame.col(i) for i in plan.val_pos]

    if dense_ok:
        if _ASYNC and hasattr(_kernels, "preagg_dense_async"):
            res = _kernels.preagg_dense_async(keys, vals, plan.accs, checked=True, val_dtypes=dts)
            if res is None:
                _dense_rejected.add(plan.shape_key)
            else:
                bi, bf, fc = res
                accs = list(plan.accs)

                def replay(keys=keys, accs=accs, shape_key=plan.shape_key):
                    _dense_rejected.add(shape_key)
                    k, a, _ = _kernels.groupby_aggregate(keys, all_vals(), accs, sort=False)
                    return k, a

                pend = Pending(bi, fc.n_int, bi.shape[1], replay, fblock=bf, layout=fc)
                # the K partial frames are row ranges [0, capacity) of the block pair: no tensor view is made
                # unless a consumer asks for one (the merges of the tree take pointers)
                src = ColumnSet.from_block(bi, bf, fc)
                nk, cap = fc.nk, bi.shape[1]
                frames = [DeviceFrame.from_range(plan.by, cols, src, 0, nk, nk + a0, nk + a1, 0, cap, pending=pend,
                                                 slot=(a0, a1))
                          for cols, a0, a1 in plan.slots]
        elif hasattr(_kernels, "preagg_dense"):
            vals = all_vals()
            res = _kernels.preagg_dense(keys, vals, plan.accs)
            if res is None:
                _dense_rejected.add(plan.shape_key)
            else:
                out_keys, out_accs = res
                n_out = int(out_keys[0].numel())
                frames = [DeviceFrame(plan.by, out_keys, cols, out_accs[a0:a1], n_out) for cols, a0, a1 in plan.slots]
    if frames is None and which == "mid":
        # mid-cardinality sum / mean / count shape: one asynchronous launch of the locked-record kernel
        per_slot = _launch_band(plan, (keys_ok, count_only, dts), [frame])
        if per_slot is not None:
            frames = [p[0] for p in per_slot]
    if frames is None:
        vals = all_vals()
        # Shared-memory pre-aggregation pays while a CTA's table holds (nearly) all groups of the chunk; a
        # shape whose rows mostly spilled is remembered and its next chunks go straight to the bucketed
        # aggregation (distinct-keys hint), until a chunk comes back with few groups again.
        distinct = plan.shape_key in _distinct_shapes
        if not distinct and hasattr(_kernels, "estimate_distinct") and len(frame) >= _PROBE_MIN_ROWS:
            # first chunk of a shape: a strided sample of its keys tells nearly-distinct keys (bucketed aggregation
            # right away) from the regimes where the on-chip pre-aggregation absorbs most rows
            sampled, uniq = _kernels.estimate_distinct(keys)
            if sampled and uniq * 4 >= sampled * 3:
                _distinct_shapes.add(plan.shape_key)
                distinct = True
        passthrough = False
        if (distinct and _PASSTHROUGH and hasattr(_kernels, "row_partials") and hasattr(_kernels, "estimate_distinct")
                and len(frame) >= _PROBE_MIN_ROWS and _group_ratio.get(plan.shape_key, 0.0) >= 0.9):
            # the last chunk of this shape that WAS grouped came back with about as many groups as rows: grouping
            # reduces nothing, and the combine / agg stages re-group the partial rows anyway
            # (aggregation.py:1130-1267) -- the rows ARE the partials.  (A sample cannot tell this regime from many
            # groups of a few rows each: 16 K sampled keys of a million groups are all distinct, too; the sample is
            # only the cheap check that this chunk still looks like the one that was measured.)
            # ... sampled for the first chunks of a run and then every 8th one (a sample is a host wait)
            streak = _pt_streak.get(plan.shape_key, 0)
            if streak >= 2 and streak % 8:
                passthrough = True
            else:
                sampled, uniq = _kernels.estimate_distinct(keys)
                passthrough = bool(sampled) and uniq * 10 >= sampled * 9
            _pt_streak[plan.shape_key] = streak + 1 if passthrough else 0
        if passthrough:
            global passthrough_chunks
            passthrough_chunks += 1
            out_keys, out_accs, stats = list(keys), _kernels.row_partials(vals, list(plan.accs)), None
        elif distinct:
            out_keys, out_accs, stats = _kernels.groupby_aggregate(keys, vals, list(plan.accs), sort=False,
                                                                   expect_distinct=True)
            if stats and (int(out_keys[0].numel()) <= _PREAGG_GROUPS or stats.get("bucket_fallback") == 2):
                _distinct_shapes.discard(plan.shape_key)   # few groups, or hot keys that pre-aggregation absorbs
        else:
            out_keys, out_accs, stats = _kernels.groupby_aggregate(keys, vals, list(plan.accs), sort=False)
            if stats and stats.get("rows_spilled", 0) * 2 > int(keys[0].numel()):
                _distinct_shapes.add(plan.shape_key)
        n_out = int(out_keys[0].numel())
        if not passthrough and len(frame):
            _group_ratio[plan.shape_key] = n_out / len(frame)
        if n_out <= 32:
            _dense_rejected.discard(plan.shape_key)      # the data changed: the one-launch path fits again
        frames = [DeviceFrame(plan.by, out_keys, cols, out_accs[a0:a1], n_out) for cols, a0, a1 in plan.slots]
    if rec is not None:   # method='auto' samples sizes: exact ones resolve the count (a host wait), the capacity bound does not
        pend0 = frames[0]._pending
        bound = None
        if pend0 is not None and pend0.value is None:
            bound = sum(pend0.cap * 8 * (len(f.columns) + len(plan.by)) for f in frames)
        _record_sizes(ctx, rec, frame.nbytes,
                      lambda: sum(len(f) * 8 * (len(f.columns) + len(plan.by)) for f in frames), bound)
    ctx[op.outputs[0].key] = tuple(frames)


_size_recorder_cls = None


def _record_sizes(ctx, name, raw_size, agg_size, bound=None):
    """agg_size may be a zero-argument callable.  A real Mars recorder is a remote object and gets the exact
    number right away (aggregation.py:1120-1126).  The mirror's recorder evaluates it when the tiler reads the
    sizes; when ``bound`` is given -- the capacity of a still speculative result, known without asking the
    device -- the tiler first decides on the bounds and reads the exact sizes (one host wait) only if the bounds
    alone would send the plan to the shuffle branch (dataframe.DataFrameGroupByAgg.tile)."""
    global _size_recorder_cls
    getter = getattr(ctx, "get_remote_object", None)
    if getter is not None:
        getter(name).record(raw_size, agg_size() if callable(agg_size) else agg_size)
        return
    if _size_recorder_cls is None:
        from .dataframe import SizeRecorder

        _size_recorder_cls = SizeRecorder
    _size_recorder_cls.lookup(name).record(raw_size, agg_size, bound)


def _merge_partials(slots: Sequence[Sequence[DeviceFrame]], steps: Sequence[_Step], sort: bool) -> List[DeviceFrame]:
    """Re-aggregate partial frames slot-wise: slot k -> groupby(level).agg(agg_func_k)
    (reference: _execute_combine, aggregation.py:1130-1164; count partials are SUMmed as int64).
    ``slots[k]`` lists the pieces of slot k (a virtual concat: the pieces are read in place)."""
    if len(slots) != len(steps):
        raise MarsGBError(1, f"{len(slots)} partial frames for {len(steps)} aggregation steps")
    n_pieces = len(slots[0])
    first = slots[0]
    shared = all(len(sl) == n_pieces and all(first[i].same_index(sl[i]) for i in range(n_pieces))
                 for sl in slots[1:])
    groups: List[List[int]] = []
    if shared:
        cur, width = [], 0
        for i, sl in enumerate(slots):
            if cur and width + len(sl[0].columns) > 64:
                groups.append(cur)
                cur, width = [], 0
            cur.append(i)
            width += len(sl[0].columns)
        groups.append(cur)
    else:
        groups = [[i] for i in range(len(slots))]
    force_sort = sort or len(groups) > 1
    out: List[Optional[DeviceFrame]] = [None] * len(slots)
    for grp in groups:
        base = slots[grp[0]]
        ops = [_OPCODE[steps[i].agg_func] for i in grp for _ in slots[i][0].columns]
        pend_out = None
        # sizes first (known on the host without touching any column): large inputs skip the small-merge attempt
        known = sum(len(f) for f in base if not f.is_pending)
        n_live = sum(1 for f in base if f.is_pending or len(f) > 0)
        small = (known <= getattr(_kernels, "SMALL_MAX_ROWS", 2048)
                 and n_live <= getattr(_kernels, "SMALL_MAX_PIECES", 16))
        if small and _ASYNC and hasattr(_kernels, "merge_small_async"):
            # sizes that are still on the device stay there: the merge reads them in the kernel
            pieces = []
            dev = base[0].device
            dts = [_kernels.F64 if slots[s][0].col_dtype(j) == torch.float64 else _kernels.I64
                   for s in grp for j in range(len(slots[s][0].columns))]

            def piece_of(i, n, cptr):
                kp, ap = base[i].ptrs_raw()
                for s_ in grp[1:]:
                    ap = ap + slots[s_][i].ptrs_raw()[1]
                return (kp, ap, n, cptr, dts, dev)

            for i in range(len(base)):
                f0 = base[i]
                if f0.is_pending:
                    pieces.append(piece_of(i, None, f0._pending.count_ptr()))
                elif len(f0) > 0:
                    pieces.append(piece_of(i, len(f0), 0))
            if not pieces:
                pieces.append(piece_of(0, 0, 0))
            res = _kernels.merge_small_async(pieces, ops, sort=force_sort, checked=True)
            if res is not None:
                bi, bf, fc = res
                pend_out = Pending(bi, fc.n_int, bi.shape[1],
                                   lambda grp=grp, ops=ops: _merge_general(slots, grp, ops, force_sort),
                                   fblock=bf, layout=fc)
                out_src = ColumnSet.from_block(bi, bf, fc)
        if pend_out is None:
            res = None
            live = [i for i in range(len(base)) if len(base[i]) > 0] or [0]
            if hasattr(_kernels, "merge_small") and sum(len(base[i]) for i in live) <= 2048 and len(live) <= 16:
                res = _kernels.merge_small([base[i].index for i in live],
                                           [[t for s in grp for t in slots[s][i].data] for i in live], ops,
                                           sort=force_sort)
            k, a = res if res is not None else _merge_general(slots, grp, ops, force_sort)
        pos = 0
        for i in grp:
            ncol = len(slots[i][0].columns)
            if pend_out is not None:
                nk_ = pend_out.layout.nk
                out[i] = DeviceFrame.from_range(base[0].index_names, slots[i][0].columns, out_src, 0, nk_, nk_ + pos,
                                                nk_ + pos + ncol, 0, pend_out.cap, pending=pend_out,
                                                slot=(pos, pos + ncol))
            else:
                out[i] = DeviceFrame(base[0].index_names, k, slots[i][0].columns, a[pos:pos + ncol], int(k[0].numel()))
            pos += ncol
    return out  # type: ignore[return-value]


def _merge_general(slots, grp, ops, sort):
    """general hash aggregation over the pieces of the slots in ``grp`` (one update per piece, no concat);
    resolves pending sizes (and replays failed speculations) of the inputs"""
    base = slots[grp[0]]
    live = [i for i in range(len(base)) if len(base[i]) > 0] or [0]
    first_vals = [t for s in grp for t in slots[s][live[0]].data]
    dts = [_kernels.dtype_code(t) for t in first_vals]
    # every piece is a partial frame with unique keys: pre-aggregating a piece on its own cannot reduce it, so
    # the pieces go to the library with the distinct-keys hint (bucketed aggregation over all pieces in place)
    agg = _kernels.Aggregator(len(base[0].index), dts, list(enumerate(ops)), expect_distinct=True)
    try:
        if hasattr(agg, "update_ptrs"):     # pointer level: shuffle blocks are unsliced row ranges
            dev = base[0].device
            for i in live:
                fr = [slots[s][i] for s in grp]
                kp, vp = fr[0].ptrs()
                for f in fr[1:]:
                    vp = vp + f.ptrs()[1]
                agg.update_ptrs(kp, vp, len(fr[0]), fr, dev)
        else:
            for i in live:
                agg.update(base[i].index, [t for s in grp for t in slots[s][i].data])
        return agg.result(sort=sort, device=base[0].device)
    finally:
        agg.close()


def _pieces_of(value) -> List[List[DeviceFrame]]:
    """partial tuple -> per slot the list of pieces (DeviceFrame: one piece; PiecewiseFrame: its pieces)"""
    if not isinstance(value, tuple):
        value = (value,)
    slots = []
    for v in value:
        if isinstance(v, bool) or v is None or isinstance(v, list):   # deliver_by flag / trailing `by`
            continue
        if isinstance(v, PiecewiseFrame):
            slots.append(list(v.pieces))
        else:
            slots.append([_as_device_frame(v, index_as_keys=True)])
    return slots


def _partials_of(value) -> List[DeviceFrame]:
    """partial tuple -> one contiguous DeviceFrame per slot (materialises virtual concats)"""
    out = []
    for pieces in _pieces_of(value):
        out.append(pieces[0] if len(pieces) == 1 else PiecewiseFrame(pieces).materialize(_kernels.concat_columns))
    return out


def _execute_agg_combine(ctx, op):
    slots = _pieces_of(ctx[op.inputs[0].key])
    steps = _plan_steps(op, None)
    ctx[op.outputs[0].key] = tuple(_merge_partials(slots, steps, sort=False))


_post_plans: Dict[Any, Any] = {}


def _post_plan(op, slots, steps, two_level):
    """ctypes post steps of a stage=agg operand whose post functions are all identity / mean / var, or None.
    Cached on the identity of ``op.post_funcs`` (compiled reductions are shared between the operands of one
    computation) and the partial schema."""
    first = [sl[0] for sl in slots]
    key = (id(op.post_funcs), two_level, tuple(tuple(f.columns) for f in first),
           tuple(f.col_dtype(j) for f in first for j in range(len(f.columns))))
    hit = _post_plans.get(key)
    if hit is not None and hit[0] is op.post_funcs:
        return hit[1]
    off, pos = [], 0
    for f in first:
        off.append(pos)
        pos += len(f.columns)
    by_key = {s.output_key: i for i, s in enumerate(steps)}
    plan = None
    post_steps, labels, is_float = [], [], []
    ok = True
    for p in op.post_funcs:
        idx = [by_key[k] for k in p.input_keys]
        cols = list(p.columns) if p.columns is not None else None
        common = [c for c in first[idx[0]].columns if all(c in first[i].columns for i in idx[1:])
                  and (cols is None or c in cols)]
        cls = classify_post(p.func, len(idx))
        if cls is None:
            ok = False
            break
        kind, ddof = cls
        for c in common:
            acc = [off[i] + first[i].columns.index(c) for i in idx]
            labels.append((c, p.func_name) if two_level else c)
            if kind == "identity":
                post_steps.append((_kernels.POST_IDENTITY, acc[0], 0, 0, 0))
                is_float.append(first[idx[0]].col_dtype(first[idx[0]].columns.index(c)) == torch.float64)
            elif kind == "mean":
                post_steps.append((_kernels.POST_MEAN, acc[0], acc[1], 0, 0))
                is_float.append(True)
            else:   # inputs of the generated var: sum(x**2), sum(x), count
                post_steps.append((_kernels.POST_VAR, acc[1], acc[0], acc[2], int(ddof)))
                is_float.append(True)
    if ok and post_steps and len(post_steps) <= _kernels.MAX_POST and len(set(labels)) == len(labels):
        ops = [_OPCODE[steps[i].agg_func] for i, f in enumerate(first) for _ in f.columns]
        dts = [_kernels.F64 if f.col_dtype(j) == torch.float64 else _kernels.I64
               for f in first for j in range(len(f.columns))]
        plan = (_kernels.PostPlan(post_steps, is_float), labels, ops, dts)
    if len(_post_plans) > 256:
        _post_plans.clear()
    _post_plans[key] = (op.post_funcs, plan)
    return plan


def _agg_one_launch(ctx, op, out_chunk, slots, steps, sort, col_value) -> bool:
    """stage=agg of a small input in ONE launch + ONE device->host copy: merge of the partial pieces, the post
    functions and the key sort happen inside mgb_merge_small_post_async; the pieces' sizes may still be on the
    device.  Returns False when the input is not small / a post function is not one of identity, mean, var /
    the speculative chain failed -- the caller then takes the general route (which replays)."""
    if not (_ASYNC and hasattr(_kernels, "merge_small_post_async")):
        return False
    first = slots[0]
    n_pieces = len(first)
    if any(len(sl) != n_pieces for sl in slots[1:]):
        return False
    for sl in slots[1:]:
        for i in range(n_pieces):
            if not first[i].same_index(sl[i]):
                return False
    if sum(len(sl[0].columns) for sl in slots) > 64:
        return False
    known = sum(len(f) for f in first if not f.is_pending)
    n_live = sum(1 for f in first if f.is_pending or len(f) > 0)
    if known > _kernels.SMALL_MAX_ROWS or n_live > _kernels.SMALL_MAX_PIECES or n_live == 0:
        return False
    plan = _post_plan(op, slots, steps, col_value.nlevels > 1)
    if plan is None:
        return False
    post, labels, ops, dts = plan
    dev = first[0].device
    pieces = []
    for i in range(n_pieces):
        f0 = first[i]
        pend = f0.is_pending
        if not pend and len(f0) == 0:
            continue
        kp, ap = f0.ptrs_raw()
        for sl in slots[1:]:
            ap = ap + sl[i].ptrs_raw()[1]
        pieces.append((kp, ap, None if pend else len(f0), f0._pending.count_ptr() if pend else 0, dts, dev))
    res = _kernels.merge_small_post_async(pieces, ops, post, sort=sort)
    if res is None:
        return False
    blk, nk, n_post = res
    n, keys, cols = _kernels.result_block_to_host(blk, nk, n_post, post.is_float)
    if n < 0:
        return False          # a speculation failed up the chain: the general route resolves and replays
    ctx[out_chunk.key] = _assemble_result(dict(zip(labels, cols)), keys,
                                          _result_meta(op, out_chunk, first[0].index_names, col_value))
    return True


def _execute_agg_agg(ctx, op):
    """stage=agg: final merge, post functions, result frame (reference: _execute_agg, aggregation.py:1166-1267)"""
    out_chunk = op.outputs[0]
    slots = _pieces_of(ctx[op.inputs[0].key])
    steps = _plan_steps(op, None)
    sort = bool(op.groupby_params.get("sort", True))
    if len(slots) == len(steps) and _agg_one_launch(ctx, op, out_chunk, slots, steps, sort,
                                                    out_chunk.columns_value.to_pandas()):
        return
    merged = _merge_partials(slots, steps, sort=sort)
    col_value = out_chunk.columns_value.to_pandas()
    two_level = col_value.nlevels > 1
    by_key = {s.output_key: f for s, f in zip(steps, merged)}

    def post_columns(raw: bool):
        """device columns of the result (post functions applied); raw: work on capacity-length columns of a
        still-pending merge -- rows beyond the (device-side) count hold garbage that is cut off on the host"""
        pieces: Dict[Any, Any] = {}
        host_eval = []
        for p in op.post_funcs:
            inputs = [by_key[k] for k in p.input_keys]
            cols = list(p.columns) if p.columns is not None else None
            common = [c for c in inputs[0].columns if all(c in f.columns for f in inputs[1:])
                      and (cols is None or c in cols)]
            cls = classify_post(p.func, len(inputs))
            for c in common:
                label = (c, p.func_name) if two_level else c
                if cls is None:
                    host_eval.append((label, p, c, inputs))
                    continue
                kind, ddof = cls
                ts = [(f.raw_data if raw else f.data)[f.columns.index(c)] for f in inputs]
                if kind == "identity":
                    res = ts[0]
                elif kind == "mean":
                    res = _kernels.post_mean(ts[0], ts[1])
                else:
                    res = _kernels.post_var(ts[1], ts[0], ts[2], ddof)   # inputs: sum(x**2), sum(x), count
                pieces[label] = res
        return pieces, host_eval

    # all result columns, the group keys and -- when the merge is still pending -- its row count cross to the
    # host with ONE synchronisation
    base = merged[0]
    arrays = None
    pend = base._pending if base.is_pending else None
    if pend is not None and all(f._pending is pend for f in merged) and hasattr(_kernels, "to_host_counted"):
        pieces, host_eval = post_columns(raw=True)
        if not host_eval:
            labels_dev = list(pieces.keys())
            n, arrays = _kernels.to_host_counted([pieces[l] for l in labels_dev] + list(base.raw_index), pend.block,
                                                 pend.row)
            if n >= 0:
                pend.settle(n)
            else:
                arrays = None   # speculation failed somewhere up the chain: resolve (replays) and redo eagerly
    if arrays is None:
        pieces, host_eval = post_columns(raw=False)
        labels_dev = list(pieces.keys())
        if (not host_eval and _DEVICE_RESULTS and len(base) >= _DEVICE_RESULT_ROWS
                and hasattr(_kernels, "to_host_concat")):
            # large results stay in HBM; the D2H happens once, for all result chunks, at fetch time
            ctx[out_chunk.key] = DeviceResult(labels_dev, [pieces[l] for l in labels_dev], base.index,
                                              _result_meta(op, out_chunk, base.index_names, col_value))
            return
        arrays = _to_host([pieces[l] for l in labels_dev] + list(base.index))
    host = dict(zip(labels_dev, arrays[:len(labels_dev)]))
    idx_arrays = arrays[len(labels_dev):]
    for label, p, c, inputs in host_eval:
        args = [pd.Series(f.column(c).cpu().numpy()) for f in inputs]
        try:
            host[label] = np.asarray(p.func(*args, gpu=False))
        except TypeError:
            host[label] = np.asarray(p.func(*args))
    ctx[out_chunk.key] = _assemble_result(host, idx_arrays, _result_meta(op, out_chunk, base.index_names, col_value))


def _result_meta(op, out_chunk, index_names, col_value):
    names = index_names
    out_names = getattr(out_chunk.index_value, "names", None) if out_chunk.index_value is not None else None
    if out_names and len(out_names) == len(index_names) and any(n is not None for n in out_names):
        names = list(out_names)
    return dict(names=list(names), col_value=col_value, two_level=col_value.nlevels > 1,
                as_index=op.groupby_params.get("as_index", True), dtypes=getattr(out_chunk, "dtypes", None))


def _assemble_result(host: Dict[Any, Any], idx_arrays, meta) -> pd.DataFrame:
    """result frame from host arrays: labels / order / dtypes of the chunk meta (aggregation.py:1233-1267)"""
    names, col_value = meta["names"], meta["col_value"]
    if len(idx_arrays) == 1:
        index = pd.Index(idx_arrays[0], name=names[0] if names else None, copy=False)
    else:
        index = _multi_index(idx_arrays, names)
    labels = list(host.keys())
    result = None
    if meta["as_index"] and len(set(labels)) == len(labels) and set(labels) == set(col_value):
        # declared column order == a permutation of what was computed: assemble in that order, one block per
        # column and no consolidation copy (copy=False) -- the frame wraps the arrays that came off the device
        if len(index) < 4096 and hasattr(pd.DataFrame, "_from_arrays"):
            # tiny results: the array-list constructor skips most of the per-column validation (it copies
            # into consolidated blocks, which is free at this size)
            result = pd.DataFrame._from_arrays([host[l] for l in col_value], columns=col_value, index=index,
                                               verify_integrity=False)
        else:
            result = pd.DataFrame({i: host[l] for i, l in enumerate(col_value)}, index=index, copy=False)
            result.columns = col_value
    if result is None:
        result = pd.DataFrame({i: host[l] for i, l in enumerate(labels)}, index=index)
        result.columns = pd.MultiIndex.from_tuples(labels) if meta["two_level"] else pd.Index(labels)
        if not meta["as_index"]:
            result = result.reset_index()
        result = result.reindex(col_value, axis=1)
    dtypes = meta["dtypes"]
    if dtypes is not None and len(result) == 0:
        result = result.astype(dtypes)
    return result


class DeviceResult:
    """Result chunk of a stage=agg operand that stays in HBM (what the reference's GPU mode produces: its
    final chunks are cudf frames and the user calls ``.fetch().to_pandas()``, test_groupby_execution.py:788-793).
    Frame-like for the processor's meta code (``shape`` / ``dtypes`` / ``index`` names, subtask/utils.py:62-68);
    ``to_pandas()`` reads it back.  ``Session.fetch`` reads ALL result chunks into one pinned block and builds
    the final frame without a host-side concat."""

    __slots__ = ("labels", "columns", "index", "meta")

    def __init__(self, labels, columns, index, meta):
        self.labels, self.columns, self.index, self.meta = list(labels), list(columns), list(index), meta

    def __len__(self):
        return int(self.index[0].numel())

    @property
    def shape(self):
        return (len(self), len(self.meta["col_value"]))

    @property
    def ndim(self):
        return 2

    @property
    def dtypes(self):
        return self.meta["dtypes"]

    @property
    def nbytes(self):
        return sum(t.numel() * 8 for t in self.columns + self.index)

    def __sizeof__(self):
        return 64 + self.nbytes

    def to_pandas(self) -> pd.DataFrame:
        return DeviceResult.fetch_all([self])

    @staticmethod
    def fetch_all(parts: Sequence["DeviceResult"]) -> pd.DataFrame:
        """row-wise concatenation of result chunks, assembled on the host in one pinned block"""
        first = parts[0]
        for p in parts[1:]:
            if p.labels != first.labels or len(p.index) != len(first.index):
                raise MarsGBError(1, "result chunks differ in their columns")
        ncol = len(first.columns)
        arrays = _kernels.to_host_concat([list(p.columns) + list(p.index) for p in parts])
        return _assemble_result(dict(zip(first.labels, arrays[:ncol])), arrays[ncol:], first.meta)

    def __repr__(self):
        return f"DeviceResult(rows={len(self)}, columns={len(self.labels)})"


_small_index_cache: Dict[Any, Any] = {}


def _multi_index(arrays, names):
    # tiny results (the tree branch: a handful of groups): the same key tuples come back step after step; a
    # MultiIndex is immutable, so one that was built for exactly these keys and names is handed out again
    n = len(arrays[0])
    ck = None
    if n <= 64:
        try:
            ck = (tuple(names), tuple(a.tobytes() for a in arrays))
            hit = _small_index_cache.get(ck)
            if hit is not None:
                return hit
        except TypeError:
            ck = None
    mi = _multi_index_build(arrays, names)
    if ck is not None:
        if len(_small_index_cache) > 256:
            _small_index_cache.clear()
        _small_index_cache[ck] = mi
    return mi


def _multi_index_build(arrays, names):
    """pd.MultiIndex.from_arrays without its per-level Categorical round trip (the keys are int64 arrays)"""
    try:
        levels, codes = [], []
        for a in arrays:
            if len(a) < 4096:     # tiny results (the tree branch): the sort-based path has the smaller fixed cost
                lv, inv = np.unique(a, return_inverse=True)
            else:                 # hash-based: linear in the rows, sorts the distinct values only
                inv, lv = pd.factorize(a, sort=True)
            levels.append(pd.Index(lv))
            codes.append(inv)
        return pd.MultiIndex(levels=levels, codes=codes, names=names, verify_integrity=False)
    except Exception:  # noqa: BLE001
        return pd.MultiIndex.from_arrays(arrays, names=names)


def _to_host(columns):
    fn = getattr(_kernels, "to_host", None)
    if fn is not None:
        return fn(columns)
    return [t.cpu().numpy() for t in columns]


# ----------------------------------------------------------------------------------------------------
# band-level fusion of the tree branch: n stage=map operands + concat / stage=combine tree -> one launch
# ----------------------------------------------------------------------------------------------------
class FusedTree:
    """The part of a chunk graph below one stage=agg operand of the tree branch (aggregation.py:715-823):
    ``leaves`` are the not yet executed stage=map chunks, ``interior`` the DataFrameConcat / stage=combine
    chunks between them and the agg operand's input ``root``, ``opaque`` chunks that were executed earlier
    (method='auto' samples the first combine_size map chunks) and feed the tree as ready partial tuples."""
    __slots__ = ("agg", "root", "leaves", "interior", "opaque", "nodes")

    def __init__(self, agg, root, leaves, interior, opaque, nodes):
        self.agg, self.root, self.leaves, self.interior, self.opaque, self.nodes = agg, root, leaves, interior, opaque, nodes


def find_fused_trees(order, keep) -> List[FusedTree]:
    """``order``: the chunks about to run, topologically sorted; ``keep``: keys whose values must survive (result
    chunks).  A subtree qualifies when every node has exactly one consumer and is a stage=map / stage=combine
    DataFrameGroupByAgg or a DataFrameConcat -- i.e. when nothing but the stage=agg operand can observe the
    intermediate partial tuples."""
    if not (_ASYNC and hasattr(_kernels, "preagg_dense_batch_async")):
        return []
    pos = {c.key: i for i, c in enumerate(order)}
    nsucc: Dict[Any, int] = {}
    for c in order:
        for inp in c.op.inputs:
            nsucc[inp.key] = nsucc.get(inp.key, 0) + 1
    out = []
    cands = []
    for a in order:
        op = a.op
        name, st = type(op).__name__, _stage(op)
        if name == "DataFrameGroupByAgg" and st == "agg" and len(op.inputs) == 1 and op.inputs[0].key in pos:
            cands.append((a, op.inputs[0]))
        elif a.key in keep and ((name == "DataFrameGroupByAgg" and st == "combine") or name == "DataFrameConcat"):
            cands.append((None, a))      # the tree's root itself is a result chunk (sharded tree: the feed of the agg)
    for a, root in cands:
        leaves, interior, opaque, ok = [], [], [], True
        stack = [root]
        while stack:
            c = stack.pop()
            if c.key not in pos:
                opaque.append(c)
                continue
            name, st = type(c.op).__name__, _stage(c.op)
            if c is not root and (nsucc.get(c.key, 0) != 1 or c.key in keep):
                ok = False
                break
            if c is root and a is not None and (nsucc.get(c.key, 0) != 1 or c.key in keep):
                ok = False
                break
            if name == "DataFrameGroupByAgg" and st == "map":
                if getattr(c.op, "size_recorder_name", None) is not None or len(c.op.inputs) != 1:
                    ok = False
                    break
                leaves.append(c)
            elif (name == "DataFrameGroupByAgg" and st == "combine") or name == "DataFrameConcat":
                interior.append(c)
                stack.extend(reversed(c.op.inputs))
            else:
                ok = False
                break
        if ok and len(leaves) >= 2:
            nodes = sorted(leaves + interior, key=lambda c: pos[c.key])
            out.append(FusedTree(a, root, sorted(leaves, key=lambda c: pos[c.key]), interior, opaque, nodes))
    return out


class EagerBatch:
    """A band-level batch that was launched BEFORE the chunk graph was tiled (session.Session._execute): the
    map stage of every input chunk is needed under any plan, so its launch does not have to wait for the host
    to tile and walk the graph -- the GPU streams the band while the host builds the plan.  ``pieces`` is the
    partial tuple of the whole band (one piece per slot), ``covered`` maps the key of every input chunk that
    went into the launch to its frame."""
    __slots__ = ("plan", "op", "pieces", "covered", "used", "bound")

    def __init__(self, plan, op, pieces, covered, bound):
        self.plan, self.op, self.pieces, self.covered, self.bound = plan, op, pieces, covered, bound
        self.used = False


class EagerRef:
    """What ``ctx`` holds for a stage=map chunk whose input is covered by an eager batch: the chunk's own
    partial tuple does not exist (the band was aggregated as a whole).  A consumer that is not the fused tree
    (e.g. the shuffle branch) gets the real per-chunk result through ``materialize``."""
    __slots__ = ("eager", "op", "in_key")

    def __init__(self, eager, op, in_key):
        self.eager, self.op, self.in_key = eager, op, in_key

    def materialize(self):
        tmp = _MiniCtx({self.in_key: self.eager.covered[self.in_key]})
        saved, self.op.size_recorder_name = getattr(self.op, "size_recorder_name", None), None
        try:
            _execute_agg_map(tmp, self.op)
        finally:
            self.op.size_recorder_name = saved
        return tmp[self.op.outputs[0].key]


def exact_partial_size(ctx, key, ref) -> int:
    """exact size of a skipped map chunk's partial tuple (only asked for when the capacity bounds alone would
    send the plan to the shuffle branch): the chunk is aggregated for real"""
    v = ctx.get(key)
    if isinstance(v, EagerRef):
        v = ctx[key] = v.materialize()
    return sum(len(f) * 8 * (len(f.columns) + len(f.index_names)) for f in v)


class _MiniCtx(dict):
    def get_current_chunk(self):
        return None


def _mid_eligible(plan, sig) -> bool:
    """sum / mean / count shape: float64 value columns with SUM / COUNT only (int64 columns only counted)"""
    _, count_only, dts = sig
    for j, dt in enumerate(dts):
        if dt != _kernels.F64 and j not in count_only:
            return False
    return all(op in (_kernels.SUM, _kernels.COUNT) for _, op in plan.accs)


def _band_kernel(plan, sig, frames):
    """Which band-level batch kernel takes this shape: 'dense' (a handful of groups), 'mid' (up to ~2000 groups,
    sum / mean / count shapes) or None (general path).  The first chunk of an unknown shape is probed once
    (a strided key sample, one small host wait); afterwards the choice follows what the kernels reported."""
    key = plan.shape_key
    has_mid = hasattr(_kernels, "preagg_mid_batch_async")
    if key not in _probed and hasattr(_kernels, "estimate_distinct") and key not in _dense_rejected:
        _probed.add(key)
        if len(_probed) > 4096:
            _probed.clear()
        f0 = frames[0]
        if f0.physical_rows is not None and f0.physical_rows >= 4096:
            # (a pending row filter is not applied for the probe: the stored keys bound the cardinality)
            keys0 = [f0.physical(b) for b in plan.by] if f0.is_lazy else [f0.col(i) for i in plan.key_pos]
            sampled, uniq = _kernels.estimate_distinct(keys0)
            if uniq > 12:
                _dense_rejected.add(key)
            if uniq > 1400:
                _mid_rejected.add(key)
    if key not in _dense_rejected:
        return "dense"
    if has_mid and key not in _mid_rejected and _mid_eligible(plan, sig):
        return "mid"
    return None


def _row_program(plan, sig, frame):
    """(program, physical column names) of the fused pre-path for one chunk schema, or None when the shape does not
    fit: only float64 SUM / COUNT accumulators, the extra operands (columns that are read but not aggregated, the
    filter column) must fit into the index slots of the derived columns."""
    _, count_only, dts = sig
    if not _mid_eligible(plan, sig) or not hasattr(_kernels, "preagg_dense_batch_fused_async"):
        return None
    logical = [plan.used[j] for j in range(len(plan.used)) if j not in count_only]     # loaded logical columns
    if len(logical) > 8:
        return None
    filt, derived = frame.lazy_spec()
    derived = dict(derived)
    phys: List[Any] = [None if c in derived else c for c in logical]
    free = [i for i, c in enumerate(phys) if c is None]

    def slot_of(name):
        if name in phys:
            return phys.index(name)
        if not free or name in derived or name not in frame.columns:
            return None
        i = free.pop(0)
        phys[i] = name
        return i

    prog_logical = []
    for c in logical:
        if c not in derived:
            prog_logical.append(None)
            continue
        facs = []
        for name, alpha, beta in derived[c]:
            i = slot_of(name)
            if i is None:
                return None
            facs.append((i, alpha, beta))
        prog_logical.append(facs)
    row_filter = None
    if filt is not None:
        col, op, value = filt
        i = slot_of(col)
        if i is None:
            return None
        dt = frame.raw_data[frame.columns.index(col)].dtype
        row_filter = (i, op, _kernels.F64 if dt == torch.float64 else _kernels.I64, value)
    n_phys = max(i for i, c in enumerate(phys) if c is not None) + 1
    for i in range(n_phys):
        if phys[i] is None:
            phys[i] = plan.by[0]          # an unused slot below n_phys: any valid column, its values are never read
    # aggregated operands must be float64 (the kernel treats every loaded column as such)
    for i, name in enumerate(phys[:n_phys]):
        t = frame.raw_data[frame.columns.index(name)]
        if row_filter is not None and i == row_filter[0]:
            continue
        if t.dtype != torch.float64 and name != plan.by[0]:
            return None
    return _kernels.make_row_program(n_phys, prog_logical, row_filter), phys[:n_phys]


def _launch_band_fused(plan, sig, frames):
    """band-level batch with the row program (filter + derived columns evaluated inside the pre-aggregation pass)"""
    spec0 = frames[0].lazy_spec()
    for f in frames[1:]:
        if f.lazy_spec() != spec0:
            return None
    rp = _row_program(plan, sig, frames[0])
    if rp is None:
        return None
    prog, phys = rp
    keys_ok, count_only, dts = sig
    B = _kernels.dense_batch_max()
    dev = frames[0].device
    nk = len(plan.by)
    per_slot: List[List[DeviceFrame]] = [[] for _ in plan.slots]
    for g0 in range(0, len(frames), B):
        group = frames[g0:g0 + B]
        with _device_mod.deferred_uploads():      # host columns: nothing reads them before the launch below
            chunks = [([f.physical(b).data_ptr() for b in plan.by], [f.physical(c).data_ptr() for c in phys],
                       f.physical_rows) for f in group]
        bi, bf, fc = _kernels.preagg_dense_batch_fused_async(chunks, plan.accs, dts, nk, dev, prog)

        def replay(group=group, plan=plan):
            """not a low-cardinality band after all: lazy parts are materialised, general aggregator"""
            _dense_rejected.add(plan.shape_key)
            agg = _kernels.Aggregator(len(plan.by), [_kernels.dtype_code(group[0].col(i)) for i in plan.val_pos],
                                      list(plan.accs))
            try:
                for f in group:
                    agg.update([f.col(i) for i in plan.key_pos], [f.col(i) for i in plan.val_pos])
                return agg.result(sort=False, device=group[0].device)
            finally:
                agg.close()

        pend = Pending(bi, fc.n_int, bi.shape[1], replay, fblock=bf, layout=fc)
        src = ColumnSet.from_block(bi, bf, fc)
        cap = bi.shape[1]
        for s_, (cols, a0, a1) in enumerate(plan.slots):
            per_slot[s_].append(DeviceFrame.from_range(plan.by, cols, src, 0, nk, nk + a0, nk + a1, 0, cap,
                                                       pending=pend, slot=(a0, a1)))
    return per_slot


def _launch_band(plan, sig, frames):
    """one batch launch per mgb_dense_batch_max() chunks -> per slot the list of pieces, or None"""
    which = _band_kernel(plan, sig, frames)
    if which is None:
        return None
    fn = _kernels.preagg_dense_batch_async if which == "dense" else _kernels.preagg_mid_batch_async
    rejected = _dense_rejected if which == "dense" else _mid_rejected
    keys_ok, count_only, dts = sig
    if frames[0].is_lazy:
        fused = _launch_band_fused(plan, sig, frames) if which == "dense" else None
        if fused is not None:
            return fused
        # not a shape the fused pre-path takes: the lazy parts are materialised by the column accessors below
    chunks = []
    ck = (tuple(plan.key_pos), tuple(plan.val_pos), count_only)
    # chunks still in host memory: their uploads may complete when the loop ends (nothing reads the columns before the
    # launch below) -- unless a column accessor has lazy parts to materialise, which launches kernels on them
    defer = (any(f._home is not None for f in frames)
             and not any(f.is_lazy or f._pending is not None for f in frames))
    with _device_mod.deferred_uploads() if defer else contextlib.nullcontext():
        for f in frames:
            hit = f._ptrs_cache
            if hit is not None and hit[0] == ck:      # the column pointers of a resident chunk do not change
                chunks.append(hit[1])
                continue
            kp = [f.col(i).data_ptr() for i in plan.key_pos]
            vp = [None if j in count_only else f.col(i).data_ptr() for j, i in enumerate(plan.val_pos)]
            ent = [kp, vp, len(f)]          # (a list: the batch wrapper appends its ctypes arrays for the next call)
            if f._pending is None and f._src is None:
                f._ptrs_cache = (ck, ent)
            chunks.append(ent)
    B = _kernels.dense_batch_max()
    dev = frames[0].device
    nk = len(plan.by)
    per_slot: List[List[DeviceFrame]] = [[] for _ in plan.slots]
    for g0 in range(0, len(chunks), B):
        res = fn(chunks[g0:g0 + B], plan.accs, dts, nk, dev)
        if res is None:
            rejected.add(plan.shape_key)
            if g0 == 0:
                return None
            raise MarsGBError(6, "batch shape rejected after the first batch was launched")
        bi, bf, fc = res
        group = frames[g0:g0 + B]

        def replay(group=group, plan=plan, rejected=rejected):
            """the kernel reported -1 (more groups than it holds, ...): the band's chunks go through the general
            aggregator, and the shape is remembered"""
            rejected.add(plan.shape_key)
            agg = _kernels.Aggregator(len(plan.by), [_kernels.dtype_code(group[0].col(i)) for i in plan.val_pos],
                                      list(plan.accs))
            try:
                for f in group:
                    agg.update([f.col(i) for i in plan.key_pos], [f.col(i) for i in plan.val_pos])
                return agg.result(sort=False, device=group[0].device)
            finally:
                agg.close()

        pend = Pending(bi, fc.n_int, bi.shape[1], replay, fblock=bf, layout=fc)
        src = ColumnSet.from_block(bi, bf, fc)
        cap = bi.shape[1]
        for s_, (cols, a0, a1) in enumerate(plan.slots):
            per_slot[s_].append(DeviceFrame.from_range(plan.by, cols, src, 0, nk, nk + a0, nk + a1, 0, cap,
                                                       pending=pend, slot=(a0, a1)))
    return per_slot


def _band_frames_plan(op0, frames):
    """(plan, dtype signature) when all frames of a band share one low-cardinality-capable schema, else None"""
    for v in frames:
        if not isinstance(v, DeviceFrame) or v.raw_index is not None:
            return None
    try:
        plan = _map_plan(op0, frames[0])
    except NotImplementedError:
        return None
    sig = plan.for_dtypes(frames[0])
    if not sig[0]:
        return None
    if plan.shape_key in _dense_rejected and (plan.shape_key in _mid_rejected or not _mid_eligible(plan, sig)
                                              or not hasattr(_kernels, "preagg_mid_batch_async")):
        return None
    cols0 = frames[0].columns
    for f in frames[1:]:
        if (f.columns is not cols0 and f.columns != cols0) or plan.for_dtypes(f) is not sig:
            return None
    return plan, sig


def launch_eager_band(map_op, in_keys, frames) -> Optional[EagerBatch]:
    """Launch the band-level batch over the input chunks of a groupby-agg before its graph exists.  ``map_op``
    is a stage=map operand as the tiler will create it (same compiled functions)."""
    if not (_ASYNC and hasattr(_kernels, "preagg_dense_batch_async")) or len(frames) < 2:
        return None
    ps = _band_frames_plan(map_op, frames)
    if ps is None:
        return None
    plan, sig = ps
    per_slot = _launch_band(plan, sig, frames)
    if per_slot is None:
        return None
    cap = per_slot[0][0]._pending.cap
    bound = sum(cap * 8 * (len(cols) + len(plan.by)) for cols, _, _ in plan.slots)
    return EagerBatch(plan, map_op, per_slot, dict(zip(in_keys, frames)), bound)


def execute_fused_tree(ctx, ft: FusedTree, eager: Optional[EagerBatch] = None) -> bool:
    """Runs the fused part of the graph: ONE launch of the band-level batch kernel over the input chunks of all
    leaf map operands (per mgb_dense_batch_max() chunks) instead of n map launches and the merges of the
    combine tree; ``ctx[ft.root.key]`` receives the partial tuple the stage=agg operand expects (a virtual
    concat of the batch results and the opaque partial tuples).  Returns False -- nothing done -- when the
    shape is not a low-cardinality one or the chunks differ in schema: the caller then runs the operands one
    by one."""
    fn = getattr(_kernels, "preagg_dense_batch_async", None)
    if fn is None or not _ASYNC:
        return False
    op0 = ft.leaves[0].op
    for leaf in ft.leaves[1:]:
        if leaf.op.agg_funcs is not op0.agg_funcs and leaf.op.agg_funcs != op0.agg_funcs:
            return False
    if eager is not None and not eager.used:
        # the band was launched before tiling: the tree is served by that launch when its leaves plus the map
        # chunks that were skipped earlier (EagerRef) are exactly the chunks the launch covered
        refs = [ctx[c.key] for c in ft.opaque]
        if all(isinstance(v, EagerRef) and v.eager is eager for v in refs):
            keys = {leaf.op.inputs[0].key for leaf in ft.leaves} | {v.in_key for v in refs}
            if keys == set(eager.covered) and len(keys) == len(ft.leaves) + len(refs):
                eager.used = True
                ctx[ft.root.key] = tuple(PiecewiseFrame(p) for p in eager.pieces)
                return True
    frames = [ctx[leaf.op.inputs[0].key] for leaf in ft.leaves]
    ps = _band_frames_plan(op0, frames)
    if ps is None:
        return False
    plan, sig = ps
    per_slot = _launch_band(plan, sig, frames)
    if per_slot is None:
        return False
    for c in ft.opaque:
        v = ctx[c.key]
        if isinstance(v, EagerRef):
            v = ctx[c.key] = v.materialize()
        for s_, pcs in enumerate(_pieces_of(v)):
            per_slot[s_].extend(pcs)
    ctx[ft.root.key] = tuple(PiecewiseFrame(p) for p in per_slot)
    return True


def execute_groupby_agg(ctx, op):
    st = _stage(op)
    if st == "map":
        _execute_agg_map(ctx, op)
    elif st == "combine":
        _execute_agg_combine(ctx, op)
    elif st == "agg":
        _execute_agg_agg(ctx, op)
    else:
        raise ValueError("Aggregation operand not executable")


# ----------------------------------------------------------------------------------------------------
# DataFrameGroupByOperand: hash shuffle
# ----------------------------------------------------------------------------------------------------
def hash_split_frame(frame: DeviceFrame, on: Sequence[Any], n_partitions: int) -> List[DeviceFrame]:
    """`hash_dataframe_on(df, on=[...], size)` + `df.iloc[filter]` per reducer (mars/dataframe/utils.py:77-101)
    for a device-resident frame: the rows are hashed on the `on` COLUMNS with pandas' DataFrame semantics
    (mgb_partition_ids_frame) and split stably; every block keeps the original row labels as its index.  Shared
    by the hash-partition consumers outside the agg path: plain groupby (groupby/core.py:353-435 with `by`),
    the merge shuffle (merge/merge.py:93-128)."""
    on = list(on)
    _check_key_count(len(on))
    for c in on:
        if c not in frame.columns:
            raise NotImplementedError(f"shuffle column {c!r} is not a data column of the chunk")
    keys = [frame.column(c) for c in on]
    for c, t in zip(on, keys):
        if t.dtype != torch.int64:
            raise NotImplementedError(f"shuffle column {c!r} has dtype {t.dtype}: int64 columns only")
    n = len(frame)
    if frame.index is None:
        if frame.range_start is None:
            raise NotImplementedError("the chunk's row labels are unknown")
        index = [torch.arange(frame.range_start, frame.range_start + n, dtype=torch.int64, device=frame.device)]
        names = [None]
    else:
        index, names = list(frame.index), list(frame.index_names)
    cols = index + list(frame.data)
    if n == 0:
        outs, offsets = cols, [0] * (n_partitions + 1)
    else:
        if n_partitions <= 256 and hasattr(_kernels, "split_plan"):
            # the partition kernel with DataFrame-row hash semantics: no id array, no range-check synchronisation
            hist, off_dev = _kernels.split_plan(keys, n_partitions, 1)
            outs = [torch.empty_like(c) for c in cols]
            _kernels.split_scatter(keys, n_partitions, hist, cols, outs=outs, mode=1)
            offsets = off_dev.cpu().tolist()
        else:
            pid = _kernels.partition_ids_frame(keys, n_partitions)
            outs, offsets = _kernels.partition_scatter(pid, n_partitions, cols)
    nk = len(index)
    src = ColumnSet(outs)
    return [DeviceFrame.from_range(names, frame.columns, src, 0, nk, nk, len(cols), offsets[r], offsets[r + 1])
            for r in range(n_partitions)]


def _execute_by_shuffle_map(ctx, op, by):
    """plain groupby (apply / transform follow on the reference executors): shuffle of the raw chunk on its `by`
    columns (reference: execute_map with on = by, groupby/core.py:353-435)"""
    if callable(by) or any(not isinstance(b, (str, int, tuple)) for b in by):
        raise NotImplementedError("`by` must be column labels on the GPU path")
    chunk = op.outputs[0]
    value = ctx[op.inputs[0].key]
    if isinstance(value, tuple):
        raise NotImplementedError("`by` shuffles of partial tuples are outside the GPU path")
    frame = _as_device_frame(value)
    if getattr(op, "level", None) is not None:
        raise NotImplementedError("level= shuffles are outside the GPU path")
    blocks = hash_split_frame(frame, by, int(op.shuffle_size))
    mapper_index = ctx.get_current_chunk().index
    for r, block in enumerate(blocks):
        reducer_index = (r, chunk.index[1]) if getattr(op, "is_dataframe_obj", True) else (r,)
        ctx[chunk.key, reducer_index] = (mapper_index, block)


def execute_merge_align(ctx, op):
    """DataFrameMergeAlign (mars/dataframe/merge/merge.py:93-147): map = hash split of a chunk on its
    `shuffle_on` columns, one block per reducer incl. empty ones, payload (mapper_id, chunk index, block);
    reduce = row concat of the blocks of one side, ordered by mapper row."""
    st = _stage(op)
    chunk = op.outputs[0]
    if st == "map":
        on = op.shuffle_on
        if on is None:
            raise NotImplementedError("merge shuffles on the index stay on the reference executor")
        on = list(on) if isinstance(on, (list, tuple)) else [on]
        frame = _as_device_frame(ctx[op.inputs[0].key])
        blocks = hash_split_frame(frame, on, int(op.index_shuffle_size))
        idx = ctx.get_current_chunk().index
        for r, block in enumerate(blocks):
            ctx[chunk.key, (r, chunk.index[1])] = (op.mapper_id, idx, block)
        return
    for i, out in enumerate(op.outputs):
        got = {pidx: data for mapper_id, pidx, data in op.iter_mapper_data(ctx, skip_none=True) if mapper_id == i}
        res = [got[k] for k in sorted(got) if got[k] is not None]
        ctx[out.key] = _concat_frames([_as_device_frame(v, index_as_keys=True) for v in res])


class DeferredSplit:
    """One mapper chunk of a hash shuffle after pass 1 of the partition kernel (mgb_split_plan): the block sizes
    are on the device, the rows have not moved.  The multi-GPU exchange (parallel.Exchange, push transport)
    decides from the sizes of ALL mappers where every block goes and lets pass 2 write each row once, straight
    into the receive arena of the GPU that runs its reducer; ``emit`` is the single-process form (blocks of local
    columns under the reference's ctx keys, groupby/core.py:429-435)."""

    MODE = 4      # kernels.SPLIT_LONG_RUNS: Index-hash semantics, 8192-row tiles

    def __init__(self, chunk, op, mapper_index, keys, cols, R, hist, offsets, spans, nk, is_tuple):
        self.chunk, self.op, self.mapper_index = chunk, op, mapper_index
        self.keys, self.cols, self.R, self.hist, self.offsets = keys, cols, R, hist, offsets
        self.spans, self.nk, self.is_tuple = spans, nk, is_tuple      # spans: (slot, index_names, columns, c0, c1)
        self.n_slots = len(spans)

    def key_of(self, r):
        chunk, op = self.chunk, self.op
        return (chunk.key, (r, chunk.index[1]) if getattr(op, "is_dataframe_obj", True) else (r,))

    def schema(self):
        return dict(index_names=self.spans[0][1], nk=self.nk,
                    slots=[(columns, [str(self.cols[c].dtype) for c in range(c0, c1)]) for _, _, columns, c0, c1 in self.spans])

    def entry(self, frames):
        if self.is_tuple:
            return (self.mapper_index, tuple(frames) + (False,))
        return (self.mapper_index, frames[0])

    def emit(self, ctx):
        """pass 2 into local columns"""
        outs = [torch.empty_like(c) for c in self.cols]
        _kernels.split_scatter(self.keys, self.R, self.hist, self.cols, outs=outs, mode=self.MODE)
        offsets = self.offsets.cpu().tolist()
        src = ColumnSet(outs)
        for r in range(self.R):
            a, b = offsets[r], offsets[r + 1]
            ctx[self.key_of(r)] = self.entry([DeviceFrame.from_range(names, columns, src, 0, self.nk, d0, d1, a, b)
                                              for _, names, columns, d0, d1 in self.spans])


def _execute_shuffle_map(ctx, op):
    """stage=map: partition id = pandas hash of the group keys % shuffle_size; one block per reducer,
    empty ones included (reference: execute_map, groupby/core.py:353-435; contract shuffle.py:124-130)."""
    by = getattr(op, "by", None)
    if isinstance(by, (list, tuple)) or callable(by):
        return _execute_by_shuffle_map(ctx, op, by)
    chunk = op.outputs[0]
    value = ctx[op.inputs[0].key]
    is_tuple = isinstance(value, tuple)
    frames = _partials_of(value)
    R = int(op.shuffle_size)
    mapper_index = ctx.get_current_chunk().index
    # frames sharing their index are hashed and split together (reference hashes once per distinct index)
    groups: List[List[int]] = []
    for i, f in enumerate(frames):
        for g in groups:
            if frames[g[0]].same_index(f):
                g.append(i)
                break
        else:
            groups.append([i])
    if (getattr(ctx, "defer_split", False) and len(groups) == 1 and R <= 256 and hasattr(_kernels, "split_plan")
            and len(frames[0].index) >= 1):
        # multi-GPU run with a push transport: sizes now, rows when the exchange knows where they go
        base = frames[0]
        cols = list(base.index)
        spans, pos = [], len(cols)
        for i, f in enumerate(frames):
            cols.extend(f.data)
            spans.append((i, f.index_names, f.columns, pos, pos + len(f.columns)))
            pos += len(f.columns)
        # planned for stores that cross NVLink (long runs); emit() / the exchange scatter with the same mode
        hist, offsets = _kernels.split_plan(base.index, R, DeferredSplit.MODE)
        ctx[chunk.key, "deferred"] = DeferredSplit(chunk, op, mapper_index, list(base.index), cols, R, hist, offsets,
                                                   spans, len(base.index), is_tuple)
        return
    blocks: List[List[Optional[DeviceFrame]]] = [[None] * len(frames) for _ in range(R)]
    for g in groups:
        base = frames[g[0]]
        cols = list(base.index)
        for i in g:
            cols.extend(frames[i].data)
        if hasattr(_kernels, "hash_partition"):     # hash + split in one call, rows scattered to their blocks
            outs, offsets = _kernels.hash_partition(base.index, R, cols)
        else:
            pid = _kernels.partition_ids(base.index, R)
            outs, offsets = _kernels.partition_scatter(pid, R, cols)
        nk = len(base.index)
        # R x K blocks per chunk: row ranges of the scattered columns, sliced only if somebody asks for tensors
        src = ColumnSet(outs)
        spans, pos = [], nk
        for i in g:
            ncol = len(frames[i].columns)
            spans.append((i, frames[i].index_names, frames[i].columns, pos, pos + ncol))
            pos += ncol
        for r in range(R):
            a, b = offsets[r], offsets[r + 1]
            row = blocks[r]
            for i, names, columns, d0, d1 in spans:
                row[i] = DeviceFrame.from_range(names, columns, src, 0, nk, d0, d1, a, b)
    for r in range(R):
        reducer_index = (r, chunk.index[1]) if getattr(op, "is_dataframe_obj", True) else (r,)
        if is_tuple:
            ctx[chunk.key, reducer_index] = (mapper_index, tuple(blocks[r]) + (False,))
        else:
            ctx[chunk.key, reducer_index] = (mapper_index, blocks[r][0])


def _concat_frames(frames: Sequence[DeviceFrame], shared_keys=None) -> DeviceFrame:
    first = frames[0]
    keys = shared_keys
    if keys is None and first.index is not None:
        keys = [_kernels.concat_columns([f.index[j] for f in frames]) for j in range(len(first.index))]
    data = [_kernels.concat_columns([f.data[j] for f in frames]) for j in range(len(first.columns))]
    return DeviceFrame(first.index_names, keys, first.columns, data, sum(len(f) for f in frames))


def _concat_partial_tuples(items: Sequence[Sequence[Any]]) -> Tuple[PiecewiseFrame, ...]:
    """slot-wise row concatenation of partial tuples -- virtual: the pieces stay where they are and the
    consuming combine / agg stage reads them in place (no copy kernel)."""
    n_slots = len(items[0])
    return tuple(PiecewiseFrame([it[n] for it in items]) for n in range(n_slots))


def _execute_shuffle_reduce(ctx, op):
    """stage=reduce: gather mapper blocks ordered by mapper index, concat slot-wise
    (reference: execute_reduce, groupby/core.py:437-485)."""
    chunk = op.outputs[0]
    input_idx_to_df = dict(op.iter_mapper_data(ctx))
    res = [input_idx_to_df[i] for i in sorted(input_idx_to_df.keys()) if input_idx_to_df[i] is not None]
    if isinstance(res[0], tuple):
        if res[0][-1]:
            raise NotImplementedError("deliver_by shuffles are outside the agg path")
        items = [[_as_device_frame(v, index_as_keys=True) for v in it[:-1]] for it in res]
        r: Any = _concat_partial_tuples(items)
        if chunk.index_value is not None and getattr(chunk.index_value, "name", None) is not None:
            for f in r:
                for piece in f.pieces:
                    if len(piece.index_names) == 1:
                        piece.index_names = [chunk.index_value.name]
    else:
        r = _concat_frames([_as_device_frame(v, index_as_keys=True) for v in res])
    ctx[chunk.key] = r


def execute_groupby_operand(ctx, op):
    st = _stage(op)
    if st == "map":
        _execute_shuffle_map(ctx, op)
    elif st == "reduce":
        _execute_shuffle_reduce(ctx, op)
    else:
        raise NotImplementedError("plain groupby objects (apply/transform) are outside the GPU agg path")


# ----------------------------------------------------------------------------------------------------
# sort=True shuffle: PSRS sample / pivots / range split (reference: mars/dataframe/groupby/sort.py)
# ----------------------------------------------------------------------------------------------------
def execute_psrs_sample(ctx, op):
    """DataFramePSRSGroupbySample.execute (sort.py:70-97): n regular samples of the key-sorted first partial
    frame.  Only the sampled KEYS matter downstream (the pivot operand looks at `.index` alone), so the device
    work is: argsort the partial's key tuples, gather the n sampled positions, n x n_keys values to the host."""
    value = ctx[op.inputs[0].key]
    a = _partials_of(value)[0]
    n = int(op.n_partition)
    rows = len(a)
    names = a.index_names
    if rows == 0:
        keys_host = [np.empty(0, dtype=np.int64) for _ in a.index]
    else:
        perm = _kernels.argsort_keys(a.index)
        if rows < n:   # the reference repeats the frame num times and re-sorts: every key num times, in order
            num = n // rows + 1
            total = rows * num
            w = total * 1.0 / (n + 1)
            slc = np.linspace(max(w - 1, 0), total - 1, num=n, endpoint=False).astype(int) // num
        else:
            w = rows * 1.0 / (n + 1)
            slc = np.linspace(max(w - 1, 0), rows - 1, num=n, endpoint=False).astype(int)
        pos = torch.from_numpy(slc.astype(np.int32)).to(perm.device)
        take = _kernels.gather_rows([perm.to(torch.int64)], pos)[0].to(torch.int32)
        keys_host = _to_host(_kernels.gather_rows(list(a.index), take))
    if len(keys_host) == 1:
        index = pd.Index(keys_host[0], name=names[0] if names else None)
    else:
        index = pd.MultiIndex.from_arrays(keys_host, names=names)
    ctx[op.outputs[-1].key] = pd.DataFrame(index=index)


def execute_concat_pivot(ctx, op):
    """DataFrameGroupbyConcatPivot.execute (sort.py:44-67).  Host logic on p*p sample keys (metadata-sized):
    same statements as the reference, on the key indexes of the samples."""
    n_inputs = len(op.inputs)
    inputs = [ctx[c.key] for c in op.inputs if len(ctx[c.key]) > 0]
    a = pd.concat(inputs, axis=0).sort_index()
    index = a.index.drop_duplicates()
    p = len(inputs)
    if len(index) < p:
        num = p // len(index) + 1
        index = index.append([index] * (num - 1))
    index = index.sort_values()
    slc = np.linspace(p - 1, len(index) - 1, num=n_inputs - 1, endpoint=False).astype(int)
    ctx[op.outputs[-1].key] = index.values[slc]


def _pivot_columns(pivots, n_keys: int, device) -> List[torch.Tensor]:
    if n_keys == 1:
        cols = [np.asarray(pivots, dtype=np.int64)]
    else:
        cols = [np.asarray([t[j] for t in pivots], dtype=np.int64) for j in range(n_keys)]
    return [torch.from_numpy(np.ascontiguousarray(c)).to(device) for c in cols]


def _execute_sort_shuffle_map(ctx, op):
    """DataFrameGroupbySortShuffle._execute_map (sort.py:113-135): partition i = keys in (pivot[i-1], pivot[i]].
    The reference slices key-sorted frames by label; here every partial row finds its partition by a binary
    search over the pivots (mgb_range_partition_ids) and the rows are split stably -- the blocks hold the same
    row sets (their order inside a block differs, the stage=agg operand sorts its output anyway)."""
    value, pivots = ctx[op.inputs[0].key], ctx[op.inputs[1].key]
    frames = _partials_of(value)
    out = op.outputs[0]
    R = int(op.n_partition)
    base = frames[0]
    for f in frames[1:]:
        if not base.same_index(f):
            raise NotImplementedError("sort shuffle expects the partial frames of a tuple to share their index")
    cols = list(base.index)
    for f in frames:
        cols.extend(f.data)
    nk = len(base.index)
    if len(base) == 0:
        outs, offsets = cols, [0] * (R + 1)
    else:
        pid = _kernels.range_partition_ids(base.index, _pivot_columns(pivots, nk, base.device))
        outs, offsets = _kernels.partition_scatter(pid, R, cols)
    for r in range(R):
        a, b = offsets[r], offsets[r + 1]
        keys_r = [t[a:b] for t in outs[:nk]]
        pos = nk
        block = []
        for f in frames:
            ncol = len(f.columns)
            block.append(DeviceFrame(f.index_names, keys_r, f.columns, [t[a:b] for t in outs[pos:pos + ncol]], b - a))
            pos += ncol
        ctx[out.key, (r, 0)] = tuple(block)


def _execute_sort_shuffle_reduce(ctx, op):
    """DataFrameGroupbySortShuffle._execute_reduce (sort.py:137-151): slot-wise (virtual) concat + `by`"""
    raw_inputs = [v for v in op.iter_mapper_data(ctx) if v is not None]
    items = [[_as_device_frame(v, index_as_keys=True) for v in it if not isinstance(v, (bool, list))]
             for it in raw_inputs]
    ctx[op.outputs[0].key] = _concat_partial_tuples(items) + (getattr(op, "by", None),)


def execute_sort_shuffle(ctx, op):
    if _stage(op) == "map":
        _execute_sort_shuffle_map(ctx, op)
    else:
        _execute_sort_shuffle_reduce(ctx, op)


# ----------------------------------------------------------------------------------------------------
# concat / data movement
# ----------------------------------------------------------------------------------------------------
def execute_concat(ctx, op):
    """tuple branch of DataFrameConcat.execute (merge/concat.py:339-348): tree glue between levels."""
    inputs = [ctx[c.key] for c in op.inputs]
    if not inputs or not isinstance(inputs[0], tuple):
        raise NotImplementedError("DataFrameConcat of plain frames stays on the reference executor")
    if any(isinstance(f, (pd.DataFrame, pd.Series)) for v in inputs for f in v):
        # partial tuples produced by the reference executor (a map stage outside the GPU envelope): its own
        # concat keeps them on the host; a later combine / agg inside the envelope uploads them itself
        raise NotImplementedError("partial tuples of pandas frames are concatenated by the reference executor")
    items = [[PiecewiseFrame(pieces) for pieces in _pieces_of(v)] for v in inputs]
    ctx[op.outputs[0].key] = _concat_partial_tuples(items)


def execute_filter(ctx, op):
    """boolean-mask row selection `df[df[col] <op> value]` (reference: DataFrameIndex with a mask,
    mars/dataframe/indexing/getitem.py).  Lazy: the rows are dropped by whoever reads the chunk -- inside the
    pre-aggregation pass when the fused pre-path applies (TPC-H Q1), by a stable compaction otherwise."""
    frame = _as_device_frame(ctx[op.inputs[0].key], lazy=_LAZY_H2D)
    if op.mask_col not in frame.columns:
        raise NotImplementedError(f"filter column {op.mask_col!r} is not a data column")
    ctx[op.outputs[0].key] = frame.with_filter(op.mask_col, op.mask_op, op.mask_value)


def execute_setitem(ctx, op):
    """`df[name] = <product of affine factors of columns>` (reference: DataFrameSetitem + the arithmetic operands
    that feed it, mars/dataframe/indexing/setitem.py, mars/dataframe/arithmetic).  Lazy, like the filter."""
    frame = _as_device_frame(ctx[op.inputs[0].key], lazy=_LAZY_H2D)
    ctx[op.outputs[0].key] = frame.with_column(op.target, op.factors)


def execute_to_gpu(ctx, op):
    """DataFrameToGPU.execute (base/to_gpu.py:28-35).  The columns are uploaded when their first consumer asks
    for them, so a column no kernel reads (an int64 column that is only counted) never crosses PCIe."""
    ctx[op.outputs[0].key] = _as_device_frame(ctx[op.inputs[0].key], lazy=_LAZY_H2D)


def execute_to_cpu(ctx, op):
    v = ctx[op.inputs[0].key]
    ctx[op.outputs[0].key] = v.to_pandas() if isinstance(v, DeviceFrame) else v


def execute_datasource(ctx, op):
    ctx[op.outputs[0].key] = op.data


def register_gpu_executors():
    """Install the GPU executors on the host-mirror operand classes."""
    from . import dataframe as md

    md.DataFrameGroupByAgg.register_executor(execute_groupby_agg)
    md.DataFrameGroupByOperand.register_executor(execute_groupby_operand)
    md.DataFrameConcat.register_executor(execute_concat)
    md.DataFrameToGPU.register_executor(execute_to_gpu)
    md.DataFrameToCPU.register_executor(execute_to_cpu)
    md.DataFrameDataSource.register_executor(execute_datasource)
    md.DataFramePSRSGroupbySample.register_executor(execute_psrs_sample)
    md.DataFrameGroupbyConcatPivot.register_executor(execute_concat_pivot)
    md.DataFrameGroupbySortShuffle.register_executor(execute_sort_shuffle)
    md.DataFrameMergeAlign.register_executor(execute_merge_align)
    md.DataFrameFilter.register_executor(execute_filter)
    md.DataFrameSetitem.register_executor(execute_setitem)
