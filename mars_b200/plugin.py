"""Registration of the GPU executors on the REAL reference classes (when ``mars`` is importable).

This module is what a Mars deployment loads into every GPU-band sub-process so that the operands of the
groupby-aggregation path run on libmarsgb.so instead of pandas/cudf.  Loading mechanisms offered by the
reference (all unchanged host code):

  * service YAML ``third_party_modules: [mars_b200.plugin]`` or env ``MARS_LOAD_MODULES=mars_b200.plugin``
    (mars/deploy/utils.py:193-219, mars/oscar/backends/pool.py:712-715) -- the module is imported in every
    pool process; importing it calls ``register_on_mars()``;
  * setuptools entry point group ``mars_extensions`` -> ``init`` (mars/core/entrypoints.py:24-42);
  * per-run ``extra_config={"operand_executors": {...}}`` under the checked test processor
    (mars/services/subtask/worker/tests/subtask_processor.py:58-62).

The executors are the same functions that serve the host mirror (``mars_b200.executors``): they only use
the operand protocol (``op.stage``, ``op.inputs/outputs``, ``op.groupby_params``, ``op.pre_funcs/agg_funcs/
post_funcs``, ``op.shuffle_size``, ``op.iter_mapper_data(ctx)``, ``ctx.get_current_chunk()``).

Fallback contract (SURVEY.md section 8b "Errors"): an executor that meets something outside the GPU envelope
(object keys, custom reductions, CPU bands) raises NotImplementedError
*before touching data*; ``_with_reference_fallback`` then calls the class's own ``execute`` -- the
reference's pandas path -- explicitly, because the dispatcher itself does not fall back to it
(mars/core/operand/core.py:501-509).  Ops placed on CPU bands (``op.gpu`` false) always take the reference
path: this plugin only claims GPU bands.
"""
from __future__ import annotations

import functools

from . import executors


def _to_host_value(v):
    """device-resident value of the GPU executors -> what the reference executors expect (pandas)"""
    from .device import DeviceFrame, PiecewiseFrame
    from .executors import DeviceResult

    if isinstance(v, (DeviceFrame, PiecewiseFrame, DeviceResult)):
        return v.to_pandas()
    if isinstance(v, tuple):
        return tuple(_to_host_value(x) for x in v)
    return v


def _inputs_to_host(ctx, op):
    """An operand that falls back to the reference executor may be fed by operands that ran on the GPU (e.g.
    a stage=agg with a custom reduction after our concat): its inputs are converted in place first."""
    keys = [inp.key for inp in (op.inputs or [])]
    it = getattr(op, "iter_mapper_keys", None)
    if it is not None and getattr(getattr(op, "stage", None), "name", None) == "reduce":
        try:
            keys += list(it())
        except Exception:  # noqa: BLE001
            pass
    for k in keys:
        try:
            v = ctx[k]
        except KeyError:
            continue
        h = _to_host_value(v)
        if h is not v:
            ctx[k] = h


def _with_reference_fallback(cls, fn, gpu_only=True):
    reference_execute = cls.execute

    @functools.wraps(fn)
    def wrapped(ctx, op):
        if gpu_only and not getattr(op, "gpu", False):
            return reference_execute(ctx, op)
        try:
            return fn(ctx, op)
        except NotImplementedError:
            _inputs_to_host(ctx, op)
            # the reference's GPU mode means cudf (`xdf = cudf if op.gpu else pd`, aggregation.py:1045); the
            # inputs were just handed over as pandas, so the operand runs the reference's pandas path
            gpu = getattr(op, "gpu", None)
            try:
                if gpu:
                    op.gpu = False
                return reference_execute(ctx, op)
            finally:
                if gpu:
                    op.gpu = gpu

    return wrapped


def register_on_mars(gpu_only: bool = True):
    """Install the executors on mars.dataframe's operand classes.  Returns the list of patched classes."""
    from mars.dataframe.base.to_cpu import DataFrameToCPU
    from mars.dataframe.base.to_gpu import DataFrameToGPU
    from mars.dataframe.groupby.aggregation import DataFrameGroupByAgg
    from mars.dataframe.groupby.core import DataFrameGroupByOperand
    from mars.dataframe.groupby.sort import (DataFrameGroupbyConcatPivot, DataFrameGroupbySortShuffle,
                                             DataFramePSRSGroupbySample)
    from mars.dataframe.merge.concat import DataFrameConcat
    from mars.dataframe.merge.merge import DataFrameMergeAlign

    # the GPU-level chunk store + the serializer of the device types: partial tuples (DeviceFrame / PiecewiseFrame)
    # cross subtask boundaries by handle and worker boundaries through the reference's own transfer code
    from . import storage

    storage.register_on_mars()
    # FINAL result chunks are handed out as pandas by default: the client that calls fetch() has no GPU, and
    # chunk metas are computed from the frame (subtask/utils.py:62-68).  MGB_MARS_DEVICE_RESULTS=1 keeps large
    # result chunks in HBM as DeviceResult (frame-like; `.to_pandas()` reads it back), as the reference's GPU mode
    # does with cudf frames (test_groupby_execution.py:788-793).
    import os

    executors._DEVICE_RESULTS = os.environ.get("MGB_MARS_DEVICE_RESULTS", "0") not in ("0", "")

    table = [
        (DataFramePSRSGroupbySample, executors.execute_psrs_sample),
        (DataFrameGroupbyConcatPivot, executors.execute_concat_pivot),
        (DataFrameGroupbySortShuffle, executors.execute_sort_shuffle),
        (DataFrameGroupByAgg, executors.execute_groupby_agg),
        (DataFrameGroupByOperand, executors.execute_groupby_operand),
        (DataFrameConcat, executors.execute_concat),
        (DataFrameToGPU, executors.execute_to_gpu),
        (DataFrameToCPU, executors.execute_to_cpu),
        (DataFrameMergeAlign, executors.execute_merge_align),
    ]
    for cls, fn in table:
        cls.register_executor(_with_reference_fallback(cls, fn, gpu_only=gpu_only))
    return [cls for cls, _ in table]


def unregister_on_mars():
    from mars.dataframe.base.to_cpu import DataFrameToCPU
    from mars.dataframe.base.to_gpu import DataFrameToGPU
    from mars.dataframe.groupby.aggregation import DataFrameGroupByAgg
    from mars.dataframe.groupby.core import DataFrameGroupByOperand
    from mars.dataframe.groupby.sort import (DataFrameGroupbyConcatPivot, DataFrameGroupbySortShuffle,
                                             DataFramePSRSGroupbySample)
    from mars.dataframe.merge.concat import DataFrameConcat

    from mars.dataframe.merge.merge import DataFrameMergeAlign

    for cls in (DataFrameGroupByAgg, DataFrameGroupByOperand, DataFrameConcat, DataFrameToGPU, DataFrameToCPU,
                DataFramePSRSGroupbySample, DataFrameGroupbyConcatPivot, DataFrameGroupbySortShuffle, DataFrameMergeAlign):
        cls.unregister_executor()


def init():
    """entry point for the ``mars_extensions`` group"""
    register_on_mars()


try:  # imported through third_party_modules / MARS_LOAD_MODULES inside a Mars pool process
    import mars  # noqa: F401
except ImportError:  # the host mirror is used instead (mars_b200.dataframe)
    pass
else:
    register_on_mars()
