"""User API and operands of the chunked DataFrame groupby-aggregation path (host side, pure Python).

Mirror of the reference's surface for this path -- same operand names, stages, tiling rules and chunk metas:

  md.DataFrame(raw, chunk_size=...)          mars/dataframe/initializer.py, datasource/dataframe.py
  df.to_gpu()                                mars/dataframe/base/to_gpu.py:19-35
  df.groupby(by, sort=, as_index=)           mars/dataframe/groupby/core.py:511-534
  groupby[...].agg(func, method=, combine_size=)   mars/dataframe/groupby/aggregation.py:1287-1354
  DataFrameGroupByAgg.tile (tree / shuffle / auto) mars/dataframe/groupby/aggregation.py:428-949
  DataFrameGroupByOperand map/reduce, DataFrameShuffleProxy, DataFrameConcat

Only *graph construction* lives here; nothing in this module computes on data.  The operands have no CPU
``execute``: the GPU executors of ``mars_b200.executors`` are registered on them through
``Operand.register_executor`` -- the same seam the real reference offers (see mars_b200.plugin).
"""
from __future__ import annotations

import itertools
from collections import OrderedDict
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd

from .core import (ChunkData, IndexValue, MapReduceOperand, Operand, OperandStage, ShuffleProxy,
                   TileableData, new_key)
from .reduction import ReductionSteps, check_supported, compile_reduction


class _Options:
    """reference: mars/config.py:363-371"""
    combine_size = 4
    chunk_store_limit = 128 * 1024 ** 2
    chunk_size = None


options = _Options()


# ----------------------------------------------------------------------------------------------------
# operands
# ----------------------------------------------------------------------------------------------------
class DataFrameDataSource(Operand):
    """Holds one row-slice of the user's pandas frame (host memory)."""


class DataFrameToGPU(Operand):
    """Host -> device copy of a chunk (reference: DataFrameToGPU.execute, to_gpu.py:28-35)."""


class DataFrameToCPU(Operand):
    """Device -> host copy of a chunk (reference: dataframe/base/to_cpu.py)."""


class DataFrameShuffleProxy(ShuffleProxy):
    """reference: mars/dataframe/operands.py:461-467"""


class DataFrameConcat(Operand):
    """Row-wise concatenation; on this path only the tuple-of-partials branch (merge/concat.py:339-348)."""


class DataFramePSRSGroupbySample(Operand):
    """regular samples of a key-sorted partial (reference: groupby/sort.py:70-97); field n_partition"""


class DataFrameGroupbyConcatPivot(Operand):
    """gathers the samples and picks p - 1 pivots (reference: groupby/sort.py:36-67)"""


class DataFrameGroupbySortShuffle(MapReduceOperand):
    """range shuffle of groupby(sort=True): map = split by pivots, reduce = gather (groupby/sort.py:100-167).

    fields: by, n_partition, shuffle_size"""

    def __init__(self, by=None, n_partition=None, shuffle_size=None, **kw):
        super().__init__(by=by, n_partition=n_partition, shuffle_size=shuffle_size, **kw)


class DataFrameGroupByOperand(MapReduceOperand):
    """groupby operand; stage map/reduce implement the hash shuffle (groupby/core.py:353-485).

    fields: by, level, as_index, sort, group_keys, shuffle_size"""

    def __init__(self, by=None, level=None, as_index=True, sort=True, group_keys=None, shuffle_size=None, **kw):
        super().__init__(by=by, level=level, as_index=as_index, sort=sort, group_keys=group_keys,
                         shuffle_size=shuffle_size, **kw)

    @property
    def is_dataframe_obj(self):
        return True

    @property
    def groupby_params(self):
        return dict(by=self.by, level=self.level, as_index=self.as_index, sort=self.sort,
                    group_keys=self.group_keys)


class DataFrameFilter(Operand):
    """row selection by a comparison mask, `df[df[col] <op> value]` (reference: DataFrameIndex with a boolean
    mask, mars/dataframe/indexing/getitem.py).  fields: mask_col, mask_op ("<=", "<", ">=", ">", "==", "!="),
    mask_value"""

    def __init__(self, mask_col=None, mask_op=None, mask_value=None, **kw):
        super().__init__(mask_col=mask_col, mask_op=mask_op, mask_value=mask_value, **kw)


class DataFrameSetitem(Operand):
    """`df[target] = expression` where the expression is a product of affine factors of columns (reference:
    DataFrameSetitem fed by arithmetic operands, mars/dataframe/indexing/setitem.py).  fields: target, factors
    [(column, alpha, beta)]"""

    def __init__(self, target=None, factors=None, **kw):
        super().__init__(target=target, factors=factors, **kw)


class ColumnExpr:
    """Lazy column expression of the mirror API: a left-to-right product of affine factors alpha + beta * column --
    enough for `price * (1 - discount) * (1 + tax)`; comparisons of a plain column with a scalar give a row mask."""

    def __init__(self, df, factors):
        self.df, self.factors = df, list(factors)

    def _plain(self):
        return len(self.factors) == 1 and self.factors[0][1] == 0.0 and self.factors[0][2] == 1.0

    def _affine(self, mul, add):
        if len(self.factors) != 1:
            raise NotImplementedError("scalar arithmetic on a product is outside the mirror's expression form")
        c, a, b = self.factors[0]
        return ColumnExpr(self.df, [(c, a * mul + add, b * mul)])

    def __mul__(self, other):
        if isinstance(other, ColumnExpr):
            if other.df is not self.df:
                raise NotImplementedError("columns of two frames")
            if len(self.factors) + len(other.factors) > 3:
                raise NotImplementedError("more than three factors")
            return ColumnExpr(self.df, self.factors + other.factors)
        return self._affine(float(other), 0.0)

    __rmul__ = __mul__

    def __add__(self, other):
        return self._affine(1.0, float(other))

    __radd__ = __add__

    def __sub__(self, other):
        return self._affine(1.0, -float(other))

    def __rsub__(self, other):
        return self._affine(-1.0, float(other))

    def __neg__(self):
        return self._affine(-1.0, 0.0)

    def _mask(self, op, value):
        if not self._plain():
            raise NotImplementedError("row masks compare a stored column with a scalar")
        return RowMask(self.df, self.factors[0][0], op, value)

    def __le__(self, v):
        return self._mask("<=", v)

    def __lt__(self, v):
        return self._mask("<", v)

    def __ge__(self, v):
        return self._mask(">=", v)

    def __gt__(self, v):
        return self._mask(">", v)


class RowMask:
    def __init__(self, df, column, op, value):
        self.df, self.column, self.op, self.value = df, column, op, value


class DataFrameMergeAlign(MapReduceOperand):
    """hash shuffle of the two sides of a merge on their join columns (reference: merge/merge.py:60-147).

    fields: shuffle_on, index_shuffle_size, mapper_id"""

    def __init__(self, shuffle_on=None, index_shuffle_size=None, **kw):
        super().__init__(shuffle_on=shuffle_on, index_shuffle_size=index_shuffle_size, **kw)


class DataFrameGroupByAgg(Operand):
    """groupby aggregation operand with stages map / combine / agg (aggregation.py:164-1284)."""

    def __init__(self, raw_func=None, raw_func_kw=None, func=None, func_rename=None, raw_groupby_params=None,
                 groupby_params=None, method="auto", use_inf_as_na=False, combine_size=None,
                 chunk_store_limit=None, pre_funcs=None, agg_funcs=None, post_funcs=None, index_levels=None,
                 size_recorder_name=None, **kw):
        super().__init__(raw_func=raw_func, raw_func_kw=raw_func_kw or {}, func=func, func_rename=func_rename,
                         raw_groupby_params=raw_groupby_params, groupby_params=groupby_params, method=method,
                         use_inf_as_na=use_inf_as_na, combine_size=combine_size,
                         chunk_store_limit=chunk_store_limit, pre_funcs=pre_funcs, agg_funcs=agg_funcs,
                         post_funcs=post_funcs, index_levels=index_levels,
                         size_recorder_name=size_recorder_name, tileable_op_key=None, **kw)

    # ---- function normalisation (reference: normalize_reduction_funcs, reduction/aggregation.py:959-1006)
    def _normalize(self):
        raw = self.raw_func
        if isinstance(raw, dict):
            func = OrderedDict()
            for col, fs in raw.items():
                fs = [fs] if isinstance(fs, str) else list(fs)
                func[col] = [check_supported(f) for f in fs]
            self.func = func
        elif isinstance(raw, (list, tuple)):
            self.func = [check_supported(f) for f in raw]
        else:
            self.func = [check_supported(raw)]

    # ---- compile (reference: _compile_funcs, aggregation.py:525-549)
    @classmethod
    def _compile_funcs(cls, op) -> ReductionSteps:
        if isinstance(op.func, list):
            it = [(None, f) for f in op.func]
        else:
            it = [(col, f) for col, fs in op.func.items() for f in fs]
        return compile_reduction(it, ddof=op.raw_func_kw.get("ddof", 1))

    # ---- tile helpers ------------------------------------------------------------------------------
    @classmethod
    def _gen_map_chunks(cls, op, in_chunks, out_df, steps: ReductionSteps) -> List[ChunkData]:
        """one stage=map chunk per input chunk (aggregation.py:473-523)"""
        out = []
        for chunk in in_chunks:
            map_op = op.copy().reset_key()
            map_op.groupby_params = dict(op.groupby_params, as_index=True)
            map_op.stage = OperandStage.map
            map_op.pre_funcs = steps.pre_funcs
            map_op.agg_funcs = steps.agg_funcs
            out.append(map_op.new_chunk([chunk], shape=out_df.shape, index=(chunk.index[0], 0),
                                        index_value=out_df.index_value, columns_value=out_df.columns_value,
                                        dtypes=out_df.dtypes))
        return out

    @classmethod
    def _agg_like_op(cls, op, stage, steps, with_post):
        new = op.copy().reset_key()
        new.stage = stage
        params = dict(op.groupby_params)
        params.pop("selection", None)
        params.pop("by", None)
        params["level"] = list(range(op.index_levels))   # group on index levels after the map stage
        new.groupby_params = params
        new.agg_funcs = steps.agg_funcs
        new.pre_funcs = None
        new.post_funcs = steps.post_funcs if with_post else None
        return new

    @classmethod
    def _build_tree_chunks(cls, op, chunks, steps, combine_size, input_chunk_size=None, chunk_store_limit=None):
        """combine_size-ary tree of concat + stage=combine (aggregation.py:715-780)"""
        out_df = op.outputs[0]
        check_size = chunk_store_limit is not None
        concat_size = input_chunk_size
        while (not check_size or concat_size < chunk_store_limit) and len(chunks) > combine_size:
            new_chunks = []
            for idx, i in enumerate(range(0, len(chunks), combine_size)):
                chks = chunks[i:i + combine_size]
                if len(chks) == 1:
                    chk = chks[0]
                else:
                    for j, c in enumerate(chks):
                        c.index = (j, 0)
                    chk = DataFrameConcat(gpu=op.gpu).new_chunk(chks, dtypes=chks[0].dtypes,
                                                               shape=(np.nan, out_df.shape[1]), index=(idx, 0))
                cop = cls._agg_like_op(op, OperandStage.combine, steps, with_post=False)
                new_chunks.append(cop.new_chunk([chk], index=(idx, 0), shape=(np.nan, out_df.shape[1]),
                                                index_value=chks[0].index_value,
                                                columns_value=out_df.columns_value, dtypes=out_df.dtypes))
            chunks = new_chunks
            if concat_size is not None:
                concat_size *= combine_size
        return (chunks, concat_size) if concat_size else chunks

    @classmethod
    def _build_out_tileable(cls, op, out_df, combined, steps):
        """final concat + stage=agg (aggregation.py:782-823)"""
        if len(combined) == 1:
            chk = combined[0]
        else:
            for j, c in enumerate(combined):
                c.index = (j, 0)
            chk = DataFrameConcat(gpu=op.gpu).new_chunk(combined, dtypes=combined[0].dtypes,
                                                        shape=(np.nan, out_df.shape[1]), index=(0, 0))
        aop = cls._agg_like_op(op, OperandStage.agg, steps, with_post=True)
        aop.tileable_op_key = op.key
        chunk = aop.new_chunk([chk], index=(0, 0), **out_df.params)
        return _finish_tileable(op, out_df, [chunk])

    @classmethod
    def _gen_shuffle_chunks(cls, op, chunks):
        """hash shuffle: n map -> proxy -> n reduce, partition count == number of mapper chunks
        (aggregation.py:428-471)"""
        n = len(chunks)
        maps = []
        for chunk in chunks:
            mop = DataFrameGroupByOperand(stage=OperandStage.map, shuffle_size=n, gpu=op.gpu)
            maps.append(mop.new_chunk([chunk], shape=(np.nan, np.nan), index=chunk.index,
                                      index_value=op.outputs[0].index_value))
        proxy = DataFrameShuffleProxy(gpu=op.gpu).new_chunk(maps, shape=())
        reduces = []
        for idx in itertools.product(range(n), range(1)):
            rop = DataFrameGroupByOperand(stage=OperandStage.reduce, n_reducers=n, gpu=op.gpu)
            reduces.append(rop.new_chunk([proxy], shape=(np.nan, np.nan), index=idx,
                                         index_value=op.outputs[0].index_value))
        return reduces

    @classmethod
    def _gen_shuffle_chunks_with_pivot(cls, op, chunks):
        """sort=True: PSRS range shuffle -- sample every partial, pick pivots, split by pivots, gather
        (aggregation.py:316-426, 563-641)"""
        n = len(chunks)
        by = op.groupby_params["by"]
        samples = []
        for i, chunk in enumerate(chunks):
            sop = DataFramePSRSGroupbySample(n_partition=n, by=by, gpu=op.gpu)
            samples.append(sop.new_chunk([chunk], shape=(n, np.nan), index=(i, 0), index_value=chunk.index_value))
        pivot = DataFrameGroupbyConcatPivot(n_partition=n, by=by, gpu=op.gpu).new_chunk(samples, shape=(n,))
        maps = []
        for chunk in chunks:
            mop = DataFrameGroupbySortShuffle(stage=OperandStage.map, shuffle_size=n, n_partition=n, gpu=op.gpu)
            maps.append(mop.new_chunk([chunk, pivot], shape=(np.nan, np.nan), index=chunk.index,
                                      index_value=chunk.index_value))
        proxy = DataFrameShuffleProxy(gpu=op.gpu).new_chunk(maps, shape=())
        reduces = []
        for i, m in enumerate(maps):
            rop = DataFrameGroupbySortShuffle(stage=OperandStage.reduce, reducer_index=(i, 0), n_reducers=n, by=by,
                                              gpu=op.gpu)
            reduces.append(rop.new_chunk([proxy], shape=(np.nan, np.nan), index=m.index, index_value=m.index_value))
        return reduces

    @classmethod
    def _perform_shuffle(cls, op, agg_chunks, in_df, out_df, steps):
        """aggregation.py:643-702"""
        if op.groupby_params["sort"] and len(in_df.chunks) > 1:
            reduce_chunks = cls._gen_shuffle_chunks_with_pivot(op, agg_chunks)
        else:
            reduce_chunks = cls._gen_shuffle_chunks(op, agg_chunks)
        outs = []
        for chunk in reduce_chunks:
            aop = cls._agg_like_op(op, OperandStage.agg, steps, with_post=True)
            aop.tileable_op_key = op.key
            outs.append(aop.new_chunk([chunk], index=chunk.index, **out_df.params))
        return _finish_tileable(op, out_df, outs)

    @classmethod
    def _tile_with_tree(cls, op, in_df, out_df, steps):
        chunks = cls._gen_map_chunks(op, in_df.chunks, out_df, steps)
        chunks = cls._build_tree_chunks(op, chunks, steps, op.combine_size)
        return cls._build_out_tileable(op, out_df, chunks, steps)

    @classmethod
    def _tile_with_shuffle(cls, op, in_df, out_df, steps):
        chunks = cls._gen_map_chunks(op, in_df.chunks, out_df, steps)
        return cls._perform_shuffle(op, chunks, in_df, out_df, steps)

    @classmethod
    def _build_tree_and_shuffle_chunks(cls, op, in_df, out_df, steps, sample_map_chunks, sample_agg_sizes):
        """decision rule of method='auto' (aggregation.py:838-884)"""
        combine_size = op.combine_size
        left = cls._gen_map_chunks(op, in_df.chunks[combine_size:], out_df, steps)
        input_size = sum(sample_agg_sizes) / len(sample_agg_sizes)
        limit = op.chunk_store_limit / 4
        combined, concat_size = cls._build_tree_chunks(op, sample_map_chunks + left, steps, combine_size,
                                                       input_size, limit)
        if concat_size <= limit:
            return cls._build_out_tileable(op, out_df, combined, steps)
        return cls._perform_shuffle(op, combined, in_df, out_df, steps)

    @classmethod
    def tile(cls, op):
        """Generator: yields chunk lists that must be executed before tiling can continue (method='auto'
        samples the first combine_size map chunks, aggregation.py:887-925); returns the tiled tileable."""
        in_df = op.inputs[0]
        out_df = op.outputs[0]
        steps = cls._compile_funcs(op)
        if op.method == "auto":
            if len(in_df.chunks) <= op.combine_size:
                return cls._tile_with_tree(op, in_df, out_df, steps)
            recorder = SizeRecorder()
            chunks = cls._gen_map_chunks(op, in_df.chunks[:op.combine_size], out_df, steps)
            for c in chunks:
                c.op.size_recorder_name = recorder.name
            yield chunks
            # sizes of still speculative results are first read as their capacity BOUNDS (known without asking
            # the device): a plan that is a tree under the bounds is a tree under the exact sizes too.  Only when
            # the bounds alone would choose the shuffle branch are the exact sizes read (one host wait) -- what
            # the reference always does (aggregation.py:887-925).
            raw_sizes, agg_sizes = recorder.get()
            out = cls._build_tree_and_shuffle_chunks(op, in_df, out_df, steps, chunks, agg_sizes)
            if recorder.bounded and len(out.chunks) > 1:
                raw_sizes, agg_sizes = recorder.get(exact=True)
                out = cls._build_tree_and_shuffle_chunks(op, in_df, out_df, steps, chunks, agg_sizes)
            SizeRecorder.destroy(recorder.name)
            for c in chunks:
                c.op.size_recorder_name = None
            return out
        if op.method == "shuffle":
            return cls._tile_with_shuffle(op, in_df, out_df, steps)
        if op.method == "tree":
            return cls._tile_with_tree(op, in_df, out_df, steps)
        raise NotImplementedError(op.method)


class SizeRecorder:
    """reference: aggregation.py:79-89 (a remote object there; process-local registry here)."""
    _registry: Dict[str, "SizeRecorder"] = {}

    def __init__(self):
        self.name = new_key("sizerec")
        self._raw, self._agg, self._bound = [], [], []
        self.bounded = False
        SizeRecorder._registry[self.name] = self

    def record(self, raw_size: int, agg_size, bound=None):
        """agg_size: the exact size or a zero-argument callable giving it (may wait for the device); bound: an
        upper bound known on the host (capacity of a speculative result), or None"""
        self._raw.append(raw_size)
        self._agg.append(agg_size)
        self._bound.append(bound)
        if bound is not None:
            self.bounded = True

    def get(self, exact: bool = False):
        if exact or not self.bounded:
            self._agg = [a() if callable(a) else a for a in self._agg]
            return self._raw, self._agg
        return self._raw, [b if b is not None else (a() if callable(a) else a) for a, b in zip(self._agg, self._bound)]

    @classmethod
    def lookup(cls, name):
        return cls._registry[name]

    @classmethod
    def destroy(cls, name):
        cls._registry.pop(name, None)


def _finish_tileable(op, out_df, chunks):
    out_df.chunks = chunks
    out_df.nsplits = ((np.nan,) * len(chunks), (out_df.shape[1],))
    return out_df


# ----------------------------------------------------------------------------------------------------
# user-facing objects
# ----------------------------------------------------------------------------------------------------
class DataFrame:
    """Lazy chunked DataFrame (row chunks only, like the inputs of the groupby path)."""

    def __init__(self, data, chunk_size: Optional[int] = None, gpu: bool = False, _tileable=None):
        if _tileable is not None:
            self.data = _tileable
            return
        if not isinstance(data, pd.DataFrame):
            data = pd.DataFrame(data)
        n = len(data)
        chunk_size = int(chunk_size or options.chunk_size or max(n, 1))
        op = DataFrameDataSource(gpu=False)
        t = op.new_tileable([], shape=data.shape, dtypes=data.dtypes, index_value=IndexValue(data.index),
                            columns_value=IndexValue(data.columns))
        chunks = []
        for i, start in enumerate(range(0, max(n, 1), chunk_size)):
            piece = data.iloc[start:start + chunk_size]
            cop = DataFrameDataSource(gpu=False, data=piece)
            chunks.append(cop.new_chunk([], shape=piece.shape, index=(i, 0), dtypes=data.dtypes,
                                        index_value=IndexValue(piece.index), columns_value=IndexValue(data.columns)))
        t.chunks = chunks
        t.nsplits = (tuple(c.shape[0] for c in chunks), (data.shape[1],))
        self.data = t
        if gpu:
            self.data = self.to_gpu().data

    @classmethod
    def from_chunks(cls, chunks, gpu: Optional[bool] = None) -> "DataFrame":
        """Lazy frame over row chunks that already exist -- pandas frames in host memory (then ``.to_gpu()``
        inserts the H2D step) or ``DeviceFrame`` objects already resident in HBM (what a GPU band's chunk
        store hands to the processor, processor.py:147-175).  No data is copied or sliced."""
        from .device import DeviceFrame

        first = chunks[0]
        resident = isinstance(first, DeviceFrame)
        if gpu is None:
            gpu = resident
        dtypes = first.dtypes
        try:
            ckey = tuple(first.columns)
            columns = _columns_cache.get(ckey)
        except TypeError:
            ckey, columns = None, None
        if columns is None:
            columns = pd.Index(list(first.columns))
            if ckey is not None:
                if len(_columns_cache) > 256:
                    _columns_cache.clear()
                _columns_cache[ckey] = columns
        n = sum(len(c) for c in chunks)
        op = DataFrameDataSource(gpu=gpu)
        cv = IndexValue(columns)
        t = op.new_tileable([], shape=(n, len(columns)), dtypes=dtypes,
                            index_value=IndexValue(lambda n=n: pd.RangeIndex(n)), columns_value=cv)
        out, start = [], 0
        for i, piece in enumerate(chunks):
            cop = DataFrameDataSource(gpu=gpu, data=piece)
            m = len(piece)
            out.append(cop.new_chunk([], shape=(m, len(columns)), index=(i, 0), dtypes=dtypes,
                                     index_value=IndexValue(lambda a=start, b=start + m: pd.RangeIndex(a, b)),
                                     columns_value=cv))
            start += m
        t.chunks = out
        t.nsplits = (tuple(c.shape[0] for c in out), (len(columns),))
        return cls(None, _tileable=t)

    # metadata passthrough
    @property
    def shape(self):
        return self.data.shape

    @property
    def dtypes(self):
        return self.data.dtypes

    @property
    def chunks(self):
        return self.data.chunks

    @property
    def op(self):
        return self.data.op

    @property
    def columns_value(self):
        return self.data.columns_value

    @property
    def index_value(self):
        return self.data.index_value

    # ---- column expressions, row masks, setitem (the operators TPC-H Q1 applies before its groupby) ----------
    def __getitem__(self, item):
        if isinstance(item, RowMask):
            if item.df is not self:
                raise NotImplementedError("mask of another frame")
            return self._per_chunk(lambda gpu: DataFrameFilter(mask_col=item.column, mask_op=item.op,
                                                               mask_value=item.value, gpu=gpu),
                                   self.data.dtypes, rows_known=False)
        if isinstance(item, (list, tuple)):
            raise NotImplementedError("column projection is applied by the groupby selection in this mirror")
        if item not in self.data.dtypes.index:
            raise KeyError(item)
        return ColumnExpr(self, [(item, 0.0, 1.0)])

    def __getattr__(self, name):
        data = self.__dict__.get("data")
        if data is not None and name in data.dtypes.index:
            return self[name]
        raise AttributeError(name)

    def __setitem__(self, name, value):
        if not isinstance(value, ColumnExpr) or value.df is not self:
            raise NotImplementedError("only expressions over this frame's own columns can be assigned")
        for c, _, _ in value.factors:
            if self.data.dtypes[c] != np.float64:
                raise NotImplementedError(f"column {c!r}: float64 operands only")
        dtypes = self.data.dtypes.copy()
        dtypes[name] = np.dtype("float64")
        new = self._per_chunk(lambda gpu: DataFrameSetitem(target=name, factors=list(value.factors), gpu=gpu), dtypes,
                              rows_known=True)
        self.data = new.data          # pandas semantics: the assignment changes this (lazy) frame

    def _per_chunk(self, make_op, dtypes, rows_known):
        src = self.data
        gpu = bool(src.op.gpu)
        columns = pd.Index(list(dtypes.index))
        cv = IndexValue(columns)
        nrows = src.shape[0] if rows_known else np.nan
        t = make_op(gpu).new_tileable([src], shape=(nrows, len(columns)), dtypes=dtypes, index_value=src.index_value,
                                      columns_value=cv)
        t.chunks = [make_op(gpu).new_chunk([c], shape=(c.shape[0] if rows_known else np.nan, len(columns)), index=c.index,
                                           dtypes=dtypes, index_value=c.index_value, columns_value=cv)
                    for c in src.chunks]
        t.nsplits = src.nsplits if rows_known else ((np.nan,) * len(src.chunks), (len(columns),))
        return DataFrame(None, _tileable=t)

    def to_gpu(self) -> "DataFrame":
        src = self.data
        op = DataFrameToGPU(gpu=True)
        t = op.new_tileable([src], **src.params)
        t.chunks = [DataFrameToGPU(gpu=True).new_chunk([c], shape=c.shape, index=c.index, dtypes=c.dtypes,
                                                       index_value=c.index_value, columns_value=c.columns_value)
                    for c in src.chunks]
        t.nsplits = src.nsplits
        return DataFrame(None, _tileable=t)

    def groupby(self, by=None, level=None, as_index=True, sort=True, group_keys=True) -> "DataFrameGroupBy":
        if level is not None:
            raise NotImplementedError("groupby(level=...) on the input frame is outside the GPU path")
        if by is None:
            raise TypeError("You have to supply one of 'by' and 'level'")
        if not isinstance(by, (list, tuple)):
            by = [by]
        by = list(by)
        for b in by:
            if b not in self.data.dtypes.index:
                raise KeyError(b)
        return DataFrameGroupBy(self, by, as_index=as_index, sort=sort, group_keys=group_keys)

    # execution sugar (reference: .execute() / .fetch(), deploy/oscar/session.py)
    def execute(self, session=None, **kw):
        from .session import get_default_session
        (session or get_default_session()).execute(self, **kw)
        return self

    def fetch(self, session=None):
        from .session import get_default_session
        return (session or get_default_session()).fetch(self)


_mock_cache: Dict[Any, Any] = {}
_mock_fast: Dict[Any, Any] = {}       # keyed on the identity of the (cached) dtypes Series
_columns_cache: Dict[Any, Any] = {}
_mock_meta: Dict[int, Any] = {}       # per cached mock frame: what agg() reads from it


class DataFrameGroupBy:
    def __init__(self, df: DataFrame, by: List[Any], as_index=True, sort=True, group_keys=True, selection=None):
        self.df = df
        self.by = by
        self.as_index = as_index
        self.sort = sort
        self.group_keys = group_keys
        self.selection = selection

    def __getitem__(self, item):
        """column selection (reference: groupby/getitem.py:38-41 only records `selection`)"""
        sel = list(item) if isinstance(item, (list, tuple)) else [item]
        for c in sel:
            if c not in self.df.dtypes.index:
                raise KeyError(c)
        return DataFrameGroupBy(self.df, self.by, self.as_index, self.sort, self.group_keys, selection=sel)

    @property
    def groupby_params(self):
        p = dict(by=list(self.by), level=None, as_index=self.as_index, sort=self.sort, group_keys=self.group_keys)
        if self.selection is not None:
            p["selection"] = list(self.selection)
        return p

    def _mock_result(self, func, **kw) -> pd.DataFrame:
        """schema inference by running real pandas on a small mock frame (reference: build_mock_agg_result,
        aggregation.py:142-161)"""
        dtypes = self.df.dtypes
        fast_key = None
        try:   # same schema object, same call as before: no string building at all
            fast_key = (id(dtypes), tuple(self.by), self.as_index, self.sort,
                        None if self.selection is None else tuple(self.selection),
                        id(func) if isinstance(func, (dict, list)) else func, tuple(sorted(kw.items())))
            hit = _mock_fast.get(fast_key)
            if hit is not None and hit[0] is dtypes and (hit[1] is func or hit[1] == func):
                return hit[2]
        except TypeError:
            fast_key = None
        out = self._mock_result_slow(func, dtypes, **kw)
        if fast_key is not None:
            if len(_mock_fast) > 256:
                _mock_fast.clear()
            _mock_fast[fast_key] = (dtypes, func, out)
        return out

    def _mock_result_slow(self, func, dtypes, **kw) -> pd.DataFrame:
        try:   # the mock only depends on the schema and the call: cache it (graph building stays cheap)
            cache_key = (tuple((str(c), str(dt)) for c, dt in dtypes.items()), tuple(map(str, self.by)), self.as_index,
                         self.sort, None if self.selection is None else tuple(map(str, self.selection)), repr(func),
                         repr(sorted(kw.items())))
            hit = _mock_cache.get(cache_key)
            if hit is not None:
                return hit
        except Exception:  # noqa: BLE001
            cache_key = None

        def mock_col(dt):
            try:
                return np.arange(1, 5).astype(dt)
            except (TypeError, ValueError):
                return pd.Series(list("abcd")).astype(dt)

        mock = pd.DataFrame({c: mock_col(dt) for c, dt in dtypes.items()})
        g = mock.groupby(self.by, as_index=self.as_index, sort=self.sort)
        if self.selection is not None:
            g = g[self.selection]
        out = g.agg(func, **kw)
        if cache_key is not None:
            _mock_cache[cache_key] = out
        return out

    def agg(self, func=None, method="auto", combine_size=None, **kw) -> DataFrame:
        if method is None:
            method = "auto"
        if method not in ("shuffle", "tree", "auto"):
            raise ValueError(f"Method {method} is not available, please specify 'tree' or 'shuffle")
        gpu = bool(self.df.op.gpu)
        op = DataFrameGroupByAgg(raw_func=func, raw_func_kw=kw, method=method,
                                 raw_groupby_params=self.groupby_params, groupby_params=self.groupby_params,
                                 combine_size=combine_size or options.combine_size,
                                 chunk_store_limit=options.chunk_store_limit, use_inf_as_na=False, gpu=gpu)
        op._normalize()
        mock = self._mock_result(func, **kw)
        if isinstance(mock, pd.Series):
            raise NotImplementedError("series-valued aggregation results are outside the GPU path")
        meta = _mock_meta.get(id(mock))
        if meta is None or meta[0] is not mock:
            if len(_mock_meta) > 256:
                _mock_meta.clear()
            meta = _mock_meta[id(mock)] = (mock, mock.index.nlevels, mock.shape[1], mock.dtypes, mock.index[:0],
                                           mock.columns)
        op.index_levels = meta[1]
        t = op.new_tileable([self.df.data], shape=(np.nan, meta[2]), dtypes=meta[3],
                            index_value=IndexValue(meta[4]), columns_value=IndexValue(meta[5]))
        return DataFrame(None, _tileable=t)

    aggregate = agg

    def sum(self, **kw):
        return self.agg("sum", **kw)

    def count(self, **kw):
        return self.agg("count", **kw)

    def min(self, **kw):
        return self.agg("min", **kw)

    def max(self, **kw):
        return self.agg("max", **kw)

    def mean(self, **kw):
        return self.agg("mean", **kw)

    def var(self, **kw):
        return self.agg("var", **kw)
