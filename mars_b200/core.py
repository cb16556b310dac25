"""Host-side mirror of the slice of ``mars.core`` that the groupby-aggregation path touches.

The reference is pure Python on the host and is not installed on a GPU box, so the operand / chunk /
executor-registry protocol that the hot path plugs into is restated here with the same names and
semantics.  The GPU executors in ``mars_b200.executors`` only use this protocol (duck-typed), therefore they
also run unchanged on the *real* reference operands (``mars_b200.plugin``).

Reference counterparts
  OperandStage                     mars/core/operand/base.py:57-61
  Operand.register_executor        mars/core/operand/core.py:458-464
  execute(results, op)             mars/core/operand/core.py:475-512
  MapReduceOperand / ShuffleProxy  mars/core/operand/shuffle.py:26-130
  ProcessorContext                 mars/services/subtask/worker/processor.py:55-70
"""
from __future__ import annotations

import enum
import itertools
import uuid
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple


class OperandStage(enum.Enum):
    map = 0
    reduce = 1
    combine = 2
    agg = 3


class ExecutionError(Exception):
    """Wraps an exception raised by an executor (mars/core/base.py ExecutionError)."""

    def __init__(self, nested_error: BaseException):
        super().__init__(nested_error)
        self.nested_error = nested_error


_key_counter = itertools.count()
_key_salt = uuid.uuid4().hex[:8]   # per process: keys of different processes never collide


def new_key(prefix: str = "") -> str:
    return f"{prefix}{next(_key_counter):08x}{_key_salt}"


class IndexValue:
    """Minimal stand-in for mars IndexValue: keeps the pandas Index (metadata only).  ``index`` may be a
    zero-argument callable that builds the Index on first use (chunk metas are rarely looked at)."""

    def __init__(self, index, name=None):
        self._lazy = index if callable(index) else None
        self._index_obj = None if callable(index) else index
        self.name = name if name is not None else (None if callable(index) else getattr(index, "name", None))

    @property
    def _index(self):
        if self._index_obj is None and self._lazy is not None:
            self._index_obj = self._lazy()
            self._lazy = None
        return self._index_obj

    def to_pandas(self):
        return self._index

    @property
    def names(self):
        return list(getattr(self._index, "names", [self.name]))


class ChunkData:
    __slots__ = ("op", "key", "index", "shape", "dtypes", "index_value", "columns_value", "extra")

    def __init__(self, op, index=None, shape=None, dtypes=None, index_value=None, columns_value=None, **extra):
        self.op = op
        self.key = new_key("c")
        self.index = index
        self.shape = shape
        self.dtypes = dtypes
        self.index_value = index_value
        self.columns_value = columns_value
        self.extra = extra

    @property
    def inputs(self):
        return self.op.inputs

    @property
    def ndim(self):
        return len(self.shape) if self.shape is not None else 2

    def __repr__(self):
        return f"Chunk<{type(self.op).__name__}:{getattr(self.op.stage, 'name', None)} idx={self.index} {self.key[:9]}>"


class TileableData:
    def __init__(self, op, shape=None, dtypes=None, index_value=None, columns_value=None, chunks=None,
                 nsplits=None):
        self.op = op
        self.key = new_key("t")
        self.shape = shape
        self.dtypes = dtypes
        self.index_value = index_value
        self.columns_value = columns_value
        self.chunks = chunks
        self.nsplits = nsplits

    @property
    def inputs(self):
        return self.op.inputs

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def params(self):
        return dict(shape=self.shape, dtypes=self.dtypes, index_value=self.index_value,
                    columns_value=self.columns_value)


class Operand:
    """Base operand: attribute bag + executor registry (per class, searched along the MRO)."""

    _op_type_to_executor: Dict[type, Callable] = {}

    def __init__(self, **kw):
        self.stage: Optional[OperandStage] = kw.pop("stage", None)
        self.gpu: Optional[bool] = kw.pop("gpu", None)
        self.inputs: List[Any] = []
        self.outputs: List[Any] = []
        self.key = new_key("o")
        for k, v in kw.items():
            setattr(self, k, v)

    # -- reference: TileableOperandMixin.register_executor / unregister_executor (core.py:458-464)
    @classmethod
    def register_executor(cls, executor: Callable):
        Operand._op_type_to_executor[cls] = executor

    @classmethod
    def unregister_executor(cls):
        Operand._op_type_to_executor.pop(cls, None)

    def is_gpu(self):
        return bool(self.gpu)

    def copy(self):
        new = type(self).__new__(type(self))
        new.__dict__.update(self.__dict__)
        new.inputs = list(self.inputs)
        new.outputs = []
        return new

    def reset_key(self):
        self.key = new_key("o")
        return self

    def new_chunk(self, inputs, **kw):
        self.inputs = list(inputs or [])
        chunk = ChunkData(self, **kw)
        self.outputs = [chunk]
        return chunk

    def new_tileable(self, inputs, **kw):
        self.inputs = list(inputs or [])
        t = TileableData(self, **kw)
        self.outputs = [t]
        return t

    def pre_execute(self):
        pass

    def post_execute(self):
        pass

    @classmethod
    def execute(cls, ctx, op):
        raise NotImplementedError(
            f"{cls.__name__} has no built-in executor: this framework ships GPU executors only "
            f"(call mars_b200.register_gpu_executors())")


def execute(results: Dict, op: Operand):
    """Global dispatcher -- same contract as mars/core/operand/core.py:475-512: look up the executor
    registered for type(op) along the MRO; NotImplementedError moves on to the next candidate; finally the
    class's own ``execute``."""
    op.pre_execute()
    try:
        registry = Operand._op_type_to_executor
        executor = registry.get(type(op))
        if executor is not None:          # the common case: an executor registered for the exact class
            try:
                return executor(results, op)
            except NotImplementedError:
                pass
        for cls in type(op).__mro__[1:]:
            executor = registry.get(cls)
            if executor is None:
                continue
            try:
                return executor(results, op)
            except NotImplementedError:
                continue
        return type(op).execute(results, op)
    finally:
        op.post_execute()


class MapReduceOperand(Operand):
    """reference: mars/core/operand/shuffle.py:53-122"""

    def __init__(self, **kw):
        self.reducer_index = kw.pop("reducer_index", None)
        self.n_reducers = kw.pop("n_reducers", None)
        self.mapper_id = kw.pop("mapper_id", None)
        super().__init__(**kw)

    def iter_mapper_keys(self, input_id=0):
        proxy = self.inputs[input_id]
        return [(c.key, self.outputs[0].index) for c in proxy.op.inputs]

    def iter_mapper_data(self, ctx, input_id=0, pop=False, skip_none=False):
        """Yields the values mappers stored under (mapper_chunk_key, reducer_index)."""
        for key in self.iter_mapper_keys(input_id):
            try:
                yield ctx.pop(key) if pop else ctx[key]
            except KeyError:
                if not skip_none:
                    raise
                if not pop:
                    ctx[key] = None


class ShuffleProxy(Operand):
    """Virtual operand joining all mappers of a shuffle; never executed (shuffle.py:26-50)."""

    @classmethod
    def execute(cls, ctx, op):
        return None


class ProcessorContext(dict):
    """dict-like execution context handed to executors (processor.py:55-70)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._current_chunk = None
        # multi-GPU runs with a push transport: hash-shuffle mappers stop after pass 1 of the partition kernel and
        # the exchange moves the rows (executors.DeferredSplit)
        self.defer_split = False

    def set_current_chunk(self, chunk):
        self._current_chunk = chunk

    def get_current_chunk(self):
        return self._current_chunk


def build_chunk_graph(result_chunks: Iterable[ChunkData]) -> List[ChunkData]:
    """Topologically ordered list of all chunks reachable from result_chunks (inputs first).  Order is
    deterministic (depth-first in input order) so that every rank of a multi-GPU job walks it identically."""
    order, seen = [], set()
    stack = [(c, False) for c in reversed(list(result_chunks))]
    while stack:
        c, expanded = stack.pop()
        if c.key in seen:
            continue
        if expanded:
            seen.add(c.key)
            order.append(c)
            continue
        stack.append((c, True))
        for inp in reversed(c.op.inputs):
            if inp.key not in seen:
                stack.append((inp, False))
    return order
