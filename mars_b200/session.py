"""Minimal session: tile -> chunk graph -> run every chunk operand through ``execute(ctx, op)``.

Stands in for the reference's supervisor/worker machinery (mars/services, mars/oscar -- unchanged host code
in a real deployment) so that the GPU executors can be driven, tested and benchmarked without it.  It
reproduces the parts of that machinery the hot path depends on:

  * iterative tiling (``Tiler`` handling of ``yield`` in tile, mars/core/graph/builder/chunk.py:201-250)
  * per-chunk dispatch ``execute(ctx, chunk.op)`` in topological order with reference-counted release of
    inputs (SubtaskProcessor._execute_graph, mars/services/subtask/worker/processor.py:216-282)
  * band placement for multi-GPU runs: one process per GPU (deploy/oscar/pool.py:75-80); mapper chunk c runs
    on rank c % P, reducer r on rank r % P (the reference lets the assigner place reducers round-robin,
    services/task/analyzer/analyzer.py:382-413); shuffle blocks move between ranks through the grouped
    ncclSend/ncclRecv exchange of mars_b200.parallel instead of the storage service.
"""
from __future__ import annotations

import collections
from typing import Dict, List, Optional

import pandas as pd

from . import core
from .core import ChunkData, ProcessorContext, build_chunk_graph
from .dataframe import DataFrame, DataFrameShuffleProxy


class Session:
    def __init__(self, exchange=None):
        """exchange: a ``mars_b200.parallel.Exchange`` for multi-GPU runs (None: single GPU)."""
        from .executors import register_gpu_executors

        register_gpu_executors()
        self.exchange = exchange
        self.rank = exchange.rank if exchange is not None else 0
        self.world = exchange.world if exchange is not None else 1
        self._results: Dict[str, object] = {}
        self.executed_ops: List[str] = []      # "<OperandClass>:<stage>" log, used by tests

    # ---------------------------------------------------------------------------------- placement
    def _owner(self, chunk: ChunkData, owners: Dict[str, int]) -> int:
        if self.world == 1:
            return 0
        op = chunk.op
        if not op.inputs:
            return int(chunk.index[0]) % self.world
        if isinstance(op, DataFrameShuffleProxy):
            return -1  # collective
        if isinstance(op.inputs[0].op, DataFrameShuffleProxy):
            return int(chunk.index[0]) % self.world
        return owners[op.inputs[0].key]

    # ---------------------------------------------------------------------------------- execution
    def _run_chunks(self, result_chunks: List[ChunkData], ctx: ProcessorContext, owners: Dict[str, int]):
        order = [c for c in build_chunk_graph(result_chunks) if c.key not in owners]
        refs = collections.Counter()
        for c in order:
            for inp in c.op.inputs:
                refs[inp.key] += 1
        keep = {c.key for c in result_chunks}
        for c in order:
            owner = self._owner(c, owners)
            owners[c.key] = owner
            op = c.op
            if isinstance(op, DataFrameShuffleProxy):
                if self.world > 1:
                    self.exchange.shuffle(ctx, c, owners)
            else:
                # fetch inputs that live on another rank (tree branch: partials gathered for concat)
                if self.world > 1:
                    for inp in op.inputs:
                        src = owners.get(inp.key, owner)
                        if src not in (-1, owner):
                            self.exchange.transfer(ctx, inp.key, src, owner)
                if owner == self.rank:
                    ctx.set_current_chunk(c)
                    try:
                        core.execute(ctx, op)
                    except Exception as e:  # same wrapping as processor.py:210-214
                        raise core.ExecutionError(e) from e
                    self.executed_ops.append(f"{type(op).__name__}:{getattr(op.stage, 'name', None)}")
            for inp in op.inputs:
                refs[inp.key] -= 1
                if refs[inp.key] == 0 and inp.key not in keep:
                    ctx.pop(inp.key, None)
                    if isinstance(inp.op, core.MapReduceOperand) or isinstance(op, DataFrameShuffleProxy):
                        pass
            if isinstance(op, core.MapReduceOperand) and getattr(op.stage, "name", None) == "reduce":
                # mapper blocks of this reducer are consumed
                for key in op.iter_mapper_keys():
                    ctx.pop(key, None)

    def execute(self, df: DataFrame, **kw):
        t = df.data
        ctx = ProcessorContext()
        owners: Dict[str, int] = {}
        if t.chunks is None:
            gen = type(t.op).tile(t.op)
            try:
                to_run = next(gen)
                while True:
                    self._run_chunks(list(to_run), ctx, owners)
                    if self.world > 1:
                        self.exchange.sync_size_recorders(to_run)
                    to_run = gen.send(None)
            except StopIteration as stop:
                if stop.value is not None and stop.value is not t:
                    t = stop.value
        self._run_chunks(list(t.chunks), ctx, owners)
        for c in t.chunks:
            if owners.get(c.key, 0) == self.rank:
                self._results[c.key] = ctx[c.key]
        self._owners = owners
        return df

    def fetch(self, df: DataFrame) -> pd.DataFrame:
        """Concatenate the result chunks in chunk order (multi-GPU: gathered to every rank as pandas)."""
        t = df.data
        if t.chunks is None:
            raise ValueError("tileable was not executed")
        parts = []
        for c in sorted(t.chunks, key=lambda c: c.index):
            if self.world == 1:
                parts.append(self._results[c.key])
            else:
                parts.append(self.exchange.fetch_result(self._results.get(c.key), self._owners[c.key]))
        parts = [p if isinstance(p, pd.DataFrame) else p.to_pandas() for p in parts]
        return parts[0] if len(parts) == 1 else pd.concat(parts, axis=0)


_default: Optional[Session] = None


def new_session(exchange=None, default: bool = True) -> Session:
    global _default
    s = Session(exchange=exchange)
    if default:
        _default = s
    return s


def get_default_session() -> Session:
    global _default
    if _default is None:
        _default = Session()
    return _default
