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
import gc
import os
from typing import Dict, List, Optional

import pandas as pd

from . import core
from . import executors as _ex
from .core import ChunkData, ProcessorContext, build_chunk_graph
from .dataframe import DataFrame, DataFrameDataSource, DataFrameGroupByAgg, DataFrameShuffleProxy


_op_names: Dict[object, str] = {}
_map_name = "DataFrameGroupByAgg:map"
_gc_frozen = False


class Session:
    def __init__(self, exchange=None):
        """exchange: a ``mars_b200.parallel.Exchange`` for multi-GPU runs (None: single GPU)."""
        _ex.register_gpu_executors()
        self.exchange = exchange
        self.rank = exchange.rank if exchange is not None else 0
        self.world = exchange.world if exchange is not None else 1
        self._results: Dict[str, object] = {}
        self.executed_ops: List[str] = []      # "<OperandClass>:<stage>" log, used by tests
        # band-level fusion of the tree branch (MGB_FUSE=0: per-operand execution, what a real Mars worker does)
        self.fuse = os.environ.get("MGB_FUSE", "1") not in ("0", "")
        self.fused_trees = 0
        self._eager = None                     # executors.EagerBatch of the graph that is being executed

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
            from .parallel import reducer_rank

            return reducer_rank(int(chunk.index[0]), len(op.inputs[0].op.inputs), self.world)
        return owners[op.inputs[0].key]

    # ---------------------------------------------------------------------------------- execution
    def _run_chunks_local(self, result_chunks: List[ChunkData], ctx: ProcessorContext, owners: Dict[str, int]):
        """single-process form of ``_run_chunks``: same order, same reference-counted release of inputs, none
        of the placement / transfer bookkeeping.  The chunk operands of a subtask run back to back on one band
        (SubtaskProcessor._execute_graph, processor.py:216-282); where the graph holds the tree branch of a
        groupby-agg (stage=map operands under a concat / combine tree that only a stage=agg operand observes)
        the band-level batch executor runs that part as ONE launch (executors.execute_fused_tree) -- unless
        ``self.fuse`` is off, which reproduces the per-operand execution of a real Mars worker."""
        order = [c for c in build_chunk_graph(result_chunks) if c.key not in owners]
        refs: Dict[str, int] = {}
        for c in order:
            for inp in c.op.inputs:
                k = inp.key
                refs[k] = refs.get(k, 0) + 1
        keep = {c.key for c in result_chunks}
        log = self.executed_ops
        execute = core.execute
        fused_at, skip = {}, set()
        eager = self._eager
        if self.fuse and len(order) > 3:
            for ft in _ex.find_fused_trees(order, keep):
                fused_at[ft.root.key] = ft
                skip.update(c.key for c in ft.nodes)

        def release(op):
            for inp in op.inputs:
                k = inp.key
                n = refs[k] - 1
                refs[k] = n
                if n == 0 and k not in keep:
                    ctx.pop(k, None)

        def run_one(c):
            op = c.op
            if eager is not None:
                if (type(op) is DataFrameGroupByAgg and op.stage is core.OperandStage.map and len(op.inputs) == 1
                        and op.inputs[0].key in eager.covered and not eager.used):
                    # the band was aggregated as a whole by the launch that preceded tiling: this chunk has no
                    # partial tuple of its own; its sampled size is the capacity bound of the speculative result
                    ctx[c.key] = _ex.EagerRef(eager, op, op.inputs[0].key)
                    rec = getattr(op, "size_recorder_name", None)
                    if rec is not None:
                        ref = ctx[c.key]
                        frame = eager.covered[op.inputs[0].key]
                        _ex._record_sizes(ctx, rec, frame.nbytes,
                                          lambda ref=ref, key=c.key: _ex.exact_partial_size(ctx, key, ref), eager.bound)
                    log.append(_map_name)
                    release(op)
                    return
                for inp in op.inputs:       # a consumer outside the fused tree needs the chunk's real partials
                    v = ctx.get(inp.key)
                    if isinstance(v, _ex.EagerRef):
                        ctx[inp.key] = v.materialize()
            if not isinstance(op, DataFrameShuffleProxy):
                ctx._current_chunk = c
                try:
                    execute(ctx, op)
                except Exception as e:  # same wrapping as processor.py:210-214
                    raise core.ExecutionError(e) from e
                tag = (type(op), op.stage)
                name = _op_names.get(tag)
                if name is None:
                    name = _op_names[tag] = f"{type(op).__name__}:{getattr(op.stage, 'name', None)}"
                log.append(name)
            release(op)
            if op.stage is core.OperandStage.reduce and isinstance(op, core.MapReduceOperand):
                for key in op.iter_mapper_keys():    # mapper blocks of this reducer are consumed
                    ctx.pop(key, None)

        for c in order:
            owners[c.key] = 0
            ft = fused_at.get(c.key) if fused_at else None
            if ft is not None:
                try:
                    done = _ex.execute_fused_tree(ctx, ft, eager)
                except Exception as e:
                    raise core.ExecutionError(e) from e
                if done:
                    self.fused_trees += 1
                    for node in ft.nodes:
                        tag = (type(node.op), node.op.stage)
                        name = _op_names.get(tag)
                        if name is None:
                            name = _op_names[tag] = f"{type(node.op).__name__}:{getattr(node.op.stage, 'name', None)}"
                        log.append(name)
                        release(node.op)
                else:       # not a shape the batch kernel takes: the operands one by one, in graph order
                    for node in ft.nodes:
                        run_one(node)
                continue
            if c.key in skip:
                continue
            run_one(c)

    def _run_chunks(self, result_chunks: List[ChunkData], ctx: ProcessorContext, owners: Dict[str, int]):
        if self.world == 1:
            return self._run_chunks_local(result_chunks, ctx, owners)
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
        # The chunk graph of a large shuffle keeps ~1e5 small containers alive (frames, column lists, tensor
        # views); none of them forms a reference cycle (device.Pending holds weak references), but every few
        # hundred allocations the cyclic collector walks all of them.  It is paused while the graph runs; and the
        # objects that exist when the first graph runs (the imported libraries) are moved to the permanent
        # generation once, so that the collection which follows a run only looks at what the run left behind.
        global _gc_frozen
        if not _gc_frozen and os.environ.get("MGB_GC_FREEZE", "1") not in ("0", ""):
            gc.collect()
            gc.freeze()
            _gc_frozen = True
            # A graph leaves a few thousand short-lived, acyclic objects behind; with the default young-generation
            # threshold (700 allocations) a collection runs right after every execute().  MGB_GC_THRESHOLD (default
            # 50000, 0 = leave the interpreter's setting) raises it once.
            thr = int(os.environ.get("MGB_GC_THRESHOLD", "50000"))
            if thr > 0:
                cur = gc.get_threshold()
                if cur[0] and cur[0] < thr:
                    gc.set_threshold(thr, cur[1], cur[2])
        was_enabled = gc.isenabled()
        gc.disable()
        try:
            return self._execute(df, **kw)
        finally:
            self._eager = None
            if was_enabled:
                gc.enable()

    def _launch_eager(self, t, ctx, owners):
        """Before the graph of a groupby-agg over resident row chunks is tiled, enqueue the band-level batch over
        all its input chunks: the map stage of every chunk is needed under any plan (tree or shuffle,
        aggregation.py:473-523), so the GPU can stream the band while the host tiles and walks the graph.  The
        tree branch is then served by this launch (executors.execute_fused_tree); any other consumer gets the
        per-chunk partials on demand (executors.EagerRef)."""
        op = t.op
        if type(op) is not DataFrameGroupByAgg or op.method not in ("auto", "tree") or not op.inputs:
            return None
        src = op.inputs[0]
        chunks = getattr(src, "chunks", None)
        if not chunks or len(chunks) < 2 or len(chunks) > 4096:
            return None
        by = op.groupby_params.get("by")
        if not isinstance(by, (list, tuple)):
            return None
        steps = type(op)._compile_funcs(op)
        map_op = op.copy().reset_key()
        map_op.groupby_params = dict(op.groupby_params, as_index=True)
        map_op.stage = core.OperandStage.map
        map_op.pre_funcs, map_op.agg_funcs = steps.pre_funcs, steps.agg_funcs
        map_op.new_chunk([chunks[0]], index=(0, 0))
        if all(type(c.op) is DataFrameDataSource for c in chunks):
            for c in chunks:                      # resident chunks: the data-source operand only hands over its frame
                ctx[c.key] = c.op.data
                owners[c.key] = 0
            self.executed_ops.extend(["DataFrameDataSource:None"] * len(chunks))
        else:
            self._run_chunks_local(list(chunks), ctx, owners)   # data sources and H2D operands of the band
        frames = [ctx[c.key] for c in chunks]
        return _ex.launch_eager_band(map_op, [c.key for c in chunks], frames)

    def _execute(self, df: DataFrame, **kw):
        t = df.data
        ctx = ProcessorContext()
        ctx.defer_split = self.world > 1 and bool(getattr(getattr(self.exchange, "transport", None), "push", False))
        owners: Dict[str, int] = {}
        self._eager = None
        if t.chunks is None:
            if self.fuse and self.world == 1:
                try:
                    self._eager = self._launch_eager(t, ctx, owners)
                except NotImplementedError:
                    self._eager = None
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

    # ------------------------------------------------------------------------- sharded tree branch
    def execute_sharded_tree(self, df: DataFrame, group=None, max_groups: int = 2368, comm=None):
        """Weak-scaling execution of a low-cardinality (tree-branch) groupby-agg over row chunks that are
        sharded across ranks -- one process per GPU:

          1. every rank tiles and runs its LOCAL map -> concat/combine tree (independent, no communication);
          2. the local partial tuple is re-aggregated once (stage=combine semantics) and packed into a
             fixed-capacity block;
          3. ONE collective: all_gather of the packed blocks + row counts (NCCL over NVLink; the partials are
             a few rows, so this is latency-, not bandwidth-bound);
          4. rank 0 runs the stage=agg operand over the gathered pieces (virtual concat) and owns the result.

        ``df`` is this rank's lazy groupby-agg over its own chunks.  Returns the result frame on rank 0 and
        None elsewhere.  Falls back to NotImplementedError when a rank holds more than ``max_groups`` groups
        (then the shuffle branch with the all-to-all exchange is the right plan)."""
        import torch
        import torch.distributed as dist

        from .device import DeviceFrame, Pending, PiecewiseFrame
        from .executors import _merge_partials, _pieces_of, _plan_steps

        t = df.data
        ctx = ProcessorContext()
        owners: Dict[str, int] = {}
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        saved_rank, saved_world = self.rank, self.world
        self.rank, self.world = 0, 1                      # the local graph runs entirely on this rank
        self._eager = None
        try:
            if t.chunks is None:
                if self.fuse:
                    try:
                        self._eager = self._launch_eager(t, ctx, owners)
                    except NotImplementedError:
                        self._eager = None
                gen = type(t.op).tile(t.op)
                try:
                    to_run = next(gen)
                    while True:
                        self._run_chunks(list(to_run), ctx, owners)
                        to_run = gen.send(None)
                except StopIteration as stop:
                    if stop.value is not None and stop.value is not t:
                        t = stop.value
            if len(t.chunks) != 1 or getattr(t.chunks[0].op.stage, "name", None) != "agg":
                raise NotImplementedError("execute_sharded_tree expects the tree branch (one stage=agg chunk)")
            agg_chunk = t.chunks[0]
            feed = agg_chunk.op.inputs[0]
            self._run_chunks([feed], ctx, owners)
        finally:
            self.rank, self.world = saved_rank, saved_world
            self._eager = None
        steps = _plan_steps(agg_chunk.op, None)
        slots = _pieces_of(ctx[feed.key])
        # Every rank ships its local result as the block pair of a speculative result (int64 block with the row
        # count in its last row + float64 block, fixed capacity): ONE all-gather of identical shape on every rank,
        # no host synchronisation while the result is still speculative; rank 0's stage=agg merge reads every
        # rank's row count on the device.  The fused local tree leaves exactly such a block (one piece per slot,
        # one shared Pending): it is shipped as it is.  A rank whose partials do not fit the small-merge layout
        # ships a poisoned count (-1) instead of raising before the collective -- all ranks always enter it.
        pend = slots[0][0]._pending if len(slots[0]) == 1 else None
        if not (pend is not None and pend.value is None and pend.fblock is not None
                and all(len(sl) == 1 and sl[0]._pending is pend for sl in slots)):
            merged = _merge_partials(slots, steps, sort=False)
            pend = merged[0]._pending
            if (pend is not None and pend.value is None and pend.fblock is not None
                    and all(f._pending is pend for f in merged)):
                pass
            else:
                pend = None
        else:
            merged = [sl[0] for sl in slots]
        if pend is not None:
            bi, bf, fc = pend.block, pend.fblock, pend.layout
        else:
            bi, bf, fc = self._pack_block(merged, steps, max_groups)
        root = bi._base
        if root is None or bf._base is not root or root.shape[0] != bi.shape[0] + bf.shape[0]:
            root = torch.cat([bi, bf.view(torch.int64)], dim=0)
        g = torch.empty((world * root.shape[0], root.shape[1]), dtype=root.dtype, device=root.device)
        if comm is not None:     # the library's own NCCL all-gather (mgb_comm_allgather_i64)
            comm.allgather_i64(root, g)
        else:
            dist.all_gather_into_tensor(g, root, group=group)     # concatenation along dim 0
        g = g.view(world, root.shape[0], root.shape[1])
        gi = g[:, :bi.shape[0]]
        gf = g[:, bi.shape[0]:].view(torch.float64)
        if rank != 0:
            return None

        def no_replay():
            raise NotImplementedError("a rank's local partials do not fit the small-merge path: "
                                      "use the shuffle branch")

        pieces: List[List[DeviceFrame]] = [[] for _ in merged]
        for r in range(world):
            pr = Pending(gi[r], fc.n_int, gi.shape[2], no_replay)   # gi[r]: rows of rank r's block, contiguous
            k, a = fc.views(gi[r], gf[r])
            a0 = 0
            for i, f in enumerate(merged):
                a1 = a0 + len(f.columns)
                pieces[i].append(DeviceFrame(f.index_names, k, f.columns, a[a0:a1], pending=pr, slot=(a0, a1)))
                a0 = a1
        ctx[feed.key] = tuple(PiecewiseFrame(p) for p in pieces)
        ctx.set_current_chunk(agg_chunk)
        try:
            core.execute(ctx, agg_chunk.op)
        except Exception as e:
            raise core.ExecutionError(e) from e
        self.executed_ops.append("DataFrameGroupByAgg:agg")
        return ctx[agg_chunk.key]

    @staticmethod
    def _pack_block(merged, steps, max_groups):
        """partials whose size is known on the host -> block pair in the layout of the speculative
        small-merge result (kernels._FastCall), so that every rank enters the same collectives"""
        from . import kernels as K
        from .executors import _OPCODE

        n = len(merged[0])
        cap = K.SMALL_MAX_ROWS
        poisoned = n > min(cap, max_groups)      # does not fit: ship count -1, never raise before the collective
        cols = [t for f in merged for t in f.data]
        ops = [_OPCODE[st.agg_func] for st, f in zip(steps, merged) for _ in f.columns]
        fc = K._fast(len(merged[0].index), [K.dtype_code(t) for t in cols], tuple(enumerate(ops)))
        bi, bf = fc.alloc(cap, merged[0].device)
        keys, accs = fc.views(bi, bf)
        if not poisoned:
            for dst, src in zip(keys + accs, list(merged[0].index) + cols):
                dst[:n].copy_(src)
        bi[fc.n_int, :1].fill_(-1 if poisoned else n)
        return bi, bf, fc

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
                obj = self._results.get(c.key)
                if obj is not None and not isinstance(obj, pd.DataFrame):
                    obj = obj.to_pandas()         # device-resident result chunk: read back on its owner
                parts.append(self.exchange.fetch_result(obj, self._owners[c.key]))
        DeviceResult = _ex.DeviceResult
        if parts and all(isinstance(p, DeviceResult) for p in parts):
            return DeviceResult.fetch_all(parts)      # one pinned block, no host-side concat
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
