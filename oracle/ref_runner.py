"""TEST INFRASTRUCTURE ONLY -- drives the *real* reference operands (build container only).

``run_reference(raw, by, func, ...)`` builds ``md.DataFrame(raw, chunk_size).groupby(...).agg(...)`` with
the unmodified reference imported from ``$MARS_REF_DIR`` (default ``baseline/_ref``, produced by
``oracle/build_ref.sh``), tiles it with the reference's own ``ChunkGraphBuilder`` and executes every chunk
operand through the reference's global ``execute(ctx, op)`` (mars/core/operand/core.py:475-512) with a
``dict`` context -- the same call the worker makes (mars/services/subtask/worker/processor.py:206-214).
Every intermediate value is recorded so golden fixtures can pin partition assignment and partials.
"""
from __future__ import annotations

import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# built by oracle/build_ref.sh; git-ignored, but it travels to the GPU box with the gpurun snapshot
REF_DIR = os.environ.get("MARS_REF_DIR", os.path.join(_REPO, "baseline", "_ref"))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "mars"))


def _import_reference():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import ref_shim  # noqa: F401

    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import mars  # noqa: F401
    import mars.dataframe as md

    return md


class _Ctx(dict):
    _current = None

    def set_current_chunk(self, chunk):
        self._current = chunk

    def get_current_chunk(self):
        return self._current

    @staticmethod
    def new_custom_log_dir():
        return None


def run_reference(raw, by, func, chunk_size, method="tree", sort=True, as_index=True,
                  combine_size=None, selection=None, to_gpu=False):
    """Returns dict(result=pd.DataFrame, chunks=[...per-op records...])."""
    md = _import_reference()
    from mars.core import TileableGraph, TileableGraphBuilder, ChunkGraphBuilder
    from mars.core.mode import enter_mode
    from mars.core.operand import execute, OperandStage

    mdf = md.DataFrame(raw, chunk_size=chunk_size)
    if to_gpu:      # GPU bands: op.gpu propagates to every downstream operand (core/operand/core.py:62-76)
        mdf = mdf.to_gpu()
    g = mdf.groupby(by, sort=sort, as_index=as_index)
    if selection is not None:
        g = g[selection]
    kw = {}
    if combine_size is not None:
        kw["combine_size"] = combine_size
    t = g.agg(func, method=method, **kw)

    tg = TileableGraph([t.data])
    next(TileableGraphBuilder(tg).build())
    ctx = _Ctx()
    records = []
    result_chunks = {}
    # the chunk-graph builder is a generator: iterative tiling (method='auto', to_gpu) yields the graph in stages,
    # each stage is executed before the tiler continues (core/graph/builder/chunk.py:201-250)
    with enter_mode(build=False, kernel=True):
        for cg in ChunkGraphBuilder(tg, fuse_enabled=False).build():
            result_chunks = {c.key: c for c in cg.result_chunks}
            for c in cg.topological_iter():
                if c.key in ctx:
                    continue
                ctx.set_current_chunk(c)
                before = set(ctx.keys())
                execute(ctx, c.op)
                new_keys = [k for k in ctx.keys() if k not in before]
                step_keys = None
                if getattr(c.op, "agg_funcs", None) and getattr(c.op, "pre_funcs", None):
                    # slot order of the partial tuple is the (hash-seed dependent) order of op.agg_funcs: name
                    # every slot "<map func>_<identity|square>" so fixtures do not depend on it
                    ident = {p.output_key for p in c.op.pre_funcs if p.input_key == p.output_key}
                    step_keys = [f"{a.map_func_name}_{'identity' if a.input_key in ident else 'square'}"
                                 for a in c.op.agg_funcs]
                records.append(
                    dict(
                        step_keys=step_keys,
                        op=type(c.op).__name__,
                        stage=None if c.op.stage is None else OperandStage(c.op.stage).name,
                        index=tuple(c.index) if c.index is not None else None,
                        key=c.key,
                        new_keys=new_keys,
                    )
                )
    outs = sorted(result_chunks.values(), key=lambda c: c.index)
    import pandas as pd

    result = pd.concat([ctx[c.key] for c in outs], axis=0) if len(outs) > 1 else ctx[outs[0].key]
    return dict(result=result, records=records, ctx=ctx, chunk_results=[ctx[c.key] for c in outs],
                tileable=t)


if __name__ == "__main__":
    import numpy as np
    import pandas as pd

    rng = np.random.default_rng(0)
    raw = pd.DataFrame({"key": rng.integers(0, 50, 1000), "v0": rng.random(1000), "v1": rng.random(1000)})
    for method, sort in [("tree", True), ("shuffle", False), ("shuffle", True)]:
        out = run_reference(raw, "key", ["sum", "mean", "count", "min", "max", "var"], 250, method, sort)
        exp = raw.groupby("key", sort=True).agg(["sum", "mean", "count", "min", "max", "var"])
        pd.testing.assert_frame_equal(out["result"].sort_index(), exp, rtol=1e-9)
        print(method, sort, "OK", [(r["op"], r["stage"]) for r in out["records"]][:12])


# ----------------------------------------------------------------------------------------------------
# timed CPU arm (bench.py --impl reference / cpu_baseline): the reference's own operands on all host cores
# ----------------------------------------------------------------------------------------------------
_POOL_STATE = None     # (chunk graph order, map chunks) -- inherited by the forked workers


def _pool_run_map(i):
    """worker: the data-source + stage=map operands of map chunk i through the reference's execute(ctx, op)"""
    from mars.core.mode import enter_mode
    from mars.core.operand import execute

    order, maps = _POOL_STATE
    ctx = _Ctx()
    m = maps[i]
    with enter_mode(build=False, kernel=True):
        for src in m.op.inputs:
            ctx.set_current_chunk(src)
            execute(ctx, src.op)
        ctx.set_current_chunk(m)
        execute(ctx, m.op)
    return ctx[m.key]


class ReferenceTreeRun:
    """One tree-branch groupby-agg over row chunks with the UNMODIFIED reference: its tiling, its compiled
    reduction steps, its pandas-backed ``execute(ctx, op)`` classmethods.  The stage=map operands (one per
    row chunk, what dominates the cost) are fanned out over ``cores`` forked worker processes -- what Mars
    does with ``n_cpu`` bands; concat / combine / agg operands run in the parent, like the last subtasks of
    the reference's chunk graph.  ``step()`` runs the whole graph once and returns the result frame."""

    def __init__(self, raw, by, func, chunk_size, cores, method="tree", sort=True):
        global _POOL_STATE
        import multiprocessing as mp

        md = _import_reference()
        from mars.core import TileableGraph, TileableGraphBuilder, ChunkGraphBuilder
        from mars.core.operand import OperandStage

        t = md.DataFrame(raw, chunk_size=chunk_size).groupby(by, sort=sort).agg(func, method=method)
        tg = TileableGraph([t.data])
        next(TileableGraphBuilder(tg).build())
        cg = next(ChunkGraphBuilder(tg, fuse_enabled=False).build())
        self.order = list(cg.topological_iter())
        self.maps = [c for c in self.order
                     if type(c.op).__name__ == "DataFrameGroupByAgg" and c.op.stage == OperandStage.map]
        self.skip = {c.key for c in self.maps} | {s.key for m in self.maps for s in m.op.inputs}
        self.result_chunks = sorted(cg.result_chunks, key=lambda c: c.index)
        _POOL_STATE = (self.order, self.maps)
        self.cores = cores
        self.pool = mp.get_context("fork").Pool(cores)

    def step(self):
        from mars.core.mode import enter_mode
        from mars.core.operand import execute
        import pandas as pd

        ctx = _Ctx()
        partials = self.pool.map(_pool_run_map, range(len(self.maps)), chunksize=1)
        for m, p in zip(self.maps, partials):
            ctx[m.key] = p
        with enter_mode(build=False, kernel=True):
            for c in self.order:
                if c.key in self.skip:
                    continue
                ctx.set_current_chunk(c)
                execute(ctx, c.op)
        outs = [ctx[c.key] for c in self.result_chunks]
        return outs[0] if len(outs) == 1 else pd.concat(outs, axis=0)

    def close(self):
        self.pool.close()
        self.pool.join()
