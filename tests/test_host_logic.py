"""Host-side logic on CPU: tiling shapes, function decomposition, executor / session plumbing.

The executors run against tests/fake_kernels.py (a CPU double of mars_b200.kernels) so that everything above
the C ABI -- step planning, partial tuples, shuffle blocks incl. empty ones, tree concat, post step wiring,
result frame labels / dtypes -- is checked against the reference's golden outputs without a GPU.  The CUDA
kernels themselves are covered by the `-m gpu` tests."""
from __future__ import annotations

import numpy as np
import pandas as pd
import pytest
import torch

import fake_kernels
import mars_groupby_oracle as orc
from conftest import assert_frames_match, golden_names, load_golden


@pytest.fixture()
def md(monkeypatch):
    import mars_b200 as m
    from mars_b200 import executors

    monkeypatch.setattr(executors, "_kernels", fake_kernels)
    monkeypatch.setattr(executors, "_device", lambda: torch.device("cpu"))
    return m


def tile(md, t):
    gen = type(t.op).tile(t.op)
    try:
        next(gen)
    except StopIteration as stop:
        return stop.value
    raise AssertionError("unexpected yield")


def op_names(chunks):
    from mars_b200.core import build_chunk_graph

    return [f"{type(c.op).__name__}:{getattr(c.op.stage, 'name', None)}" for c in build_chunk_graph(chunks)]


# ---------------------------------------------------------------- tile shapes (reference test_groupby.py:97-192)
def test_tree_tiling_shape(md):
    raw = pd.DataFrame({"key": np.arange(100) % 7, "v": np.arange(100.0)})
    r = md.DataFrame(raw, chunk_size=10).groupby("key").agg(["sum", "mean", "count"], method="tree")
    t = tile(md, r.data)
    names = op_names(t.chunks)
    assert len(t.chunks) == 1 and t.chunks[0].op.stage.name == "agg"
    assert names.count("DataFrameGroupByAgg:map") == 10
    assert names.count("DataFrameGroupByAgg:combine") == 3      # 10 -> 3 -> 1 with combine_size 4
    assert names.count("DataFrameConcat:None") == 4
    agg_op = t.chunks[0].op
    assert agg_op.groupby_params["level"] == [0] and "by" not in agg_op.groupby_params


def test_hash_shuffle_tiling_shape(md):
    raw = pd.DataFrame({"key": np.arange(100) % 7, "v": np.arange(100.0)})
    r = md.DataFrame(raw, chunk_size=25).groupby("key", sort=False).agg(["sum"], method="shuffle")
    t = tile(md, r.data)
    names = op_names(t.chunks)
    assert len(t.chunks) == 4
    assert names.count("DataFrameGroupByOperand:map") == 4 and names.count("DataFrameGroupByOperand:reduce") == 4
    assert names.count("DataFrameShuffleProxy:None") == 1
    for c in t.chunks:
        red = c.op.inputs[0].op
        assert red.n_reducers == 4
        assert all(m.op.shuffle_size == 4 for m in red.inputs[0].op.inputs)     # R == number of mapper chunks


def test_sort_shuffle_tiling_shape(md):
    """sort=True + shuffle = PSRS range shuffle: sample per map chunk -> one pivot chunk -> sort-shuffle map
    (reads the pivots) -> proxy -> sort-shuffle reduce -> agg (reference test_groupby.py:97-122)"""
    raw = pd.DataFrame({"key": np.arange(100) % 7, "v": np.arange(100.0)})
    r = md.DataFrame(raw, chunk_size=25).groupby("key", sort=True).agg(["sum"], method="shuffle")
    t = tile(md, r.data)
    names = op_names(t.chunks)
    assert len(t.chunks) == 4
    assert names.count("DataFramePSRSGroupbySample:None") == 4
    assert names.count("DataFrameGroupbyConcatPivot:None") == 1
    assert names.count("DataFrameGroupbySortShuffle:map") == 4 and names.count("DataFrameGroupbySortShuffle:reduce") == 4
    assert "DataFrameGroupByOperand:map" not in names
    for c in t.chunks:
        red = c.op.inputs[0].op
        assert red.n_reducers == 4 and red.by == ["key"]
        for m in red.inputs[0].op.inputs:
            assert m.op.n_partition == 4 and type(m.op.inputs[1].op).__name__ == "DataFrameGroupbyConcatPivot"


def test_decomposition_matches_reference_compiler():
    from mars_b200.reduction import compile_reduction

    steps = compile_reduction([(None, f) for f in ["sum", "mean", "count", "min", "max", "var"]])
    got = [(a.map_func_name, a.agg_func_name, p.func.kind) for a in steps.agg_funcs
           for p in steps.pre_funcs if p.output_key == a.input_key]
    # SURVEY.md Appendix A.3 (probed from ReductionCompiler): K = 5 atomic steps, count merges with SUM
    assert sorted(got) == sorted([("sum", "sum", "identity"), ("count", "sum", "identity"), ("min", "min", "identity"),
                                  ("max", "max", "identity"), ("sum", "sum", "square")])
    var = [p for p in steps.post_funcs if p.func_name == "var"][0]
    assert [k for k in var.input_keys] == ["agg_sum_square", "agg_sum_identity", "agg_count_identity"]
    steps = compile_reduction([(None, f) for f in ["sum", "mean", "count"]])
    assert len(steps.agg_funcs) == 2 and len(steps.pre_funcs) == 1


# ------------------------------------------------------------------------------ full path through the double
@pytest.mark.parametrize("name", golden_names())
def test_executor_plumbing_reproduces_reference(md, name):
    g = load_golden(name)
    s = g.spec
    sess = md.new_session(default=False)
    r = md.DataFrame(g.raw, chunk_size=s["chunk_size"]).to_gpu().groupby(s["by"], sort=s["sort"]) \
        .agg(s["func"], method=s["method"])
    r.execute(session=sess)
    got = r.fetch(session=sess)
    assert_frames_match(got, g.result, sort_index=not s["sort"])
    chunks = sorted(r.data.chunks, key=lambda c: c.index)
    assert len(chunks) == len(g.chunk_results)
    for c, exp in zip(chunks, g.chunk_results):
        assert_frames_match(sess._results[c.key], exp, sort_index=not s["sort"])


def test_auto_method_samples_then_decides(md):
    rng = np.random.default_rng(0)
    raw = pd.DataFrame({"key": rng.integers(0, 50, 4000), "v": rng.random(4000)})
    sess = md.new_session(default=False)
    r = md.DataFrame(raw, chunk_size=400).to_gpu().groupby("key").agg(["sum", "count"], method="auto")
    r.execute(session=sess)
    assert len(r.data.chunks) == 1                                   # small partials -> tree
    assert_frames_match(r.fetch(session=sess), raw.groupby("key").agg(["sum", "count"]), sort_index=False)
    # force the shuffle decision with a tiny store limit (reference test_groupby.py:172-192)
    from mars_b200 import options
    old = options.chunk_store_limit
    options.chunk_store_limit = 64
    try:
        sess = md.new_session(default=False)
        r = md.DataFrame(raw, chunk_size=400).to_gpu().groupby("key", sort=False).agg(["sum", "count"], method="auto")
        r.execute(session=sess)
        assert len(r.data.chunks) > 1
        assert "DataFrameGroupByOperand:reduce" in sess.executed_ops
        assert_frames_match(r.fetch(session=sess), raw.groupby("key").agg(["sum", "count"]), sort_index=True)
    finally:
        options.chunk_store_limit = old


def test_shuffle_mapper_emits_every_reducer_block(md):
    from mars_b200.core import ProcessorContext, build_chunk_graph, execute

    raw = pd.DataFrame({"key": np.array([3, 3, 3, 3], dtype=np.int64), "v": np.arange(4.0)})
    r = md.DataFrame(raw, chunk_size=1).to_gpu().groupby("key", sort=False).agg(["sum"], method="shuffle")
    t = tile(md, r.data)
    ctx = ProcessorContext()
    for c in build_chunk_graph(t.chunks):
        if type(c.op).__name__ == "DataFrameShuffleProxy":
            break
        ctx.set_current_chunk(c)
        execute(ctx, c.op)
    mappers = [c for c in build_chunk_graph(t.chunks) if type(c.op).__name__ == "DataFrameGroupByOperand"
               and c.op.stage.name == "map"]
    target = int(orc.partition_ids([np.array([3])], 4)[0])
    for m in mappers:
        for i in range(4):
            idx, payload = ctx[(m.key, (i, 0))]           # all 4 blocks exist (shuffle.py:124-130)
            assert idx == m.index and payload[-1] is False
            assert len(payload[0]) == (1 if i == target else 0)


def test_errors_wrap_into_execution_error(md):
    from mars_b200 import ExecutionError

    raw = pd.DataFrame({"key": ["a", "b"], "v": [1.0, 2.0]})
    r = md.DataFrame(raw, chunk_size=1).to_gpu().groupby("key").agg(["sum"], method="tree")
    with pytest.raises(ExecutionError):
        r.execute(session=md.new_session(default=False))


# ------------------------------------------------------------------- frames that are not sliced until asked
def test_row_range_frames_give_pointers_without_slicing():
    """shuffle blocks / result partials are row ranges of shared columns (DeviceFrame.from_range): pointers by
    arithmetic, tensors only on demand, and both views agree"""
    from mars_b200.device import ColumnSet, DeviceFrame

    cols = [torch.arange(100, dtype=torch.int64), torch.arange(100, dtype=torch.float64) * 2,
            torch.arange(100, dtype=torch.float64) * 3]
    src = ColumnSet(cols)
    f = DeviceFrame.from_range(["k"], ["a", "b"], src, 0, 1, 1, 3, 10, 25)
    g = DeviceFrame.from_range(["k"], ["a", "b"], src, 0, 1, 1, 3, 10, 25)
    h = DeviceFrame.from_range(["k"], ["a", "b"], src, 0, 1, 1, 3, 25, 40)
    assert len(f) == 15 and f.nbytes == 15 * 8 * 3 and f.col_dtype(0) == torch.float64
    assert f.same_index(g) and not f.same_index(h)
    kp, dp = f.ptrs()
    assert f._src is not None                        # nothing was sliced so far
    assert kp == [cols[0].data_ptr() + 80] and dp == [cols[1].data_ptr() + 80, cols[2].data_ptr() + 80]
    assert f.index[0].tolist() == list(range(10, 25)) and f._src is None      # now it is
    assert f.data[1].tolist() == [3.0 * i for i in range(10, 25)]
    assert (kp, dp) == f.ptrs()                      # same pointers from the materialised tensors
    assert f.same_index(g)                           # mixed: one side sliced, the other not


def test_block_backed_column_set_matches_the_block_views():
    """ColumnSet.from_block: column pointers of a result block pair by arithmetic == pointers of the views"""
    from mars_b200 import kernels as K
    from mars_b200.device import ColumnSet, DeviceFrame, Pending

    accs = ((0, K.SUM), (1, K.SUM), (0, K.COUNT), (2, K.MAX), (2, K.COUNT))
    fc = K._FastCall(2, [K.F64, K.F64, K.I64], accs)
    bi, bf = fc.alloc(64, "cpu")
    src = ColumnSet.from_block(bi, bf, fc)
    keys, outs = fc.views(bi, bf)
    assert src.ptrs == [t.data_ptr() for t in keys + outs]
    assert src.dtypes == [t.dtype for t in keys + outs]
    assert src._cols is None                         # no view was made for the pointers
    # a pending partial frame over accumulators [1, 3): capacity rows until the count settles
    bi[fc.n_int, 0] = 7
    pend = Pending(bi, fc.n_int, 64, lambda: None, fblock=bf, layout=fc)
    f = DeviceFrame.from_range(["x", "y"], ["p", "q"], src, 0, 2, 2 + 1, 2 + 3, 0, 64, pending=pend, slot=(1, 3))
    assert f.is_pending and f.ptrs_raw()[1] == [outs[1].data_ptr(), outs[2].data_ptr()]
    assert len(f) == 7 and not f.is_pending          # resolves: one 8-byte read
    assert [t.numel() for t in f.data] == [7, 7] and f.data[0].data_ptr() == outs[1].data_ptr()


def test_result_frame_assembly_small_and_large_paths_agree():
    """tiny results take the cheap constructors, large ones the zero-copy ones: same frame either way"""
    from mars_b200 import executors as E

    rng = np.random.default_rng(3)
    for n in (5, 5000):
        k1, k2 = rng.integers(0, 50, n), rng.integers(0, 7, n)
        cols = pd.MultiIndex.from_tuples([("v", "sum"), ("v", "count"), ("w", "mean")])
        host = {("v", "count"): rng.integers(1, 9, n), ("w", "mean"): rng.random(n), ("v", "sum"): rng.random(n)}
        meta = dict(names=["a", "b"], col_value=cols, two_level=True, as_index=True, dtypes=None)
        got = E._assemble_result(dict(host), [k1, k2], meta)
        exp = pd.DataFrame({c: host[c] for c in cols}, index=pd.MultiIndex.from_arrays([k1, k2], names=["a", "b"]))
        pd.testing.assert_frame_equal(got, exp)
        assert list(got.columns) == list(cols) and got.dtypes.tolist() == exp.dtypes.tolist()


def test_session_pauses_and_restores_the_cyclic_gc(md):
    import gc

    raw = pd.DataFrame({"key": np.arange(40) % 5, "v": np.arange(40.0)})
    r = md.DataFrame(raw, chunk_size=10).to_gpu().groupby("key").agg(["sum"], method="tree")
    assert gc.isenabled()
    got = r.execute(session=md.new_session(default=False))
    assert gc.isenabled()
    assert got is r


# ---------------------------------------------------------------- band-level fusion of the tree branch
def _run_tree(md, raw, by, func, chunk, method="tree", fuse=True, sort=True):
    sess = md.new_session(default=False)
    sess.fuse = fuse
    r = md.DataFrame(raw, chunk_size=chunk).to_gpu().groupby(by, sort=sort).agg(func, method=method)
    r.execute(session=sess)
    return r.fetch(session=sess), sess


@pytest.mark.parametrize("func", [["sum", "mean", "count"], ["min", "max", "var"],
                                  {"v0": ["sum", "mean"], "v1": ["count"]}])
def test_fused_tree_is_one_batch_launch_and_one_agg_launch(md, func):
    rng = np.random.default_rng(1)
    n = 4000
    raw = pd.DataFrame({"k1": rng.integers(0, 3, n), "k2": rng.integers(0, 2, n), "v0": rng.normal(10, 3, n),
                        "v1": rng.integers(0, 9, n)})
    from mars_b200 import executors

    executors._dense_rejected.clear()
    fake_kernels.calls.clear()
    got, sess = _run_tree(md, raw, ["k1", "k2"], func, 400)
    assert sess.fused_trees == 1
    assert fake_kernels.calls.count("preagg_dense_batch_async") == 1          # 10 map operands + 3 combines: one launch
    assert fake_kernels.calls.count("merge_small_post_async") == 1            # merge + post functions + sort: one launch
    assert "preagg_dense_async" not in fake_kernels.calls and "post_mean" not in fake_kernels.calls
    # the plan the reference's tiler produced is still what was (logically) executed
    assert sess.executed_ops.count("DataFrameGroupByAgg:map") == 10
    assert sess.executed_ops.count("DataFrameGroupByAgg:combine") == 3
    exp = orc.run_groupby_agg(raw, ["k1", "k2"], func, 400, method="tree", sort=True)
    assert_frames_match(got, exp, sort_index=False)
    # per-operand execution (what a real Mars worker does) gives the same frame
    got2, sess2 = _run_tree(md, raw, ["k1", "k2"], func, 400, fuse=False)
    assert sess2.fused_trees == 0
    assert_frames_match(got2, exp, sort_index=False)


def test_fused_tree_under_method_auto_merges_the_sampled_partials(md):
    """method='auto' runs the first combine_size map operands on their own (size sampling,
    aggregation.py:887-925); the rest of the tree is fused and the sampled partial tuples join as ready pieces"""
    rng = np.random.default_rng(2)
    n = 6000
    raw = pd.DataFrame({"akey": rng.integers(0, 5, n), "av0": rng.random(n), "av1": rng.random(n)})
    fake_kernels.calls.clear()
    got, sess = _run_tree(md, raw, "akey", ["sum", "mean", "count"], 500, method="auto")
    assert sess.fused_trees == 1
    # the band is launched once, BEFORE tiling (session._launch_eager): the sampled map operands are not run on
    # their own, their sizes are recorded as the capacity bound of the speculative result (no host wait)
    assert fake_kernels.calls.count("preagg_dense_batch_async") == 1 and "preagg_dense_async" not in fake_kernels.calls
    assert sess.executed_ops.count("DataFrameGroupByAgg:map") == 12
    exp = orc.run_groupby_agg(raw, "akey", ["sum", "mean", "count"], 500, method="tree", sort=True)
    assert_frames_match(got, exp, sort_index=False)
    # without the band-level batch: 4 sampled + 8 further map operands, each on its own
    fake_kernels.calls.clear()
    got2, sess2 = _run_tree(md, raw, "akey", ["sum", "mean", "count"], 500, method="auto", fuse=False)
    assert fake_kernels.calls.count("preagg_dense_async") == 12 and "preagg_dense_batch_async" not in fake_kernels.calls
    assert_frames_match(got2, exp, sort_index=False)


def test_eager_band_gives_way_when_auto_chooses_the_shuffle_branch(md, monkeypatch):
    """the band is launched before tiling; when the sampled sizes send method='auto' to the shuffle branch (here:
    a tiny chunk_store_limit), the skipped map chunks are aggregated for real (EagerRef.materialize) and the
    graph runs operand by operand"""
    rng = np.random.default_rng(6)
    n = 6000
    raw = pd.DataFrame({"bkey": rng.integers(0, 5, n), "bv0": rng.random(n)})
    monkeypatch.setattr(md.options, "chunk_store_limit", 64)
    fake_kernels.calls.clear()
    sess = md.new_session(default=False)
    r = md.DataFrame(raw, chunk_size=500).to_gpu().groupby("bkey", sort=False).agg(["sum", "count"], method="auto")
    r.execute(session=sess)
    got = r.fetch(session=sess)
    assert len(r.data.chunks) > 1 and "DataFrameGroupByOperand:map" in sess.executed_ops
    assert sess.fused_trees == 0
    assert_frames_match(got, raw.groupby("bkey").agg(["sum", "count"]), sort_index=True)


def test_fused_tree_replays_when_the_band_is_not_low_cardinality(md):
    from mars_b200 import executors

    rng = np.random.default_rng(3)
    n = 3000
    raw = pd.DataFrame({"fkey": rng.integers(0, 200, n), "fv": rng.random(n)})     # 200 groups > what a CTA holds
    before = set(executors._dense_rejected)
    fake_kernels.calls.clear()
    got, sess = _run_tree(md, raw, "fkey", ["sum", "count", "max"], 300)
    exp = orc.run_groupby_agg(raw, "fkey", ["sum", "count", "max"], 300, method="tree", sort=True)
    assert_frames_match(got, exp, sort_index=False)
    assert fake_kernels.calls.count("preagg_dense_batch_async") == 1 and "groupby_aggregate" in fake_kernels.calls
    assert set(executors._dense_rejected) - before          # remembered: the next run goes per operand
    fake_kernels.calls.clear()
    got2, sess2 = _run_tree(md, raw, "fkey", ["sum", "count", "max"], 300)
    assert sess2.fused_trees == 0 and "preagg_dense_batch_async" not in fake_kernels.calls
    assert_frames_match(got2, exp, sort_index=False)


def test_shuffle_branch_is_not_fused(md):
    rng = np.random.default_rng(4)
    raw = pd.DataFrame({"key": rng.integers(0, 5, 2000), "v": rng.random(2000)})
    got, sess = _run_tree(md, raw, "key", ["sum"], 250, method="shuffle", sort=False)
    assert sess.fused_trees == 0
    assert_frames_match(got, raw.groupby("key").agg(["sum"]), sort_index=True)


def test_mid_cardinality_band_takes_the_mid_batch_kernel(md):
    """sum / mean / count over a few hundred groups: too many for the dense kernel, one launch of the
    mid-cardinality batch kernel once the shape is known (the first run learns it)"""
    from mars_b200 import executors

    rng = np.random.default_rng(8)
    n = 6000
    raw = pd.DataFrame({"mkey": rng.integers(0, 300, n), "mv0": rng.random(n), "mv1": rng.random(n)})
    exp = orc.run_groupby_agg(raw, "mkey", ["sum", "mean", "count"], 500, method="tree", sort=True)
    got, sess = _run_tree(md, raw, "mkey", ["sum", "mean", "count"], 500)          # dense is tried, fails, replays
    assert_frames_match(got, exp, sort_index=False)
    assert any(k[0] == ("mkey",) for k in executors._dense_rejected)
    fake_kernels.calls.clear()
    got, sess = _run_tree(md, raw, "mkey", ["sum", "mean", "count"], 500)
    assert sess.fused_trees == 1
    assert fake_kernels.calls.count("preagg_mid_batch_async") == 1 and "preagg_dense_batch_async" not in fake_kernels.calls
    assert fake_kernels.calls.count("merge_small_post_async") == 1 and "groupby_aggregate" not in fake_kernels.calls
    assert_frames_match(got, exp, sort_index=False)
    # per operand: every map operand is one asynchronous mid launch
    fake_kernels.calls.clear()
    got, sess = _run_tree(md, raw, "mkey", ["sum", "mean", "count"], 500, fuse=False)
    assert fake_kernels.calls.count("preagg_mid_batch_async") == 12 and "groupby_aggregate" not in fake_kernels.calls
    assert_frames_match(got, exp, sort_index=False)


@pytest.mark.parametrize("method", ["tree", "shuffle"])
@pytest.mark.parametrize("by", [["akey"], ["akey", "ak2"]])
def test_as_index_false_returns_the_keys_as_columns(md, method, by):
    """groupby(..., as_index=False).agg(...): same frame as pandas -- the keys are columns, a default RangeIndex
    (reference: reset_index at the end of _execute_agg, aggregation.py:1259-1263)"""
    rng = np.random.default_rng(12)
    raw = pd.DataFrame({"akey": rng.integers(0, 50, 5000), "ak2": rng.integers(0, 3, 5000), "av": rng.random(5000),
                        "aw": rng.integers(0, 9, 5000)})
    func = {"av": ["sum", "max"], "aw": ["count"]}
    sess = md.new_session(default=False)
    got = md.DataFrame(raw, chunk_size=1200).to_gpu().groupby(by, as_index=False).agg(func, method=method) \
        .execute(session=sess).fetch(session=sess)
    exp = raw.groupby(by, as_index=False).agg(func)
    got = got.sort_values([(b, "") for b in by]).reset_index(drop=True)
    assert list(got.columns) == list(exp.columns)
    pd.testing.assert_frame_equal(got, exp.reset_index(drop=True), rtol=1e-9, atol=0, check_exact=False)


@pytest.mark.parametrize("method,sort", [("shuffle", False), ("tree", True)])
def test_nearly_distinct_chunks_skip_the_map_side_fold(md, method, sort):
    """a chunk that was grouped and came back with ~as many groups as rows makes the following chunks of that shape
    hand their rows on as partials (row_partials) -- until a chunk's sample stops looking distinct (the first chunks of
    a run are all sampled, later ones every 8th); the final frame is
    the oracle's either way (NaN values and duplicate keys inside a chunk included)"""
    from mars_b200 import executors

    rng = np.random.default_rng(23)
    n, chunk = 350_000, 70_000
    raw = pd.DataFrame({"dkey": rng.integers(0, 40_000_000, n), "dv": rng.normal(3, 2, n)})
    raw.loc[rng.integers(0, n, 2000), "dv"] = np.nan
    raw.loc[[5, 70_007, 70_008, 279_999, 349_999], "dkey"] = 77
    raw.loc[140_000:209_999, "dkey"] = rng.integers(0, 50, 70_000)    # chunk 2 is NOT distinct
    raw["dkey"] = raw["dkey"].astype(np.int64)
    funcs = ["sum", "mean", "count", "min", "max", "var"]
    fake_kernels.calls.clear()
    for learned in (executors._pt_streak, executors._group_ratio, executors._distinct_shapes):
        learned.clear()                       # what earlier chunks of this shape taught the map stage (per process)
    before = executors.passthrough_chunks
    sess = md.new_session(default=False)
    got = md.DataFrame(raw, chunk_size=chunk).to_gpu().groupby("dkey", sort=sort).agg(funcs, method=method) \
        .execute(session=sess).fetch(session=sess)
    exp = orc.run_groupby_agg(raw, ["dkey"], funcs, chunk, method=method, sort=sort)
    assert_frames_match(got, exp, sort_index=not sort)
    # chunk 0 is measured, 1 passes through, 2 fails its sample and is grouped, 3 is grouped (and measured) again, 4 passes
    assert executors.passthrough_chunks - before == 2
    assert fake_kernels.calls.count("row_partials") == 2
