"""GPU parity tests of the whole hot path through the operand API: md.DataFrame(...).groupby(...).agg(...)
tiled into DataFrameGroupByAgg map/combine/agg + DataFrameGroupByOperand map/reduce chunks, every chunk
executed through the registered `fn(ctx, op)` executors (which call libmarsgb.so through the C ABI), compared
with (i) the committed outputs of the unmodified reference (tests/golden) and (ii) the oracle on fresh seeded
inputs.  These read like the reference's own tests (mars/dataframe/groupby/tests/test_groupby_execution.py:
316-476) with pandas replaced by the golden frames."""
from __future__ import annotations

import numpy as np
import pandas as pd
import pytest

import mars_groupby_oracle as orc
from conftest import assert_frames_match, golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def md(cuda_device):
    import mars_b200 as m

    m.new_session()
    return m


def run(md, raw, by, func, chunk_size, method, sort, combine_size=None, selection=None):
    mdf = md.DataFrame(raw, chunk_size=chunk_size).to_gpu()
    g = mdf.groupby(by, sort=sort)
    if selection is not None:
        g = g[selection]
    r = g.agg(func, method=method, combine_size=combine_size)
    sess = md.new_session()
    r.execute(session=sess)
    return r.fetch(session=sess), sess, r


@pytest.mark.parametrize("name", golden_names())
def test_matches_reference_golden(md, name):
    g = load_golden(name)
    s = g.spec
    got, sess, r = run(md, g.raw, s["by"], s["func"], s["chunk_size"], s["method"], s["sort"])
    assert_frames_match(got, g.result, sort_index=not s["sort"])
    # every result chunk separately: with the hash shuffle chunk i holds exactly the keys of reducer i, i.e.
    # this pins the partition assignment of every group key to the reference's
    chunks = sorted(r.data.chunks, key=lambda c: c.index)
    assert len(chunks) == len(g.chunk_results)
    for c, exp in zip(chunks, g.chunk_results):
        assert_frames_match(sess._results[c.key], exp, sort_index=not s["sort"])


@pytest.mark.parametrize("name", ["single_key_all6_shuffle", "two_keys_minmaxvar_shuffle", "q1_dict_tree"])
def test_map_stage_partials_match_reference(md, name):
    """stage=map output (tuple of K partial frames) against the reference's partial tuples."""
    from mars_b200.core import ProcessorContext, build_chunk_graph, execute

    g = load_golden(name)
    s = g.spec
    mdf = md.DataFrame(g.raw, chunk_size=s["chunk_size"]).to_gpu()
    r = mdf.groupby(s["by"], sort=s["sort"]).agg(s["func"], method=s["method"])
    t = r.data
    gen = type(t.op).tile(t.op)
    try:
        next(gen)
    except StopIteration as stop:
        t = stop.value
    ctx = ProcessorContext()
    map_i = 0
    for c in build_chunk_graph(t.chunks):
        if type(c.op).__name__ == "DataFrameGroupByAgg" and c.op.stage.name != "map":
            continue
        if type(c.op).__name__ not in ("DataFrameDataSource", "DataFrameToGPU", "DataFrameGroupByAgg"):
            continue
        ctx.set_current_chunk(c)
        execute(ctx, c.op)
        if type(c.op).__name__ == "DataFrameGroupByAgg":
            tup = ctx[c.key]
            exp = g.maps[map_i]
            assert len(tup) == len(exp)
            for step, frame in zip(c.op.agg_funcs, tup):
                pre = [p for p in c.op.pre_funcs if p.output_key == step.input_key][0]
                key = f"{step.map_func_name}_{pre.func.kind}"
                e = exp[key]
                got = frame.to_pandas()
                e = e[list(got.columns)]     # the reference also carries dead columns (squared keys)
                assert_frames_match(got.sort_index(), e.sort_index(), sort_index=False)
            map_i += 1
    assert map_i == len(g.maps)


CASES = [
    # n, key spec, value cols, func, chunk, method, sort
    (50_000, ("key", 1000), 4, ["sum", "mean", "count"], 5_000, "tree", True),
    (50_000, ("key", 1000), 4, ["sum", "mean", "count"], 5_000, "shuffle", False),
    (50_000, ("key", 1000), 4, ["sum", "mean", "count"], 5_000, "auto", True),
    (60_000, ("key", 40_000), 1, ["sum", "count"], 7_500, "shuffle", False),
    (60_000, ("key", 40_000), 1, ["sum", "count"], 7_500, "auto", False),
    (40_000, ("two", 900), 8, ["min", "max", "var"], 5_000, "shuffle", False),
    (40_000, ("zipf", 0), 1, ["mean", "count"], 5_000, "shuffle", False),
    (60_000, ("key", 40_000), 1, ["sum", "count"], 7_500, "shuffle", True),      # PSRS range shuffle
    (40_000, ("two", 900), 3, ["min", "max", "var"], 5_000, "shuffle", True),
    (20_000, ("key", 3), 2, ["sum", "mean"], 2_500, "shuffle", True),            # fewer keys than partitions
    (10_000, ("key", 50), 2, "sum", 10_000, "tree", True),       # single chunk, single function
    (999, ("key", 7), 3, ["var"], 100, "tree", True),             # ragged last chunk
]


def make_inputs(n, keyspec, nv, seed):
    rng = np.random.default_rng(seed)
    kind, card = keyspec
    if kind == "key":
        cols = {"key": rng.integers(0, card, n)}
        by = "key"
    elif kind == "two":
        c1 = int(np.sqrt(card))
        cols = {"k1": rng.integers(0, c1, n), "k2": rng.integers(0, card // c1, n)}
        by = ["k1", "k2"]
    else:
        cols = {"key": np.minimum(rng.zipf(1.2, n), 10 ** 8).astype(np.int64)}
        by = "key"
    for i in range(nv):
        cols[f"v{i}"] = rng.normal(10, 3, n)
    return pd.DataFrame(cols), by


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[1][0]}{c[1][1]}-{c[5]}-{c[6]}")
def test_matches_oracle_on_seeded_inputs(md, case):
    n, keyspec, nv, func, chunk, method, sort = case
    raw, by = make_inputs(n, keyspec, nv, seed=n + nv)
    got, sess, r = run(md, raw, by, func, chunk, method, sort)
    omethod = method
    if method == "auto":
        omethod = "shuffle" if len(r.data.chunks) > 1 else "tree"
    exp = orc.run_groupby_agg(raw, by, func, chunk, method=omethod, sort=sort)
    assert_frames_match(got, exp, sort_index=not sort)
    ops = set(sess.executed_ops)
    if omethod == "shuffle" and not sort:
        assert "DataFrameGroupByOperand:map" in ops and "DataFrameGroupByOperand:reduce" in ops
    if omethod == "shuffle" and sort:
        assert "DataFrameGroupbySortShuffle:map" in ops and "DataFrameGroupbySortShuffle:reduce" in ops
        # chunk i of the result holds exactly the keys of range partition i
        rec = {}
        orc.run_groupby_agg(raw, by, func, chunk, method="shuffle", sort=True, record=rec)
        chunks = sorted(r.data.chunks, key=lambda c: c.index)
        for c, e in zip(chunks, rec["chunk_results"]):
            assert_frames_match(sess._results[c.key], e, sort_index=False)
    assert "DataFrameGroupByAgg:map" in ops and "DataFrameGroupByAgg:agg" in ops


def test_column_selection_and_dict_funcs(md):
    raw, by = make_inputs(30_000, ("two", 6), 5, seed=11)
    func = {"v0": ["sum", "mean"], "v1": ["sum"], "v3": ["mean", "count"]}
    got, _, _ = run(md, raw, by, func, 4_000, "tree", True, selection=["v0", "v1", "v3"])
    exp = orc.run_groupby_agg(raw, by, func, 4_000, method="tree", sort=True, selection=["v0", "v1", "v3"])
    assert_frames_match(got, exp, sort_index=False)
    exp_pd = raw.groupby(by)[["v0", "v1", "v3"]].agg(func)
    assert_frames_match(got, exp_pd, sort_index=False)


def test_empty_reducers_and_tiny_input(md):
    # more reducers than distinct keys: most shuffle blocks are empty but must still be emitted
    raw = pd.DataFrame({"key": np.array([5, 5, 9, 5, 9, 9, 5, 5], dtype=np.int64), "v": np.arange(8.0)})
    got, sess, r = run(md, raw, "key", ["sum", "count"], 1, "shuffle", False)
    exp = raw.groupby("key").agg(["sum", "count"])
    assert_frames_match(got, exp, sort_index=True)
    assert len(r.data.chunks) == 8
    sizes = sorted(len(sess._results[c.key]) for c in r.data.chunks)
    assert sizes == [0] * 6 + [1, 1]


@pytest.mark.parametrize("nkeys,sort", [(1, False), (2, False), (1, True)])
def test_large_result_chunks_stay_on_the_device_until_fetch(md, nkeys, sort):
    """result chunks of more than 2048 groups are DeviceResult objects (the reference's GPU mode returns cudf frames);
    fetch() reads all of them into one pinned block -- same frame as pandas"""
    from mars_b200.executors import DeviceResult

    rng = np.random.default_rng(77 + nkeys)
    n = 900_000
    raw = pd.DataFrame({"k1": rng.integers(0, 500_000, n), "k2": rng.integers(0, 2, n), "v0": rng.normal(10, 3, n),
                        "v1": rng.random(n)})
    by = ["k1", "k2"][:nkeys] if nkeys > 1 else "k1"
    if nkeys == 1:
        raw = raw.drop(columns="k2")
    # (no var here: the expectation below is pandas' one-shot variance, and with 2-3 rows per group the reference's
    # sum-of-squares formula (var.py:38-48) cancels differently -- var is compared with the ORACLE, which uses that
    # formula, in the config-shaped tests and test_nearly_distinct_chunks_hand_their_rows_on_as_partials)
    func = ["sum", "mean", "count", "min", "max"]
    got, sess, r = run(md, raw, by, func, 300_000, "shuffle", sort)
    kinds = {type(sess._results[c.key]) for c in r.data.chunks}
    assert kinds == {DeviceResult}, kinds
    exp = raw.groupby(by).agg(func)
    assert_frames_match(got, exp, sort_index=not sort)
    # a single chunk reads back on its own, too
    one = sess._results[r.data.chunks[0].key]
    assert_frames_match(one.to_pandas().sort_index(), exp.loc[one.to_pandas().sort_index().index], sort_index=False)


@pytest.mark.parametrize("cfg,rows,chunk", [("cfg3", 4_000_000, 500_000), ("cfg5", 4_000_000, 500_000)])
def test_config_runner_record(md, cfg, rows, chunk, capsys):
    """the record bench.py attaches per BASELINE.json config: value, roofline fraction, oracle parity on sampled
    groups + invariants (full-frame parity per config: tests/test_baseline_configs.py)"""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import config_runner

    rec = config_runner.run_record(cfg, rows, chunk, repeats=1, n_sample=512)
    capsys.readouterr()
    assert rec["parity"]["ok"], rec["parity"]
    assert "DataFrameGroupByOperand:map" in rec["plan"] and "DataFrameGroupByOperand:reduce" in rec["plan"]
    assert 0 < rec["roofline"]["frac"] < 1 and rec["value"] > 0


def test_unsupported_function_raises(md):
    raw, by = make_inputs(100, ("key", 5), 1, seed=3)
    with pytest.raises(NotImplementedError):
        md.DataFrame(raw, chunk_size=50).to_gpu().groupby(by).agg(["median"])


@pytest.mark.parametrize("n,card,nkeys", [(20_000_000, 1000, 1), (16_000_000, 6, 2), (8_000_000, 3_000_000, 1)])
def test_full_size_properties(md, n, card, nkeys):
    """Size-independent properties at bench-like sizes (the oracle would take too long):
    sum of counts == n, every key's reducer == hash(key) % R, keys unique, sum of sums == column total,
    min/max bounded by the column extrema."""
    import torch
    from mars_b200 import kernels as K

    gen = torch.Generator(device="cuda").manual_seed(n % 1000 + card)
    if nkeys == 1:
        keys = [torch.randint(0, card, (n,), generator=gen, device="cuda", dtype=torch.int64)]
    else:
        keys = [torch.randint(0, 3, (n,), generator=gen, device="cuda", dtype=torch.int64),
                torch.randint(0, 2, (n,), generator=gen, device="cuda", dtype=torch.int64)]
    v = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)
    k, a, stats = K.groupby_aggregate(keys, [v], [(0, K.SUM), (0, K.COUNT), (0, K.MIN), (0, K.MAX)], sort=True)
    assert int(a[1].sum()) == n
    total = float(v.sum())
    assert abs(float(a[0].sum()) - total) <= 1e-9 * abs(total)
    assert float(a[2].min()) == float(v.min()) and float(a[3].max()) == float(v.max())
    assert bool((a[2] <= a[3]).all())
    if nkeys == 1:
        assert bool((k[0][1:] > k[0][:-1]).all())            # sorted & unique
        cnt = torch.bincount(keys[0], minlength=card)
        present = cnt.nonzero().flatten()
        assert torch.equal(present, k[0]) and torch.equal(cnt[present], a[1])
    else:
        code = k[0] * 2 + k[1]
        assert bool((code[1:] > code[:-1]).all())
    # partition assignment property on the partial keys
    R = 64
    pid = K.partition_ids(k, R)
    outs, offsets = K.partition_scatter(pid, R, list(k))
    h = orc.hash_keys([t.cpu().numpy() for t in outs])
    for r in range(R):
        assert bool((h[offsets[r]:offsets[r + 1]] % np.uint64(R) == r).all())


def test_speculative_dense_path_replays_when_a_chunk_is_not_dense(md):
    """The map / combine / agg chain is enqueued without waiting for intermediate sizes.  Here the first
    chunks are dense (3 groups) and the later ones are not (thousands of groups): their device-side count
    comes back -1, the -1 propagates through the pending merges, and the agg stage replays the affected
    operands on the general path from the inputs it kept alive."""
    from mars_b200 import executors

    rng = np.random.default_rng(21)
    n = 80_000
    key = np.concatenate([rng.integers(0, 3, n // 2), rng.integers(0, 5000, n // 2)])
    raw = pd.DataFrame({"spec_key": key, "spec_v0": rng.normal(10, 3, n), "spec_v1": rng.random(n)})
    func = ["sum", "mean", "count", "min", "max", "var"]
    before = set(executors._dense_rejected)
    got, sess, r = run(md, raw, "spec_key", func, 10_000, "tree", True)
    exp = orc.run_groupby_agg(raw, "spec_key", func, 10_000, method="tree", sort=True)
    assert_frames_match(got, exp, sort_index=False)
    assert any(k[0] == ("spec_key",) for k in set(executors._dense_rejected) - before), "no replay happened"
    # and again, now that the shape is known not to be dense (general path from the start)
    got2, _, _ = run(md, raw, "spec_key", func, 10_000, "tree", True)
    assert_frames_match(got2, exp, sort_index=False)


def test_pending_partials_feed_the_shuffle_branch(md):
    # dense map output (size still on the device) -> shuffle map must resolve it before partitioning
    rng = np.random.default_rng(22)
    n = 40_000
    raw = pd.DataFrame({"pk": rng.integers(0, 7, n), "pv": rng.random(n)})
    got, sess, r = run(md, raw, "pk", ["sum", "count"], 5_000, "shuffle", False)
    exp = raw.groupby("pk").agg(["sum", "count"])
    assert_frames_match(got, exp, sort_index=True)


@pytest.mark.parametrize("method", ["tree", "auto"])
@pytest.mark.parametrize("func", [["sum", "mean", "count"], ["min", "max", "var"],
                                  {"v0": ["sum", "mean"], "v1": ["count"], "v2": ["var", "max"]}])
def test_tree_branch_runs_as_one_batch_launch(md, method, func):
    """band-level batch: the map operands + concat / combine tree of the tree branch are ONE launch of the dense
    kernel over all chunks (ragged last chunk), the agg stage ONE launch; same frame as the oracle and as the
    per-operand execution of a real Mars worker (fuse off)"""
    from mars_b200 import executors, kernels

    rng = np.random.default_rng(31)
    n = 1_000_003
    raw = pd.DataFrame({"fk1": rng.integers(0, 3, n), "fk2": rng.integers(0, 2, n), "v0": rng.normal(10, 3, n),
                        "v1": rng.integers(0, 9, n), "v2": rng.random(n)})
    executors._dense_rejected.clear()
    mdf = md.DataFrame(raw, chunk_size=90_000).to_gpu()
    exp = orc.run_groupby_agg(raw, ["fk1", "fk2"], func, 90_000, method="tree", sort=True)
    warm = mdf.groupby(["fk1", "fk2"]).agg(func, method=method)      # first sight of the shape: cardinality probe
    warm.execute(session=md.new_session(default=False))
    results = {}
    for fuse in (True, False):
        sess = md.new_session(default=False)
        sess.fuse = fuse
        r = mdf.groupby(["fk1", "fk2"]).agg(func, method=method)
        before = kernels.launch_counter
        r.execute(session=sess)
        got = r.fetch(session=sess)
        results[fuse] = (got, kernels.launch_counter - before, sess)
        assert_frames_match(got, exp, sort_index=False)
    got, launches, sess = results[True]
    assert sess.fused_trees == 1
    # one band-level batch (launched before tiling; a sum / mean / count shape sends the register-accumulator kernel
    # first and the dense kernel behind it) + one agg-stage launch
    assert launches in (2, 3), launches
    assert results[False][2].fused_trees == 0 and results[False][1] > launches
    assert sess.executed_ops.count("DataFrameGroupByAgg:map") == 12


def test_infinite_values_and_nan_partials_follow_the_reference(cuda_device):
    """+inf and -inf in one group of one chunk make that chunk's partial sum NaN; the reference's combine / agg
    stages re-aggregate the partials with pandas' skipna sum, i.e. they DROP it (aggregation.py:1130-1164).  The
    per-operand path must give the oracle's (= the reference's) answer, not the all-rows-at-once one."""
    import mars_b200 as md

    rng = np.random.default_rng(3)
    n = 40_000
    raw = pd.DataFrame({"k": rng.integers(0, 5, n), "v": rng.normal(size=n), "w": rng.random(n)})
    raw.loc[10, ["k", "v"]] = [2, np.inf]
    raw.loc[11, ["k", "v"]] = [2, -np.inf]          # both in chunk 0
    raw.loc[30_000, ["k", "w"]] = [3, np.inf]       # a lone +inf: the sum of group 3 is +inf everywhere
    raw["k"] = raw["k"].astype(np.int64)
    funcs = ["sum", "mean", "count", "min", "max"]
    exp = orc.run_groupby_agg(raw, ["k"], funcs, 10_000, method="tree", sort=True)
    sess = md.new_session(default=False)
    sess.fuse = False                               # every chunk operand on its own, like a Mars worker
    got = md.DataFrame(raw, chunk_size=10_000).to_gpu().groupby("k").agg(funcs, method="tree").execute(session=sess) \
        .fetch(session=sess)
    assert_frames_match(got, exp, sort_index=False)
    assert np.isposinf(got.loc[3, ("w", "sum")]) and np.isfinite(got.loc[2, ("v", "sum")])
    assert got.loc[2, ("v", "min")] == -np.inf and got.loc[2, ("v", "max")] == np.inf


@pytest.mark.parametrize("n", [131_072, 1_048_577, 5_000_003])
def test_pageable_columns_are_uploaded_through_the_pinned_staging_ring(cuda_device, n):
    """uploads from pageable host memory (pandas blocks of a Mars chunk): same bytes on the device as the plain copy,
    for float64 and int64, sizes that are not a multiple of the slab, several columns back to back (slab reuse)"""
    import torch

    from mars_b200 import device as D

    rng = np.random.default_rng(n)
    cols = [rng.random(n), rng.integers(-2 ** 62, 2 ** 62, n, dtype=np.int64), rng.normal(size=n)]
    before = D.h2d_bytes
    ts = [D._to_device(c, cuda_device) for c in cols]
    torch.cuda.synchronize()
    assert D.h2d_bytes - before == 3 * 8 * n
    for t, c in zip(ts, cols):
        assert t.dtype == (torch.float64 if c.dtype == np.float64 else torch.int64)
        np.testing.assert_array_equal(t.cpu().numpy(), c)


@pytest.mark.parametrize("method,sort", [("shuffle", False), ("tree", True), ("shuffle", True)])
def test_nearly_distinct_chunks_hand_their_rows_on_as_partials(md, method, sort):
    """chunks whose keys are ~all distinct: after the first chunk has shown it, the map stage does not group the
    following chunks (the rows are their own partials, COUNT / SUMSQ written per row) -- the final frame is the
    oracle's, NaN values, an all-NaN group and duplicate keys inside a chunk included"""
    from mars_b200 import executors as E

    rng = np.random.default_rng(17)
    n, chunk = 600_000, 100_000
    raw = pd.DataFrame({"key": rng.integers(0, 40_000_000, n), "v": rng.normal(3, 2, n), "w": rng.integers(-50, 50, n)})
    raw.loc[rng.integers(0, n, 5000), "v"] = np.nan
    raw.loc[[5, 250_007, 599_999], "key"] = 123_456_789          # one key in three chunks ...
    raw.loc[[250_008, 250_009], "key"] = 123_456_789             # ... three times in one of them
    raw.loc[[7, 100_007], ["key", "v"]] = [[-5, np.nan], [-5, np.nan]]   # a group without any value
    raw["key"] = raw["key"].astype(np.int64)
    funcs = ["sum", "mean", "count", "min", "max", "var"]
    before = E.passthrough_chunks
    sess = md.new_session(default=False)
    got = md.DataFrame(raw, chunk_size=chunk).to_gpu().groupby("key", sort=sort).agg(funcs, method=method) \
        .execute(session=sess).fetch(session=sess)
    assert E.passthrough_chunks - before >= 4, "the chunks after the first were expected to skip the fold"
    exp = orc.run_groupby_agg(raw, ["key"], funcs, chunk, method=method, sort=sort)
    assert_frames_match(got, exp, sort_index=not sort)


@pytest.mark.parametrize("method,n_groups", [("tree", 40), ("shuffle", 40), ("shuffle", 30_000)])
def test_as_index_false(md, method, n_groups):
    """groupby(..., as_index=False): the keys come back as columns (aggregation.py:1259-1263), for small results
    and for result chunks that stay on the device"""
    rng = np.random.default_rng(5)
    n = 200_000
    raw = pd.DataFrame({"k": rng.integers(0, n_groups, n), "k2": rng.integers(0, 3, n), "v": rng.normal(1, 2, n),
                        "w": rng.integers(0, 9, n)})
    func = {"v": ["sum", "max"], "w": ["count", "mean"]}
    sess = md.new_session(default=False)
    got = md.DataFrame(raw, chunk_size=50_000).to_gpu().groupby(["k", "k2"], as_index=False).agg(func, method=method) \
        .execute(session=sess).fetch(session=sess)
    exp = raw.groupby(["k", "k2"], as_index=False).agg(func)
    got = got.sort_values([("k", ""), ("k2", "")]).reset_index(drop=True)
    assert list(got.columns) == list(exp.columns)
    pd.testing.assert_frame_equal(got, exp.reset_index(drop=True), rtol=1e-9, atol=0, check_exact=False)
