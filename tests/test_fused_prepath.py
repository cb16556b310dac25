"""TPC-H Q1 end to end (reference benchmarks/tpch/run_queries.py:110-163): row filter
on shipdate + two derived columns + groupby(returnflag, linestatus).agg(...).  On the sum / mean / count shape the
filter and the derived columns are evaluated INSIDE the band-level pre-aggregation pass (one launch for the whole
band, no extra pass over lineitem); other shapes materialise them with the elementwise / compaction kernels.
Checked against pandas applying the same operators, through the oracle's stages."""
from __future__ import annotations

import numpy as np
import pandas as pd
import pytest
import torch

import fake_kernels
import mars_groupby_oracle as orc
from conftest import assert_frames_match

BY = ["returnflag", "linestatus"]
Q1_FUNC = {"quantity": ["sum", "mean"], "extendedprice": ["sum", "mean"], "disc_price": ["sum"], "charge": ["sum"],
           "discount": ["mean"], "orderkey": ["count"]}
CUTOFF = 10471          # 1998-09-02 as days since the epoch


def lineitem(n, seed=0):
    rng = np.random.default_rng(seed)
    grp = rng.choice(4, size=n, p=[0.2466, 0.0006, 0.5063, 0.2465])
    return pd.DataFrame({
        "returnflag": np.array([0, 1, 1, 2], dtype=np.int64)[grp], "linestatus": np.array([0, 0, 1, 0], dtype=np.int64)[grp],
        "quantity": rng.integers(1, 51, n).astype(np.float64), "extendedprice": rng.uniform(900.0, 105000.0, n),
        "discount": rng.integers(0, 11, n) / 100.0, "tax": rng.integers(0, 9, n) / 100.0,
        "shipdate": rng.integers(8036, 10562, n, dtype=np.int64), "orderkey": rng.integers(1, 60_000_000, n, dtype=np.int64)})


def q01(md, raw, chunk, func, fuse=True):
    f = md.DataFrame(raw, chunk_size=chunk).to_gpu()
    f = f[f["shipdate"] <= CUTOFF]
    f["disc_price"] = f["extendedprice"] * (1 - f["discount"])
    f["charge"] = f["extendedprice"] * (1 - f["discount"]) * (1 + f["tax"])
    sess = md.new_session(default=False)
    sess.fuse = fuse
    r = f.groupby(BY).agg(func)
    r.execute(session=sess)
    return r.fetch(session=sess), sess


def expected(raw, chunk, func):
    e = raw[raw["shipdate"] <= CUTOFF].copy()
    e["disc_price"] = e["extendedprice"] * (1 - e["discount"])
    e["charge"] = e["extendedprice"] * (1 - e["discount"]) * (1 + e["tax"])
    return orc.run_groupby_agg(e.reset_index(drop=True), BY, func, chunk, method="tree", sort=True)


@pytest.fixture()
def md_cpu(monkeypatch):
    import mars_b200 as m
    from mars_b200 import device, executors

    monkeypatch.setattr(executors, "_kernels", fake_kernels)
    monkeypatch.setattr(device, "_lazy_kernels", fake_kernels)
    monkeypatch.setattr(executors, "_device", lambda: torch.device("cpu"))
    return m


def _check(md, calls=None, n=40_000, chunk=6_000):
    from mars_b200 import executors

    raw = lineitem(n)
    executors._dense_rejected.clear()
    exp = expected(raw, chunk, Q1_FUNC)
    if calls is not None:
        calls.clear()
    got, sess = q01(md, raw, chunk, Q1_FUNC)
    assert_frames_match(got, exp, sort_index=False)
    assert int(got[("orderkey", "count")].sum()) == int((raw["shipdate"] <= CUTOFF).sum())
    if calls is not None:      # one fused launch for the band: no compaction pass, no elementwise pass
        assert calls.count("preagg_dense_batch_fused_async") == 1
        assert "filter_rows" not in calls and "eval_product" not in calls and "preagg_dense_async" not in calls
    # per-operand execution: every map operand is one fused launch over its own chunk
    if calls is not None:
        calls.clear()
    got2, _ = q01(md, raw, chunk, Q1_FUNC, fuse=False)
    assert_frames_match(got2, exp, sort_index=False)
    if calls is not None:
        assert calls.count("preagg_dense_batch_fused_async") == -(-n // chunk) and "filter_rows" not in calls
    # a shape the fused pre-path does not take (min / max): the lazy parts are materialised, same answer
    func = {"quantity": ["min", "max"], "disc_price": ["sum", "max"], "charge": ["mean"]}
    if calls is not None:
        calls.clear()
    got3, _ = q01(md, raw, chunk, func)
    assert_frames_match(got3, expected(raw, chunk, func), sort_index=False)
    if calls is not None:
        assert "filter_rows" in calls and "eval_product" in calls and "preagg_dense_batch_fused_async" not in calls


def test_q1_prepath_host_logic(md_cpu):
    _check(md_cpu, fake_kernels.calls)


def test_expression_algebra_of_the_mirror():
    import mars_b200 as m

    raw = lineitem(10)
    f = m.DataFrame(raw, chunk_size=5)
    e = f["extendedprice"] * (1 - f["discount"]) * (1 + f["tax"])
    assert e.factors == [("extendedprice", 0.0, 1.0), ("discount", 1.0, -1.0), ("tax", 1.0, 1.0)]
    assert (2 * f["tax"] + 3).factors == [("tax", 3.0, 2.0)]
    with pytest.raises(NotImplementedError):
        e * f["quantity"]                      # four factors
    with pytest.raises(NotImplementedError):
        (f["tax"] * f["discount"]) <= 1         # masks compare a stored column


@pytest.mark.gpu
def test_q1_prepath_on_the_gpu(cuda_device):
    import mars_b200 as m
    from mars_b200 import kernels

    _check(m, None, n=1_000_003, chunk=90_000)
    # launch accounting of the fused band: cardinality probe aside, one batch launch + one agg-stage launch
    raw = lineitem(500_000, seed=3)
    q01(m, raw, 60_000, Q1_FUNC)
    before = kernels.launch_counter
    q01(m, raw, 60_000, Q1_FUNC)
    # (+ the dense kernel's immediate exit behind the register-accumulator kernel)
    assert kernels.launch_counter - before in (2, 3)


@pytest.mark.gpu
def test_eval_product_and_filter_rows_match_numpy(cuda_device):
    from mars_b200 import kernels as K

    rng = np.random.default_rng(4)
    n = 300_001
    a, b, c = rng.uniform(900, 105000, n), rng.integers(0, 11, n) / 100.0, rng.integers(0, 9, n) / 100.0
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()   # noqa: E731
    got = K.eval_product([dev(a), dev(b), dev(c)], [(0, 0.0, 1.0), (1, 1.0, -1.0), (2, 1.0, 1.0)]).cpu().numpy()
    np.testing.assert_array_equal(got, a * (1 - b) * (1 + c))             # same operations, same roundings
    d = rng.integers(0, 100, n, dtype=np.int64)
    for op, fn in enumerate([np.less_equal, np.less, np.greater_equal, np.greater, np.equal, np.not_equal]):
        outs, kept = K.filter_rows(dev(d), op, 42, [dev(a), dev(d)])
        m = fn(d, 42)
        assert kept == int(m.sum())
        np.testing.assert_array_equal(outs[0].cpu().numpy(), a[m])
        np.testing.assert_array_equal(outs[1].cpu().numpy(), d[m])
    an = a.copy()
    an[::7] = np.nan
    outs, kept = K.filter_rows(dev(an), 0, 50000.0, [dev(an)])
    assert kept == int((an <= 50000.0).sum())
