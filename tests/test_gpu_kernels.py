"""GPU parity tests of the individual C-ABI entry points (include/mgb.h) against the oracle."""
from __future__ import annotations

import numpy as np
import pandas as pd
import pytest
import torch

import mars_groupby_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K(cuda_device):
    from mars_b200 import kernels

    sms, cc = kernels.device_info()
    assert cc >= 100, f"sm_100a library on compute capability {cc}"
    return kernels


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------ hash
def test_hash_known_answers(K):
    # SURVEY.md Appendix C (pandas docstring / installed pandas)
    h = K.hash_keys([dev(np.array([1, 2, 3], dtype=np.int64))]).cpu().numpy().view(np.uint64)
    assert h.tolist() == [6238072747940578789, 15839785061582574730, 2185194620014831856]
    h = K.hash_keys([dev(np.array([0, -1, 2 ** 63 - 1, -2 ** 63], dtype=np.int64))]).cpu().numpy().view(np.uint64)
    assert h.tolist() == [0, 13029008266876403067, 6514504133438201533, 2720858781877447050]
    h = K.hash_keys([dev(np.array([1, 2, 3], dtype=np.int64)), dev(np.array([10, 20, 30], dtype=np.int64))])
    assert h.cpu().numpy().view(np.uint64).tolist() == [912933651372556365, 2058288944137990103,
                                                         1448569499672317777]
    pid = K.partition_ids([dev(np.array([1, 2, 3], dtype=np.int64))], 8).cpu().numpy()
    assert pid.tolist() == [5, 2, 0]


@pytest.mark.parametrize("n_keys", [1, 2])
@pytest.mark.parametrize("n", [0, 1, 1000, 1 << 20])
def test_hash_matches_pandas_and_oracle(K, n_keys, n):
    rng = np.random.default_rng(n + n_keys)
    cols = [rng.integers(-2 ** 63, 2 ** 63 - 1, n, dtype=np.int64) for _ in range(n_keys)]
    got = K.hash_keys([dev(c) for c in cols]).cpu().numpy().view(np.uint64)
    np.testing.assert_array_equal(got, orc.hash_keys(cols))
    if n:
        idx = pd.Index(cols[0]) if n_keys == 1 else pd.MultiIndex.from_arrays(cols)
        exp = pd.util.hash_pandas_object(idx, categorize=False).to_numpy()
        np.testing.assert_array_equal(got, exp)
    for R in (1, 3, 8, 64, 1000):
        pid = K.partition_ids([dev(c) for c in cols], R).cpu().numpy()
        np.testing.assert_array_equal(pid, orc.partition_ids(cols, R))


# ------------------------------------------------------------------------------------------- partition
@pytest.mark.parametrize("n,R", [(0, 4), (1, 1), (5000, 7), (100000, 64), (300000, 256), (200000, 1000),
                                 (1 << 20, 2)])
def test_partition_scatter_is_stable_split(K, n, R):
    rng = np.random.default_rng(n * 31 + R)
    keys = rng.integers(0, max(n // 3, 1) + 1, n, dtype=np.int64)
    vals = rng.random(n)
    ivals = rng.integers(-5, 5, n, dtype=np.int64)
    pid = orc.partition_ids([keys], R).astype(np.int32)
    outs, offsets = K.partition_scatter(dev(pid), R, [dev(keys), dev(vals), dev(ivals)])
    order = np.argsort(pid, kind="stable")
    exp_off = np.searchsorted(pid[order], np.arange(R + 1)).tolist()
    assert offsets == exp_off
    np.testing.assert_array_equal(outs[0].cpu().numpy(), keys[order])
    np.testing.assert_array_equal(outs[1].cpu().numpy(), vals[order])
    np.testing.assert_array_equal(outs[2].cpu().numpy(), ivals[order])


# ----------------------------------------------------------------------------------------- aggregation
ALL_OPS = ["sum", "sumsq", "count", "min", "max"]


def pandas_agg(keys, vals, accs):
    """expected accumulators via pandas at the reference's call sites (groupby(...).agg)"""
    df = pd.DataFrame({f"k{j}": k for j, k in enumerate(keys)})
    for i, v in enumerate(vals):
        df[f"v{i}"] = v
    g = df.groupby([f"k{j}" for j in range(len(keys))], sort=True)
    out = []
    for col, op in accs:
        s = g[f"v{col}"]
        if op == "sumsq":
            d2 = df.copy()
            d2[f"v{col}"] = d2[f"v{col}"] ** 2
            r = d2.groupby([f"k{j}" for j in range(len(keys))], sort=True)[f"v{col}"].sum()
        else:
            r = getattr(s, op)()
        out.append(r)
    idx = out[0].index
    key_arrays = [idx.get_level_values(j).to_numpy() for j in range(idx.nlevels)]
    return key_arrays, [r.to_numpy() for r in out]


def run_agg(K, keys, vals, accs, sort=True, batches=1, distinct=False):
    code = dict(sum=K.SUM, sumsq=K.SUMSQ, count=K.COUNT, min=K.MIN, max=K.MAX)
    agg = K.Aggregator(len(keys), [K.F64 if v.dtype == np.float64 else K.I64 for v in vals],
                       [(c, code[o]) for c, o in accs], expect_distinct=distinct)
    n = len(keys[0])
    bounds = np.linspace(0, n, batches + 1).astype(int)
    for b in range(batches):
        s, e = bounds[b], bounds[b + 1]
        agg.update([dev(k[s:e]) for k in keys], [dev(v[s:e]) for v in vals])
    k, a = agg.result(sort=sort)
    stats = agg.stats
    agg.close()
    return [t.cpu().numpy() for t in k], [t.cpu().numpy() for t in a], stats


def check_agg(K, keys, vals, accs, batches=1, expect_path=None, distinct=False):
    gk, ga, stats = run_agg(K, keys, vals, accs, sort=True, batches=batches, distinct=distinct)
    ek, ea = pandas_agg(keys, vals, accs)
    for a, b in zip(gk, ek):
        np.testing.assert_array_equal(a, b)
    for (col, op), a, b in zip(accs, ga, ea):
        if op in ("count", "min", "max") or b.dtype.kind in "iu":
            np.testing.assert_array_equal(a, b.astype(a.dtype), err_msg=f"{op} of column {col}")
        else:
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=0, equal_nan=True, err_msg=f"{op} of column {col}")
    if expect_path is not None:
        assert stats[expect_path] > 0, stats
    # unsorted emit returns the same rows in table order
    uk, ua, _ = run_agg(K, keys, vals, accs, sort=False, batches=batches, distinct=distinct)
    order = np.lexsort(tuple(reversed(uk)))
    for a, b in zip(uk, gk):
        np.testing.assert_array_equal(a[order], b)
    for a, b in zip(ua, ga):
        if a.dtype.kind in "iu":
            np.testing.assert_array_equal(a[order], b)
        else:
            np.testing.assert_allclose(a[order], b, rtol=1e-9, equal_nan=True)   # two runs: atomics / fold order differ
    return stats


@pytest.mark.parametrize("n_keys", [1, 2])
@pytest.mark.parametrize("n,card,path", [
    (1, 1, None), (1000, 3, "rows_smem"), (200000, 4, "rows_smem"), (200000, 1000, "rows_smem"),
    (500000, 50000, None), (300000, 300000, None)])
def test_aggregate_matches_pandas(K, n_keys, n, card, path):
    rng = np.random.default_rng(n + card + n_keys)
    if n_keys == 1:
        keys = [rng.integers(-(card // 2), card - card // 2, n, dtype=np.int64)]
    else:
        c1 = max(int(np.sqrt(card)), 1)
        keys = [rng.integers(0, c1, n, dtype=np.int64), rng.integers(-3, max(card // c1, 1) - 3, n, dtype=np.int64)]
    vals = [rng.normal(10, 3, n), rng.random(n), rng.integers(-100, 100, n, dtype=np.int64)]
    accs = [(0, o) for o in ALL_OPS] + [(1, "sum"), (1, "count"), (2, "sum"), (2, "min"), (2, "max"), (2, "count")]
    check_agg(K, keys, vals, accs, expect_path=path)


def test_aggregate_many_batches_and_wide(K):
    rng = np.random.default_rng(5)
    n = 120000
    keys = [rng.integers(0, 300, n, dtype=np.int64)]
    vals = [rng.normal(10, 3, n) for _ in range(11)]      # > 8 value columns: two column groups
    accs = [(c, o) for c in range(11) for o in ("sum", "min", "max")]
    check_agg(K, keys, vals, accs, batches=5)


def test_aggregate_nan_and_marker_keys(K):
    rng = np.random.default_rng(6)
    n = 50000
    keys = [rng.choice(np.array([0, -1, 2 ** 63 - 1, -2 ** 63, 7, 1 << 40], dtype=np.int64), n)]
    v0 = rng.random(n)
    v0[rng.random(n) < 0.3] = np.nan
    v1 = rng.normal(0, 1, n)
    v1[keys[0] == 7] = np.nan
    accs = [(0, o) for o in ALL_OPS] + [(1, o) for o in ALL_OPS]
    check_agg(K, keys, [v0, v1], accs)
    # two keys, both equal to the empty marker (INT64_MIN)
    k1 = rng.choice(np.array([-2 ** 63, 0, 5], dtype=np.int64), n)
    k2 = rng.choice(np.array([-2 ** 63, 1], dtype=np.int64), n)
    check_agg(K, [k1, k2], [v0, v1], accs)
    # many distinct keys including the marker -> global table path
    k = rng.integers(-2 ** 63, 2 ** 63 - 1, n, dtype=np.int64)
    k[::97] = -2 ** 63
    check_agg(K, [k], [v0, v1], accs)


@pytest.mark.parametrize("n_keys", [1, 2])
@pytest.mark.parametrize("n,card", [(1000, 5), (300_000, 1000), (300_000, 2500), (400_000, 20_000), (250_000, 250_000)])
def test_aggregate_plain_shapes_take_the_lean_hash_kernel(K, n_keys, n, card):
    """sum / mean / count shapes (float64 columns, SUM + COUNT): k_preagg_hash, incl. NaNs (spilled rows),
    a frozen table (more groups than slots) and marker keys"""
    rng = np.random.default_rng(n + card + n_keys)
    keys = [rng.integers(-(card // 2), card - card // 2, n, dtype=np.int64)]
    if n_keys == 2:
        keys.append(rng.integers(0, 3, n, dtype=np.int64))
    keys[0][::101] = -2 ** 63
    vals = [rng.normal(10, 3, n), rng.random(n), rng.random(n)]
    vals[1][rng.random(n) < 0.05] = np.nan
    accs = [(0, "sum"), (1, "sum"), (2, "sum"), (0, "count"), (1, "count")]
    stats = check_agg(K, keys, vals, accs, batches=2)
    assert stats["rows_smem"] > 0


def test_aggregate_skewed_keys(K):
    rng = np.random.default_rng(7)
    n = 400000
    keys = [np.minimum(rng.zipf(1.2, n), 10 ** 8).astype(np.int64)]
    vals = [rng.random(n)]
    check_agg(K, keys, vals, [(0, "sum"), (0, "count")])


def test_empty_aggregator(K):
    agg = K.Aggregator(1, [K.F64], [(0, K.SUM)])
    k, a = agg.result(sort=True)
    assert k[0].numel() == 0 and a[0].numel() == 0
    agg.close()


# ------------------------------------------------------------------------------------------------ post
def test_post_functions_follow_reference_operation_order(K):
    rng = np.random.default_rng(8)
    n = 10000
    c = rng.integers(1, 50, n)
    s = rng.normal(10, 3, n) * c
    q = s * s / c + rng.random(n) * c
    got = K.post_mean(dev(s), dev(c)).cpu().numpy()
    np.testing.assert_array_equal(got, s / c)
    for ddof in (1, 0):
        got = K.post_var(dev(s), dev(q), dev(c), ddof).cpu().numpy()
        exp = orc.post_var(pd.Series(q), pd.Series(s), pd.Series(c), ddof=ddof).to_numpy()
        np.testing.assert_array_equal(got.view(np.int64), exp.view(np.int64))   # bit-exact: same op order, no FMA


# ----------------------------------------------------------------------------------------- sort / concat
@pytest.mark.parametrize("n_keys", [1, 2])
def test_argsort_and_gather(K, n_keys):
    rng = np.random.default_rng(9)
    n = 70000
    cols = [rng.integers(-2 ** 62, 2 ** 62, n, dtype=np.int64) if j == 0 else rng.integers(-4, 4, n, dtype=np.int64)
            for j in range(n_keys)]
    if n_keys == 2:
        cols[0] = rng.integers(-50, 50, n, dtype=np.int64)
    perm = K.argsort_keys([dev(c) for c in cols]).cpu().numpy()
    exp = np.lexsort(tuple(reversed(cols)))
    np.testing.assert_array_equal(perm, exp)      # LSD radix sort is stable, like lexsort
    vals = rng.random(n)
    out = K.gather_rows([dev(vals)], dev(perm.astype(np.int32)))[0].cpu().numpy()
    np.testing.assert_array_equal(out, vals[exp])


def test_concat_columns(K):
    a, b, c = np.arange(5.0), np.arange(0.0), np.arange(7.0) * 2
    out = K.concat_columns([dev(a), dev(b), dev(c)]).cpu().numpy()
    np.testing.assert_array_equal(out, np.concatenate([a, b, c]))


def test_cpu_tensor_is_rejected_loudly(K):
    from mars_b200 import MarsGBError

    with pytest.raises(MarsGBError):
        K.hash_keys([torch.arange(4, dtype=torch.int64)])


# --------------------------------------------------------------------------- low-cardinality fast paths
def _codes(K, accs):
    code = dict(sum=K.SUM, sumsq=K.SUMSQ, count=K.COUNT, min=K.MIN, max=K.MAX)
    return [(c, code[o]) for c, o in accs]


def check_dense(K, keys, vals, accs, expect_fit=True):
    res = K.preagg_dense([dev(k) for k in keys], [dev(v) for v in vals], _codes(K, accs))
    if not expect_fit:
        assert res is None
        return
    assert res is not None, "dense path rejected a shape that fits"
    gk = [t.cpu().numpy() for t in res[0]]
    ga = [t.cpu().numpy() for t in res[1]]
    ek, ea = pandas_agg(keys, vals, accs)
    order = np.lexsort(tuple(reversed(gk)))
    for a, b in zip(gk, ek):
        np.testing.assert_array_equal(a[order], b)
    for (col, op), a, b in zip(accs, ga, ea):
        a = a[order]
        assert a.dtype == (np.int64 if (op == "count" or b.dtype.kind in "iu") else np.float64)
        if op in ("count", "min", "max") or b.dtype.kind in "iu":
            np.testing.assert_array_equal(a, b.astype(a.dtype), err_msg=f"{op} of column {col}")
        else:
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=0, equal_nan=True, err_msg=f"{op} of column {col}")


@pytest.mark.parametrize("n", [1, 255, 1024, 100_003, 3_000_000])
@pytest.mark.parametrize("n_keys", [1, 2])
def test_dense_sum_count_shapes(K, n, n_keys):
    rng = np.random.default_rng(n + n_keys)
    keys = [rng.integers(0, 3, n, dtype=np.int64)] + ([rng.integers(-1, 1, n, dtype=np.int64)] if n_keys == 2 else [])
    vals = [rng.random(n), rng.normal(5, 2, n), rng.integers(1, 10 ** 6, n, dtype=np.int64), rng.random(n),
            rng.random(n), rng.random(n)]
    vals[3][rng.random(n) < 0.1] = np.nan
    accs = [(0, "sum"), (1, "sum"), (3, "sum"), (4, "sum"), (5, "sum"), (0, "count"), (1, "count"), (3, "count"),
            (2, "count"), (2, "sum")]
    check_dense(K, keys, vals, accs)


@pytest.mark.parametrize("ncol", [1, 2, 3, 5, 8])
def test_dense_all_ops(K, ncol):
    rng = np.random.default_rng(ncol)
    n = 200_000
    # all five ops on 8 columns leave room for 2 groups per CTA; narrower shapes hold more
    keys = [rng.integers(-2, 2 if ncol < 8 else 0, n, dtype=np.int64)]
    vals = [rng.normal(10, 3, n) if c % 3 else rng.integers(-1000, 1000, n, dtype=np.int64) for c in range(ncol)]
    if ncol > 1:
        vals[1][rng.random(n) < 0.2] = np.nan
        vals[1][keys[0] == 1] = np.nan     # an all-NaN group: sum 0, count 0, min/max NaN
    accs = [(c, o) for c in range(ncol) for o in ALL_OPS]
    if len(accs) > 64:
        accs = accs[:64]
    check_dense(K, keys, vals, accs)


def test_dense_rejects_high_cardinality_and_marker_keys(K):
    rng = np.random.default_rng(1)
    n = 50_000
    check_dense(K, [rng.integers(0, 500, n, dtype=np.int64)], [rng.random(n)], [(0, "sum")], expect_fit=False)
    # INT64_MIN / extreme keys are ordinary keys on the dense path
    keys = [rng.choice(np.array([-2 ** 63, 2 ** 63 - 1, 0], dtype=np.int64), n)]
    check_dense(K, keys, [rng.random(n)], [(0, "sum"), (0, "count"), (0, "min"), (0, "max")])
    # exactly at the per-CTA group capacity boundary: a handful of groups, one of them very rare
    keys = [rng.integers(0, 6, n, dtype=np.int64)]
    keys[0][n // 2] = 77
    check_dense(K, keys, [rng.random(n)], [(0, "sum"), (0, "count")])


@pytest.mark.parametrize("n_keys", [1, 2])
@pytest.mark.parametrize("sort", [False, True])
def test_merge_small_matches_pandas(K, n_keys, sort):
    rng = np.random.default_rng(3 + n_keys)
    pieces_k, pieces_a, frames = [], [], []
    for i, n in enumerate([7, 0, 300, 1, 900]):
        ks = [rng.integers(-20, 20, n, dtype=np.int64) for _ in range(n_keys)]
        s = rng.normal(3, 1, n)
        c = rng.integers(0, 5, n, dtype=np.int64)
        mn = rng.normal(0, 1, n)
        mn[rng.random(n) < 0.3] = np.nan
        mx = rng.integers(-9, 9, n, dtype=np.int64)
        pieces_k.append([dev(k) for k in ks])
        pieces_a.append([dev(s), dev(c), dev(mn), dev(mn.copy()), dev(mx)])
        frames.append(pd.DataFrame(dict({f"k{j}": k for j, k in enumerate(ks)}, s=s, c=c, mn=mn, mx2=mn, mx=mx)))
    res = K.merge_small(pieces_k, pieces_a, [K.SUM, K.SUM, K.MIN, K.MAX, K.MAX], sort=sort)
    assert res is not None
    gk = [t.cpu().numpy() for t in res[0]]
    ga = [t.cpu().numpy() for t in res[1]]
    exp = pd.concat(frames).groupby([f"k{j}" for j in range(n_keys)], sort=True).agg(
        s=("s", "sum"), c=("c", "sum"), mn=("mn", "min"), mx2=("mx2", "max"), mx=("mx", "max"))
    order = np.lexsort(tuple(reversed(gk)))
    if sort:
        np.testing.assert_array_equal(order, np.arange(len(order)))
    for j in range(n_keys):
        np.testing.assert_array_equal(gk[j][order], exp.index.get_level_values(j).to_numpy())
    np.testing.assert_allclose(ga[0][order], exp["s"].to_numpy(), rtol=1e-12)
    np.testing.assert_array_equal(ga[1][order], exp["c"].to_numpy())
    np.testing.assert_array_equal(ga[2][order], exp["mn"].to_numpy())
    np.testing.assert_array_equal(ga[3][order], exp["mx2"].to_numpy())
    np.testing.assert_array_equal(ga[4][order], exp["mx"].to_numpy())


def test_merge_small_declines_large_inputs(K):
    k = dev(np.arange(3000, dtype=np.int64))
    assert K.merge_small([[k]], [[dev(np.ones(3000))]], [K.SUM]) is None
    res = K.merge_small([[k[:0]]], [[dev(np.ones(0))]], [K.SUM])
    assert res is not None and res[0][0].numel() == 0


def test_to_host_single_sync(K):
    a, b = np.arange(10.0), np.arange(10, dtype=np.int64)
    out = K.to_host([dev(a), dev(b)])
    np.testing.assert_array_equal(out[0], a)
    np.testing.assert_array_equal(out[1], b)
    assert out[0].dtype == np.float64 and out[1].dtype == np.int64


@pytest.mark.parametrize("plain", [True, False])
def test_aggregate_with_distinct_keys_hint_skips_preaggregation(K, plain):
    rng = np.random.default_rng(31)
    n = 300_000
    keys = [rng.integers(-2 ** 62, 2 ** 62, n, dtype=np.int64)]
    keys[0][::13] = keys[0][0]         # some duplicates all the same (one bucket of many tiles)
    vals = [rng.normal(0, 1, n), rng.random(n)]
    vals[1][rng.random(n) < 0.1] = np.nan
    accs = [(0, "sum"), (1, "sum"), (0, "count"), (1, "count")] + ([] if plain else [(1, "min"), (0, "max"), (0, "sumsq")])
    code = dict(sum=K.SUM, sumsq=K.SUMSQ, count=K.COUNT, min=K.MIN, max=K.MAX)
    k, a, stats = K.groupby_aggregate([dev(x) for x in keys], [dev(v) for v in vals], [(c, code[o]) for c, o in accs],
                                      sort=True, expect_distinct=True)
    assert stats["rows_smem"] == 0 and stats["rows_spilled"] == n and stats["table_capacity"] == 0
    ek, ea = pandas_agg(keys, vals, accs)
    np.testing.assert_array_equal(k[0].cpu().numpy(), ek[0])
    for (col, op), got, exp in zip(accs, a, ea):
        got = got.cpu().numpy()
        if op in ("count", "min", "max"):
            np.testing.assert_array_equal(got, exp.astype(got.dtype))
        else:
            np.testing.assert_allclose(got, exp, rtol=1e-9, atol=0, equal_nan=True)


# ------------------------------------------------------------------------ bucketed aggregation (mgb_bucket.cu)
def _bucketed(stats):
    """the distinct-keys hint was honoured by the bucketed path: no global table was built"""
    return stats["table_capacity"] == 0 and stats["rows_spilled"] > 0


@pytest.mark.parametrize("n_keys", [1, 2])
@pytest.mark.parametrize("n,card,batches", [(5000, 4000, 1), (300_000, 300_000, 1), (300_000, 40_000, 7),
                                            (1_200_000, 150_000, 3), (200_000, 13, 2)])
def test_bucketed_aggregation_matches_pandas(K, n_keys, n, card, batches):
    """inputs that are not pre-aggregated (distinct-keys hint): all ops, both dtypes, NaNs, several sources;
    card=13 makes every bucket one long multi-tile run of a few big groups (warp-per-group fold)"""
    rng = np.random.default_rng(n + card + n_keys)
    if n_keys == 1:
        keys = [rng.integers(-(card // 2), card - card // 2, n, dtype=np.int64)]
    else:
        c1 = max(int(np.sqrt(card)), 1)
        keys = [rng.integers(0, c1, n, dtype=np.int64), rng.integers(-3, max(card // c1, 1) - 3, n, dtype=np.int64)]
    keys[0][::1001] = -2 ** 63
    vals = [rng.normal(10, 3, n), rng.random(n), rng.integers(-100, 100, n, dtype=np.int64)]
    vals[1][rng.random(n) < 0.2] = np.nan
    accs = [(0, o) for o in ALL_OPS] + [(1, o) for o in ALL_OPS] + [(2, "sum"), (2, "min"), (2, "max"), (2, "count")]
    stats = check_agg(K, keys, vals, accs, batches=batches, distinct=True)
    assert _bucketed(stats), stats


def test_bucketed_aggregation_wide_min_max_var_shape(K):
    """cfg4-shaped: two keys, 8 float64 columns x (min, max, sum, sumsq, count) = 40 accumulators"""
    rng = np.random.default_rng(44)
    n = 400_000
    keys = [rng.integers(0, 300, n, dtype=np.int64), rng.integers(0, 200, n, dtype=np.int64)]
    vals = [rng.normal(10, 3, n) for _ in range(8)]
    accs = [(c, o) for o in ("min", "max", "sum", "sumsq", "count") for c in range(8)]
    stats = check_agg(K, keys, vals, accs, batches=2, distinct=True)
    assert _bucketed(stats), stats


def test_bucketed_aggregation_many_columns(K):
    """24 value columns: the row-major staging tile of the pack kernel needs more than 48 KB of shared memory
    (opt-in attribute), and the fold walks 24 (group, column) items per group"""
    rng = np.random.default_rng(48)
    n = 150_000
    keys = [rng.integers(0, 60_000, n, dtype=np.int64)]
    vals = [rng.normal(10, 3, n) if c % 3 else rng.integers(-50, 50, n, dtype=np.int64) for c in range(24)]
    accs = [(c, "sum") for c in range(24)] + [(c, "max") for c in range(24)]
    stats = check_agg(K, keys, vals, accs, batches=3, distinct=True)
    assert _bucketed(stats), stats


def test_bucketed_aggregation_skew_and_merge_ops(K):
    """one key holds 4 % of the rows (a bucket of many tiles, folded by warps, merged across tiles);
    reducer-side merge shape: int64 count partials summed, float64 min / max partials with NaN = no value"""
    rng = np.random.default_rng(45)
    n = 600_000
    k = rng.integers(0, 200_000, n, dtype=np.int64)
    k[rng.random(n) < 0.04] = 77
    s = rng.normal(10, 3, n)         # positive mean: the hot key's sum must not cancel (1e-9 relative bar)
    c = rng.integers(0, 50, n, dtype=np.int64)
    mn = rng.random(n)
    mn[rng.random(n) < 0.3] = np.nan
    mn[k == 5] = np.nan
    accs = [(0, "sum"), (1, "sum"), (2, "min"), (2, "max")]
    stats = check_agg(K, [k], [s, c, mn], accs, batches=16, distinct=True)
    assert _bucketed(stats), stats


def test_bucketed_aggregation_declines_heavily_skewed_keys(K):
    """a key with a third of the rows would serialise its bucket on one CTA: the input goes to the hash table
    and the statistics say why (the executors then stop giving the hint for this shape)"""
    rng = np.random.default_rng(47)
    n = 600_000
    k = rng.integers(0, 200_000, n, dtype=np.int64)
    k[rng.random(n) < 0.33] = 77
    stats = check_agg(K, [k], [rng.normal(10, 3, n)], [(0, "sum"), (0, "count"), (0, "min")], distinct=True)
    assert stats["bucket_fallback"] == 2 and stats["table_capacity"] > 0


def test_bucketed_aggregation_overflow_falls_back_to_the_table(K):
    """more groups in one bucket than its on-chip table holds: the library reports nothing wrong, it takes
    the hash-table path (keys chosen so that the top bits of their slot hash collide)"""
    cand = np.arange(1, 3_000_000, dtype=np.int64)
    h = orc.hash_keys([cand]) * np.uint64(0x9e3779b97f4a7c15)
    chosen = cand[(h >> np.uint64(64 - 9)) == 0][:2000]
    assert len(chosen) == 2000
    rng = np.random.default_rng(46)
    # 200k rows -> 512 buckets; bucket 0 gets 2000 groups x 15 rows (more groups than its table, fewer rows
    # than the skew limit), the other rows spread evenly
    keys = [np.concatenate([np.repeat(chosen, 15), rng.integers(10_000_000, 20_000_000, 170_000, dtype=np.int64)])]
    rng.shuffle(keys[0])
    vals = [rng.random(len(keys[0]))]
    stats = check_agg(K, keys, vals, [(0, "sum"), (0, "count"), (0, "max")], distinct=True)
    assert stats["table_capacity"] > 0 and stats["bucket_fallback"] == 1


@pytest.mark.parametrize("n_keys", [1, 2])
def test_range_partition_ids_match_oracle(K, n_keys):
    rng = np.random.default_rng(41 + n_keys)
    n = 100_000
    cols = [rng.integers(-1000, 1000, n, dtype=np.int64) for _ in range(n_keys)]
    cols[0][:4] = [-2 ** 63, 2 ** 63 - 1, 0, -1]
    if n_keys == 1:
        pivots = np.array([-2 ** 63, -500, -500, 0, 17, 999, 2 ** 63 - 1], dtype=np.int64)
        pcols = [pivots]
    else:
        tuples = sorted({(int(a), int(b)) for a, b in zip(rng.integers(-1000, 1000, 9), rng.integers(-1000, 1000, 9))})
        tuples.insert(3, tuples[3])                     # a duplicated pivot: an empty partition
        pivots = np.empty(len(tuples), dtype=object)
        pivots[:] = tuples
        pcols = [np.array([t[j] for t in tuples], dtype=np.int64) for j in range(2)]
    got = K.range_partition_ids([dev(c) for c in cols], [dev(p) for p in pcols]).cpu().numpy()
    np.testing.assert_array_equal(got, orc.range_partition_ids(cols, pivots))


# ------------------------------------------------------------------- band-level batch / one-launch agg stage
def _batch(K, chunks_keys, chunks_vals, accs, skip_cols=()):
    """mgb_preagg_dense_batch_async over host chunks -> (n, key arrays, accumulator arrays) read back"""
    dk = [[dev(k) for k in ks] for ks in chunks_keys]
    dv = [[None if c in skip_cols else dev(v) for c, v in enumerate(vs)] for vs in chunks_vals]
    dts = [K.F64 if v.dtype == np.float64 else K.I64 for v in chunks_vals[0]]
    chunks = [([k.data_ptr() for k in ks], [None if v is None else v.data_ptr() for v in vs], len(chunks_keys[i][0]))
              for i, (ks, vs) in enumerate(zip(dk, dv))]
    res = K.preagg_dense_batch_async(chunks, tuple(_codes(K, accs)), dts, len(chunks_keys[0]), torch.device("cuda", 0))
    if res is None:
        return None
    bi, bf, fc = res
    torch.cuda.synchronize()
    n = int(bi[fc.n_int, 0])
    if n < 0:
        return n, None, None
    k, a = fc.views(bi, bf, n)
    return n, [t.cpu().numpy() for t in k], [t.cpu().numpy() for t in a]


@pytest.mark.parametrize("n_keys", [1, 2])
@pytest.mark.parametrize("sizes", [[5000, 0, 1024, 777, 300_000, 1, 2048], [4_194_304, 1_265_796], [10] * 64, [0, 0]])
def test_dense_batch_streams_all_chunks_in_one_launch(K, n_keys, sizes):
    """ragged chunk sizes (empty, shorter than a tile, exact multiples of the tile): one launch == pandas on the
    concatenation; sum / mean / count shape with NaNs and a count-only int64 column that is never uploaded"""
    rng = np.random.default_rng(len(sizes) + n_keys)
    ck, cv = [], []
    for n in sizes:
        ck.append([rng.integers(0, 3, n, dtype=np.int64)] + ([rng.integers(-1, 1, n, dtype=np.int64)] if n_keys == 2 else []))
        v = [rng.random(n), rng.normal(5, 2, n), rng.integers(1, 10 ** 6, n, dtype=np.int64)]
        v[1][rng.random(n) < 0.1] = np.nan
        cv.append(v)
    accs = [(0, "sum"), (1, "sum"), (0, "count"), (1, "count"), (2, "count")]
    before = K.launch_counter
    got = _batch(K, ck, cv, accs, skip_cols=(2,))
    assert K.launch_counter - before == (2 if sum(sizes) else 0)   # register-accumulator kernel + the dense one behind it
    n, gk, ga = got
    keys = [np.concatenate([c[j] for c in ck]) for j in range(n_keys)]
    vals = [np.concatenate([c[j] for c in cv]) for j in range(3)]
    if sum(sizes) == 0:
        assert n == 0
        return
    ek, ea = pandas_agg(keys, vals, accs)
    assert n == len(ek[0])
    order = np.lexsort(tuple(reversed(gk)))
    for a, b in zip(gk, ek):
        np.testing.assert_array_equal(a[order], b)
    for (col, op), a, b in zip(accs, ga, ea):
        if op == "count":
            np.testing.assert_array_equal(a[order], b)
        else:
            np.testing.assert_allclose(a[order], b, rtol=1e-9, atol=0)


def test_dense_batch_all_ops_and_overflow(K):
    rng = np.random.default_rng(9)
    sizes = [70_000, 999, 130_000]
    ck = [[rng.integers(-2, 3, n, dtype=np.int64)] for n in sizes]
    cv = [[rng.normal(10, 3, n), rng.integers(-1000, 1000, n, dtype=np.int64)] for n in sizes]
    accs = [(c, o) for c in range(2) for o in ALL_OPS]
    n, gk, ga = _batch(K, ck, cv, accs)
    ek, ea = pandas_agg([np.concatenate([c[0] for c in ck])], [np.concatenate([c[j] for c in cv]) for j in range(2)], accs)
    order = np.lexsort(tuple(reversed(gk)))
    np.testing.assert_array_equal(gk[0][order], ek[0])
    for (col, op), a, b in zip(accs, ga, ea):
        if op in ("count", "min", "max") or b.dtype.kind in "iu":
            np.testing.assert_array_equal(a[order], b.astype(a.dtype), err_msg=f"{op} of column {col}")
        else:
            np.testing.assert_allclose(a[order], b, rtol=1e-9, atol=0, err_msg=f"{op} of column {col}")
    # a chunk in the middle of the band with too many groups poisons the whole launch (-1), nothing else
    ck[1] = [rng.integers(0, 900, sizes[1], dtype=np.int64)]
    n, _, _ = _batch(K, ck, cv, accs)
    assert n == -1


@pytest.mark.parametrize("sort", [False, True])
def test_merge_small_post_is_the_whole_agg_stage(K, sort):
    """merge of partial pieces (sizes on the device) + identity / mean / var post functions in one launch"""
    rng = np.random.default_rng(17)
    frames, pieces, keep = [], [], []
    for n in [40, 0, 700, 3]:
        k = rng.integers(-30, 30, n, dtype=np.int64)
        s = rng.normal(50, 5, n)
        q = s * s + rng.random(n)
        c = rng.integers(1, 9, n, dtype=np.int64)
        si = rng.integers(-50, 50, n, dtype=np.int64)
        frames.append(pd.DataFrame(dict(k=k, s=s, q=q, c=c, si=si)))
        # capacity-length columns, like the blocks of the speculative chain: the row count lives on the device
        pad = lambda a: dev(np.concatenate([a, np.zeros(8, dtype=a.dtype)]))   # noqa: E731
        cols = [pad(k), pad(s), pad(q), pad(c), pad(si)]
        cnt = dev(np.array([n], dtype=np.int64))
        keep.append((cols, cnt))
        pieces.append(([cols[0].data_ptr()], [t.data_ptr() for t in cols[1:]], None, cnt.data_ptr(),
                       [K.F64, K.F64, K.I64, K.I64], torch.device("cuda", 0)))
    steps = [(K.POST_IDENTITY, 0, 0, 0, 0), (K.POST_MEAN, 0, 2, 0, 0), (K.POST_VAR, 0, 1, 2, 1), (K.POST_VAR, 0, 1, 2, 0),
             (K.POST_IDENTITY, 2, 0, 0, 0), (K.POST_MEAN, 3, 2, 0, 0)]
    post = K.PostPlan(steps, [True, True, True, True, False, True])
    blk, nk, npost = K.merge_small_post_async(pieces, [K.SUM, K.SUM, K.SUM, K.SUM], post, sort=sort)
    n, keys, cols = K.result_block_to_host(blk, nk, npost, post.is_float)
    m = pd.concat(frames).groupby("k", sort=True).sum()
    assert n == len(m)
    order = np.argsort(keys[0], kind="stable")
    if sort:
        np.testing.assert_array_equal(order, np.arange(n))
    np.testing.assert_array_equal(keys[0][order], m.index.to_numpy())
    S, Q, Cn, SI = m["s"], m["q"], m["c"], m["si"]
    np.testing.assert_allclose(cols[0][order], S.to_numpy(), rtol=1e-12)
    np.testing.assert_allclose(cols[1][order], (S / Cn).to_numpy(), rtol=1e-12)
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(cols[2][order], orc.post_var(Q, S, Cn, ddof=1).to_numpy(), rtol=1e-9, equal_nan=True)
        np.testing.assert_allclose(cols[3][order], orc.post_var(Q, S, Cn, ddof=0).to_numpy(), rtol=1e-9, equal_nan=True)
    np.testing.assert_array_equal(cols[4][order], Cn.to_numpy())
    np.testing.assert_allclose(cols[5][order], (SI / Cn).to_numpy(), rtol=1e-15)
    # a piece whose device count says "failed" (-1) poisons the result; too many rows are declined on the host
    bad = dev(np.array([-1], dtype=np.int64))
    pieces[0] = pieces[0][:3] + (bad.data_ptr(),) + pieces[0][4:]
    blk, nk, npost = K.merge_small_post_async(pieces, [K.SUM] * 4, post, sort=sort)
    assert K.result_block_to_host(blk, nk, npost, post.is_float)[0] == -1
    big = (pieces[1][0], pieces[1][1], 5000, 0) + pieces[1][4:]
    assert K.merge_small_post_async([big], [K.SUM] * 4, post) is None


# ------------------------------------------------------------------------------- mid-cardinality batch kernel
def _mid(K, chunks_keys, chunks_vals, accs, skip_cols=()):
    dk = [[dev(k) for k in ks] for ks in chunks_keys]
    dv = [[None if c in skip_cols else dev(v) for c, v in enumerate(vs)] for vs in chunks_vals]
    dts = [K.F64 if v.dtype == np.float64 else K.I64 for v in chunks_vals[0]]
    chunks = [([k.data_ptr() for k in ks], [None if v is None else v.data_ptr() for v in vs], len(chunks_keys[i][0]))
              for i, (ks, vs) in enumerate(zip(dk, dv))]
    bi, bf, fc = K.preagg_mid_batch_async(chunks, tuple(_codes(K, accs)), dts, len(chunks_keys[0]), torch.device("cuda", 0))
    torch.cuda.synchronize()
    n = int(bi[fc.n_int, 0])
    if n < 0:
        return n, None, None
    k, a = fc.views(bi, bf, n)
    return n, [t.cpu().numpy() for t in k], [t.cpu().numpy() for t in a]


@pytest.mark.parametrize("n_keys", [1, 2])
@pytest.mark.parametrize("ncol,card,sizes", [(4, 1000, [1_000_000, 777, 0, 300_000]), (1, 1500, [50_000]),
                                             (8, 40, [200_000, 1024]), (2, 1500, [3_000_000])])
def test_mid_batch_matches_pandas(K, n_keys, ncol, card, sizes):
    rng = np.random.default_rng(card + ncol + n_keys)
    ck, cv = [], []
    c1 = max(1, int(np.sqrt(card)))
    for n in sizes:
        if n_keys == 1:
            ck.append([rng.integers(-card // 2, card - card // 2, n, dtype=np.int64)])
        else:
            ck.append([rng.integers(0, c1, n, dtype=np.int64), rng.integers(0, card // c1, n, dtype=np.int64)])
        cv.append([rng.normal(5, 2, n) for _ in range(ncol)] + [rng.integers(0, 99, n, dtype=np.int64)])
    accs = [(c, "sum") for c in range(ncol)] + [(0, "count"), (ncol, "count")]
    before = K.launch_counter
    n, gk, ga = _mid(K, ck, cv, accs, skip_cols=(ncol,))
    assert K.launch_counter - before == 2           # the batch kernel + the compaction of its global table
    keys = [np.concatenate([c[j] for c in ck]) for j in range(n_keys)]
    vals = [np.concatenate([c[j] for c in cv]) for j in range(ncol + 1)]
    ek, ea = pandas_agg(keys, vals, accs)
    assert n == len(ek[0])
    order = np.lexsort(tuple(reversed(gk)))
    for a, b in zip(gk, ek):
        np.testing.assert_array_equal(a[order], b)
    for (col, op), a, b in zip(accs, ga, ea):
        if op == "count":
            np.testing.assert_array_equal(a[order], b)
        else:
            np.testing.assert_allclose(a[order], b, rtol=1e-9, atol=0)


def test_mid_batch_declines_what_it_does_not_take(K):
    rng = np.random.default_rng(5)
    n = 200_000
    k = [rng.integers(0, 500, n, dtype=np.int64)]
    v = [rng.random(n)]
    assert _mid(K, [k], [v], [(0, "sum"), (0, "count")])[0] == 500
    assert _mid(K, [k], [v], [(0, "sum"), (0, "max")])[0] == -1                       # not a sum / count shape
    assert _mid(K, [[rng.integers(0, 5000, n, dtype=np.int64)]], [v], [(0, "sum")])[0] == -1   # too many groups
    vn = [v[0].copy()]
    vn[0][12345] = np.nan
    assert _mid(K, [k], [vn], [(0, "sum"), (0, "count")])[0] == -1                    # NaN: general path keeps the books
    vi = [rng.integers(0, 9, n, dtype=np.int64)]
    assert _mid(K, [k], [vi], [(0, "sum")])[0] == -1                                  # int64 sums: general path
    # heavily skewed keys: every lane of a warp on the same slot
    ks = [np.where(rng.random(n) < 0.9, 7, rng.integers(0, 300, n)).astype(np.int64)]
    nn, gk, ga = _mid(K, [ks], [v], [(0, "sum"), (0, "count")])
    ek, ea = pandas_agg(ks, v, [(0, "sum"), (0, "count")])
    order = np.argsort(gk[0])
    np.testing.assert_array_equal(gk[0][order], ek[0])
    np.testing.assert_array_equal(ga[1][order], ea[1])
    np.testing.assert_allclose(ga[0][order], ea[0], rtol=1e-9)


def test_estimate_distinct_tells_the_regimes_apart(K):
    rng = np.random.default_rng(2)
    n = 1_000_000
    for card, lo, hi in [(4, 4, 4), (1000, 900, 1000), (10 ** 8, 15_000, 16_384)]:
        s, d = K.estimate_distinct([dev(rng.integers(0, card, n, dtype=np.int64))])
        assert s == 16384 and lo <= d <= hi, (card, s, d)
    s, d = K.estimate_distinct([dev(rng.integers(0, 30, 5000, dtype=np.int64)), dev(rng.integers(0, 30, 5000, dtype=np.int64))])
    assert s == 5000 and 850 <= d <= 900


@pytest.mark.parametrize("n_keys", [1, 2])
@pytest.mark.parametrize("n,R", [(0, 4), (1, 1), (5000, 7), (300_000, 64), (1_000_003, 256), (200_000, 1000)])
def test_hash_partition_equals_ids_plus_stable_split(K, n_keys, n, R):
    """a5 in one call: same blocks, same order inside every block (ascending input positions f_i of the
    reference, groupby/core.py:389-435) as mgb_partition_ids + mgb_partition_scatter, and as the oracle"""
    rng = np.random.default_rng(n + R)
    keys = [rng.integers(-10 ** 9, 10 ** 9, n, dtype=np.int64) for _ in range(n_keys)]
    cols = keys + [rng.random(n), np.arange(n, dtype=np.int64)]
    outs, offsets = K.hash_partition([dev(k) for k in keys], R, [dev(c) for c in cols])
    pid = orc.partition_ids(keys, R)
    order = np.argsort(pid, kind="stable")
    exp_off = np.searchsorted(pid[order], np.arange(R + 1)).tolist()
    assert offsets == exp_off
    for got, src in zip(outs, cols):
        np.testing.assert_array_equal(got.cpu().numpy(), src[order])


@pytest.mark.parametrize("n_keys,mode", [(1, 0), (2, 0), (1, 1), (2, 1)])
@pytest.mark.parametrize("n,R", [(1, 3), (2047, 5), (2049, 8), (123_457, 8), (500_001, 256), (40_000, 1)])
def test_split_two_passes_with_destination_table(K, n_keys, mode, n, R):
    """the partition kernel in its two passes (mgb_split_plan / mgb_split_scatter): block sizes first, then every row
    written once to an address taken from a destination table -- here blocks laid out back to front with gaps, in
    one buffer per column; column views at odd 8-byte offsets (no TMA, no 16-byte stores) give the same bytes"""
    rng = np.random.default_rng(n * 7 + R + mode)
    keys = [rng.integers(-10 ** 6, 10 ** 6, n, dtype=np.int64) for _ in range(n_keys)]
    cols = keys + [rng.random(n), np.arange(n, dtype=np.int64), rng.normal(size=n)]
    pid = ((orc.hash_frame_rows(keys) % np.uint64(R)).astype(np.int64) if mode == 1 else orc.partition_ids(keys, R))
    order = np.argsort(pid, kind="stable")
    counts = np.bincount(pid, minlength=R)
    dk = [dev(k) for k in keys]
    # odd element offset for every second column: source not 16-byte aligned
    dcols = []
    for i, c in enumerate(cols):
        if i % 2:
            buf = torch.empty(n + 1, dtype=torch.from_numpy(c).dtype, device="cuda")
            buf[1:] = dev(c)
            dcols.append(buf[1:])
        else:
            dcols.append(dev(c))
    dcols[:n_keys] = dk
    hist, offsets = K.split_plan(dk, R, mode)
    off = offsets.cpu().numpy()
    np.testing.assert_array_equal(off, np.concatenate([[0], np.cumsum(counts)]))
    # local mode
    outs = [torch.full_like(c, -1) for c in dcols]
    K.split_scatter(dk, R, hist, dcols, outs=outs, mode=mode)
    for got, src in zip(outs, cols):
        np.testing.assert_array_equal(got.cpu().numpy(), src[order])
    # table mode: block p of column c at element start[p] (+1 for odd columns: unaligned destinations) of arena c
    gap = 3
    start = np.zeros(R, dtype=np.int64)
    pos = 0
    for p in reversed(range(R)):
        start[p] = pos
        pos += (counts[p] + gap + 1) & ~1
    arenas = [torch.full((pos + 2,), -7, dtype=torch.int64, device="cuda") for _ in cols]
    table = np.zeros((len(cols), R), dtype=np.int64)
    for c in range(len(cols)):
        table[c] = arenas[c].data_ptr() + 8 * (start + (c % 2))
    K.split_scatter(dk, R, hist, dcols, table=dev(table.reshape(-1)), mode=mode)
    torch.cuda.synchronize()
    for c, src in enumerate(cols):
        a = arenas[c].cpu().numpy()
        sorted_src = src[order]
        touched = np.zeros(len(a), dtype=bool)
        for p in range(R):
            s0 = start[p] + (c % 2)
            blk = a[s0:s0 + counts[p]]
            np.testing.assert_array_equal(blk.view(src.dtype), sorted_src[off[p]:off[p + 1]])
            touched[s0:s0 + counts[p]] = True
        assert (a[~touched] == -7).all(), "a store outside the blocks"


@pytest.mark.parametrize("n,card,block", [(200_000, 300, 0), (1_000_003, 50_000, 65_536), (0, 1, 0)])
def test_host_buffer_entry_matches_pandas(K, n, card, block):
    """mgb_groupby_agg_host: the whole path from HOST columns (blocked H2D overlapped with the kernels, result into
    host arrays) -- what a binding without device arrays calls; every op, both dtypes, sorted keys"""
    import ctypes as C

    from mars_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(n + card)
    k = rng.integers(-card, card, n, dtype=np.int64)
    v = rng.normal(5, 2, n)
    if n:
        v[rng.integers(0, n, max(n // 50, 1))] = np.nan
    w = rng.integers(-100, 100, n, dtype=np.int64)
    accs = [(0, K.SUM), (0, K.COUNT), (0, K.MIN), (0, K.MAX), (0, K.SUMSQ), (1, K.SUM), (1, K.MIN), (1, K.MAX), (1, K.COUNT)]
    spec = _lib.make_spec(1, [_lib.F64, _lib.I64], accs)
    cap = max(2 * card + 8, 8)
    out_k = np.zeros(cap, dtype=np.int64)
    out_a = [np.zeros(cap, dtype=np.float64 if (c == 0 and op != K.COUNT) else np.int64) for c, op in accs]
    keys_p = _lib.ptr_array([k.ctypes.data])
    vals_p = _lib.ptr_array([v.ctypes.data, w.ctypes.data])
    ok_p = _lib.ptr_array([out_k.ctypes.data])
    oa_p = _lib.ptr_array([a.ctypes.data for a in out_a])
    G = C.c_int64()
    _lib.check(lib.mgb_groupby_agg_host(C.byref(spec), keys_p, vals_p, n, block, ok_p, oa_p, cap, 1, C.byref(G),
                                         torch.cuda.current_stream().cuda_stream))
    df = pd.DataFrame({"k": k, "v": v, "w": w})
    g = df.groupby("k", sort=True)
    assert G.value == g.ngroups
    m = G.value
    if m == 0:
        return
    np.testing.assert_array_equal(out_k[:m], np.array(sorted(g.groups.keys()), dtype=np.int64))
    exp = g.agg(v_sum=("v", "sum"), v_count=("v", "count"), v_min=("v", "min"), v_max=("v", "max"),
                w_sum=("w", "sum"), w_min=("w", "min"), w_max=("w", "max"), w_count=("w", "count"))
    sq = (df["v"] ** 2).groupby(df["k"], sort=True).sum()
    np.testing.assert_allclose(out_a[0][:m], exp["v_sum"].to_numpy(), rtol=1e-9, atol=0)
    np.testing.assert_array_equal(out_a[1][:m], exp["v_count"].to_numpy())
    np.testing.assert_array_equal(out_a[2][:m], exp["v_min"].to_numpy())
    np.testing.assert_array_equal(out_a[3][:m], exp["v_max"].to_numpy())
    np.testing.assert_allclose(out_a[4][:m], sq.to_numpy(), rtol=1e-9, atol=0)
    np.testing.assert_array_equal(out_a[5][:m], exp["w_sum"].to_numpy())
    np.testing.assert_array_equal(out_a[6][:m], exp["w_min"].to_numpy())
    np.testing.assert_array_equal(out_a[7][:m], exp["w_max"].to_numpy())
    np.testing.assert_array_equal(out_a[8][:m], exp["w_count"].to_numpy())


@pytest.mark.parametrize("n_keys,ncol", [(1, 1), (2, 5), (1, 6), (2, 3)])
@pytest.mark.parametrize("layout", ["mixed4", "sorted12", "sorted40", "five"])
def test_register_accumulator_kernel_and_its_hand_over(K, n_keys, ncol, layout):
    """sum / count shapes go first to the kernel that keeps <= 4 groups per CTA in registers (TMA-staged tiles,
    mgb_tiny.cu); whatever it cannot take -- a fifth key tuple in one CTA, more than 16 tuples in total -- is done by
    the dense kernel enqueued behind it, in the same call.  mixed4: 4 tuples everywhere; sorted12 / sorted40: every CTA
    sees few tuples but the band has 12 / 40; five: a fifth tuple appears in the last rows of the last chunk"""
    rng = np.random.default_rng(n_keys * 10 + ncol)
    sizes = [300_000, 1, 70_001, 1024, 2048 + 17]
    total = sum(sizes)
    if layout == "mixed4":
        key = rng.integers(0, 4, total, dtype=np.int64)
    elif layout == "five":
        key = rng.integers(0, 4, total, dtype=np.int64)
        key[-3:] = 9
    else:
        groups = 12 if layout == "sorted12" else 40
        key = np.sort(rng.integers(0, groups, total, dtype=np.int64)) * 1_000_003 - 7
    keys_all = [key] + ([(key % 2) * (2 ** 40)] if n_keys == 2 else [])
    vals_all = [rng.normal(100, 30, total) for _ in range(ncol)] + [rng.integers(1, 99, total, dtype=np.int64)]
    vals_all[0][rng.random(total) < 0.05] = np.nan
    ck, cv, pos = [], [], 0
    for n in sizes:
        ck.append([k[pos:pos + n] for k in keys_all])
        cv.append([v[pos:pos + n] for v in vals_all])
        pos += n
    accs = [(c, "sum") for c in range(ncol)] + [(0, "count")] + ([(ncol - 1, "count")] if ncol > 1 else []) + [(ncol, "count")]
    n, gk, ga = _batch(K, ck, cv, accs, skip_cols=(ncol,))
    ek, ea = pandas_agg(keys_all, vals_all, accs)
    assert n == len(ek[0])
    order = np.lexsort(tuple(reversed(gk)))
    for a, b in zip(gk, ek):
        np.testing.assert_array_equal(a[order], b)
    for (col, op), a, b in zip(accs, ga, ea):
        if op == "count":
            np.testing.assert_array_equal(a[order], b)
        else:
            np.testing.assert_allclose(a[order], b, rtol=1e-9, atol=0)


@pytest.mark.parametrize("n,R,nk", [(300_001, 8, 1), (1_000_003, 64, 2), (262_144, 3, 1), (100_000, 64, 1)])
def test_split_with_long_runs_gives_the_same_blocks(K, n, R, nk):
    """mode | SPLIT_LONG_RUNS (8192-row tiles, what the push shuffle asks for): same blocks, same order inside every
    block as the default geometry; chunks below 2^18 rows keep the small tiles"""
    rng = np.random.default_rng(n + R)
    keys = [rng.integers(-10 ** 9, 10 ** 9, n, dtype=np.int64) for _ in range(nk)]
    cols = keys + [rng.random(n), np.arange(n, dtype=np.int64)]
    dk, dc = [dev(k) for k in keys], [dev(c) for c in cols]
    pid = orc.partition_ids(keys, R)
    order = np.argsort(pid, kind="stable")
    hist, offsets = K.split_plan(dk, R, K.SPLIT_LONG_RUNS)
    np.testing.assert_array_equal(offsets.cpu().numpy(), np.searchsorted(pid[order], np.arange(R + 1)))
    outs = [torch.full_like(c, -1) for c in dc]
    K.split_scatter(dk, R, hist, dc, outs=outs, mode=K.SPLIT_LONG_RUNS)
    for got, src in zip(outs, cols):
        np.testing.assert_array_equal(got.cpu().numpy(), src[order])
