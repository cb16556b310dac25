"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's groupby-aggregation hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this module, and only as the checker or as the timed CPU baseline.  Nothing under ``mars_b200/``
imports it; the product path fails loudly when libmarsgb.so is missing.

What is restated (paths relative to the reference checkout, mars-project/mars):

  hash_keys / partition_ids     pd.util.hash_pandas_object(idx, categorize=False) % size as called by
                                hash_dataframe_on (mars/dataframe/utils.py:77-101).  The arithmetic lives in
                                the third-party dependency pandas (pandas/core/util/hashing.py:
                                _hash_ndarray tail, hash_tuples, combine_hash_arrays; reference pin
                                pandas>=1.0,<2.0, setup.cfg:30; installed here 3.0.2 -- the int64 algorithm
                                is unchanged over that range).  Restated in numpy uint64 arithmetic.
  compile_steps                 ReductionCompiler output for sum/count/min/max/mean/var
                                (mars/dataframe/reduction/core.py:908-1250; mean.py:26-33; var.py:38-48)
  agg_map                       DataFrameGroupByAgg._execute_map      (groupby/aggregation.py:1043-1128)
  shuffle_map                   DataFrameGroupByOperand.execute_map   (groupby/core.py:353-435)
  shuffle_reduce                DataFrameGroupByOperand.execute_reduce (groupby/core.py:437-485)
  concat_partials               DataFrameConcat.execute tuple branch  (merge/concat.py:339-348)
  agg_combine                   DataFrameGroupByAgg._execute_combine  (groupby/aggregation.py:1130-1164)
  agg_agg                       DataFrameGroupByAgg._execute_agg      (groupby/aggregation.py:1166-1267)
  psrs_sample / concat_pivot /  DataFramePSRSGroupbySample, DataFrameGroupbyConcatPivot, DataFrameGroupbySortShuffle
  sort_shuffle_map / _reduce    (groupby/sort.py:36-167): the range shuffle taken when sort=True
  run_groupby_agg               tile rules tree / shuffle (aggregation.py:428-471, 563-702, 715-835) + the above

Parity pinning: tests/test_oracle.py checks this module against (i) the known-answer hash vectors of
pandas and (ii) tests/golden/*.npz -- outputs of the *unmodified reference* executed
in the build container by oracle/make_golden.py (every stage: partials, shuffle blocks, reducer index per
key, final frames).  The per-group arithmetic itself (Kahan group_sum, group_min/max, count) is pandas',
called here at the same sites the reference calls it.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

# ----------------------------------------------------------------------------------------------------
# hashing (numpy restatement, bit-exact)
# ----------------------------------------------------------------------------------------------------
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def mix64(x: np.ndarray) -> np.ndarray:
    """pandas `_hash_ndarray` tail for 8-byte numeric input (splitmix64 finaliser)."""
    x = np.ascontiguousarray(x).view(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= _M1
        x ^= x >> np.uint64(27)
        x *= _M2
        x ^= x >> np.uint64(31)
    return x


def hash_keys(key_cols: Sequence[np.ndarray]) -> np.ndarray:
    """uint64 hash per key tuple: one key -> mix64; several -> hash_tuples / combine_hash_arrays."""
    cols = [np.asarray(c, dtype=np.int64) for c in key_cols]
    if len(cols) == 1:
        return mix64(cols[0])
    n_levels = len(cols)
    out = np.full(len(cols[0]), 0x345678, dtype=np.uint64)
    mult = np.uint64(1000003)
    with np.errstate(over="ignore"):
        for i, c in enumerate(cols):
            inverse_i = n_levels - i
            out ^= mix64(c)
            out *= mult
            mult = mult + np.uint64(82520 + inverse_i + inverse_i)
        out += np.uint64(97531)
    return out


def partition_ids(key_cols: Sequence[np.ndarray], n_partitions: int) -> np.ndarray:
    """reducer index of every key tuple: hash % size as uint64 (mars/dataframe/utils.py:100)."""
    return (hash_keys(key_cols) % np.uint64(n_partitions)).astype(np.int64)


def hash_frame_rows(cols: Sequence[np.ndarray]) -> np.ndarray:
    """uint64 hash of every ROW of `df[on]` -- pd.util.hash_pandas_object(df[on], index=False, categorize=False) as
    called by hash_dataframe_on(df, on=[...]) (mars/dataframe/utils.py:88-99): the per-column hashes are always
    combined (combine_hash_arrays), also for a single column."""
    cols = [np.asarray(c, dtype=np.int64) for c in cols]
    if len(cols) > 1:
        return hash_keys(cols)
    with np.errstate(over="ignore"):
        return (np.uint64(0x345678) ^ mix64(cols[0])) * np.uint64(1000003) + np.uint64(97531)


def hash_dataframe_on_columns(df: pd.DataFrame, on: Sequence[Any], size: int) -> List[np.ndarray]:
    """hash_dataframe_on(df, on=[...], size) restated: ascending row positions per reducer"""
    pid = (hash_frame_rows([df[c].to_numpy() for c in on]) % np.uint64(size)).astype(np.int64)
    order = np.argsort(pid, kind="stable")
    bounds = np.searchsorted(pid[order], np.arange(size + 1))
    return [order[bounds[i]:bounds[i + 1]] for i in range(size)]


def hash_dataframe_on_index(index: pd.Index, size: int) -> List[np.ndarray]:
    """hash_dataframe_on(df, on=None, size) restated: ascending row positions per reducer."""
    if isinstance(index, pd.MultiIndex):
        cols = [index.get_level_values(i).to_numpy() for i in range(index.nlevels)]
    else:
        cols = [index.to_numpy()]
    pid = partition_ids(cols, size)
    order = np.argsort(pid, kind="stable")
    bounds = np.searchsorted(pid[order], np.arange(size + 1))
    return [order[bounds[i]:bounds[i + 1]] for i in range(size)]


# ----------------------------------------------------------------------------------------------------
# function decomposition
# ----------------------------------------------------------------------------------------------------
SUPPORTED = ("sum", "count", "min", "max", "mean", "var")
# func -> atomic steps (pre, map func, agg func) in the compiler's registration order
_DECOMP = {
    "sum": [("identity", "sum", "sum")],
    "count": [("identity", "count", "sum")],
    "min": [("identity", "min", "min")],
    "max": [("identity", "max", "max")],
    "mean": [("identity", "sum", "sum"), ("identity", "count", "sum")],
    "var": [("identity", "count", "sum"), ("square", "sum", "sum"), ("identity", "sum", "sum")],
}


class Step:
    def __init__(self, pre, map_func, agg_func):
        self.pre, self.map_func, self.agg_func = pre, map_func, agg_func
        self.key = f"{map_func}_{pre}"

    def __repr__(self):
        return f"Step({self.key}->{self.agg_func})"


class Post:
    def __init__(self, func, inputs, cols):
        self.func, self.inputs, self.cols = func, inputs, cols


def normalize_func(func) -> "OrderedDict | list":
    if isinstance(func, dict):
        return OrderedDict((c, [f] if isinstance(f, str) else list(f)) for c, f in func.items())
    if isinstance(func, str):
        return [func]
    return list(func)


def compile_steps(func, ddof: int = 1) -> Tuple[List[Step], List[Post]]:
    """Atomic steps (deduplicated) + post steps in user order."""
    func = normalize_func(func)
    items = [(None, f) for f in func] if isinstance(func, list) else [(c, f) for c, fs in func.items() for f in fs]
    steps: "OrderedDict[str, Step]" = OrderedDict()
    posts: "OrderedDict[str, Post]" = OrderedDict()
    for col, f in items:
        if f not in SUPPORTED:
            raise NotImplementedError(f)
        keys = []
        for pre, m, a in _DECOMP[f]:
            s = Step(pre, m, a)
            steps.setdefault(s.key, s)
            keys.append(s.key)
        if f == "var":
            keys = [keys[1], keys[2], keys[0]]  # sum(x**2), sum(x), count
        p = posts.setdefault(f, Post(f, keys, None if col is None else []))
        if col is not None and p.cols is not None:
            p.cols.append(col)
    return list(steps.values()), list(posts.values())


def _columns_of_step(step: Step, func, value_cols: Sequence[Any]) -> List[Any]:
    """columns on which a step is consumed by some user function"""
    func = normalize_func(func)
    if isinstance(func, list):
        per_col = {c: func for c in value_cols}
    else:
        per_col = func
    out = []
    for c in value_cols:
        for f in per_col.get(c, []):
            if any(pre == step.pre and m == step.map_func for pre, m, _ in _DECOMP[f]):
                out.append(c)
                break
    return out


# ----------------------------------------------------------------------------------------------------
# stages
# ----------------------------------------------------------------------------------------------------
def agg_map(chunk: pd.DataFrame, by: List[Any], func, selection=None, sort: bool = False) -> Tuple[pd.DataFrame, ...]:
    """_execute_map: tuple of K partial frames indexed by the group key(s).  As in the reference the map
    stage always runs with sort as given and as_index=True; `sort` does not matter for parity because the
    merge re-groups.  Columns no user function consumes are not materialised (the reference computes them
    and drops them in the post step's column selection, aggregation.py:1214-1217)."""
    steps, _ = compile_steps(func)
    value_cols = [c for c in chunk.columns if c not in by and (selection is None or c in selection)]
    if isinstance(normalize_func(func), OrderedDict):
        value_cols = [c for c in value_cols if c in normalize_func(func)]
    out = []
    for s in steps:
        cols = _columns_of_step(s, func, value_cols)
        data = chunk[by + cols]
        if s.pre == "square":
            data = data.copy()
            data[cols] = data[cols].pow(2)
        g = data.groupby(by, sort=sort)[cols]
        res = g.agg([s.map_func])
        res.columns = res.columns.droplevel(-1)
        out.append(res)
    return tuple(out)


def shuffle_map(partials: Tuple[pd.DataFrame, ...], shuffle_size: int, mapper_index) -> Dict[int, Tuple]:
    """execute_map: reducer i -> (mapper_index, (df_0.iloc[f_i], ..., False)); empty blocks included."""
    filters = hash_dataframe_on_index(partials[0].index, shuffle_size)
    out = {}
    for i in range(shuffle_size):
        out[i] = (mapper_index, tuple(p.iloc[filters[i]] for p in partials) + (False,))
    return out


def shuffle_reduce(blocks: Sequence[Tuple]) -> Tuple[pd.DataFrame, ...]:
    """execute_reduce: blocks of one reducer, concatenated slot-wise in ascending mapper index."""
    blocks = sorted(blocks, key=lambda b: b[0])
    payloads = [b[1] for b in blocks]
    n = len(payloads[0]) - 1
    return tuple(pd.concat([p[k] for p in payloads], axis=0) for k in range(n))


def concat_partials(items: Sequence[Tuple[pd.DataFrame, ...]]) -> Tuple[pd.DataFrame, ...]:
    return tuple(pd.concat([it[k] for it in items], axis=0) for k in range(len(items[0])))


def _regroup(df: pd.DataFrame, sort: bool):
    return df.groupby(level=list(range(df.index.nlevels)), sort=sort)


def agg_combine(partials: Tuple[pd.DataFrame, ...], func) -> Tuple[pd.DataFrame, ...]:
    steps, _ = compile_steps(func)
    out = []
    for s, p in zip(steps, partials):
        res = _regroup(p, sort=False).agg([s.agg_func])
        res.columns = res.columns.droplevel(-1)
        out.append(res)
    return tuple(out)


def post_mean(s, c):
    return s / c


def post_var(q, s, c, ddof=1):
    """operation order of the generated source (reduction/var.py:38-48)"""
    if ddof == 0:
        return q / c - (s / c) ** 2
    t = s.pow(2) if hasattr(s, "pow") else s ** 2
    t = t / c
    u = q - t
    d = c - ddof
    return u / d


def agg_agg(partials: Tuple[pd.DataFrame, ...], func, sort: bool, ddof: int = 1) -> pd.DataFrame:
    steps, posts = compile_steps(func, ddof)
    list_form = isinstance(normalize_func(func), list)
    single = isinstance(func, str)
    merged = {}
    for s, p in zip(steps, partials):
        res = _regroup(p, sort=sort).agg([s.agg_func])
        res.columns = res.columns.droplevel(-1)
        merged[s.key] = res
    pieces = []
    for p in posts:
        ins = [merged[k] for k in p.inputs]
        if p.cols is not None:
            ins = [x[[c for c in p.cols if c in x.columns]] for x in ins]
        common = [c for c in ins[0].columns if all(c in x.columns for x in ins[1:])]
        ins = [x[common] for x in ins]
        if p.func == "mean":
            r = post_mean(*ins)
        elif p.func == "var":
            r = post_var(*ins, ddof=ddof)
        else:
            r = ins[0]
        if not single:
            r = r.copy()
            r.columns = pd.MultiIndex.from_product([r.columns, [p.func]])
        pieces.append(r)
    result = pd.concat(pieces, axis=1)
    # declared column order: value-column major in frame order, function order as given by the user
    if not single:
        if list_form:
            order = [(c, f) for c in _ordered_unique([c for c, _ in result.columns]) for f in normalize_func(func)]
        else:
            order = [(c, f) for c, fs in normalize_func(func).items() for f in fs]
        result = result.reindex(pd.MultiIndex.from_tuples(order), axis=1)
    return result


def _ordered_unique(xs):
    seen, out = set(), []
    for x in xs:
        if x not in seen:
            seen.add(x)
            out.append(x)
    return out


# ----------------------------------------------------------------------------------------------------
# sort=True shuffle: parallel sorting by regular sampling (PSRS) over the partial frames
# ----------------------------------------------------------------------------------------------------
def psrs_sample(first_partial: pd.DataFrame, n_partition: int) -> pd.DataFrame:
    """DataFramePSRSGroupbySample.execute (sort.py:70-97): n regular samples of the key-sorted first frame"""
    a = first_partial
    n = n_partition
    if a.shape[0] < n:
        num = n // a.shape[0] + 1
        a = pd.concat([a] * num).sort_index()
    w = a.shape[0] * 1.0 / (n + 1)
    slc = np.linspace(max(w - 1, 0), a.shape[0] - 1, num=n, endpoint=False).astype(int)
    return a.iloc[slc]


def concat_pivot(samples: Sequence[pd.DataFrame]) -> np.ndarray:
    """DataFrameGroupbyConcatPivot.execute (sort.py:44-67): p - 1 pivots out of the gathered samples"""
    n_inputs = len(samples)
    inputs = [x for x in samples if len(x) > 0]
    a = pd.concat(inputs, axis=0).sort_index()
    index = a.index.drop_duplicates()
    p = len(inputs)
    if len(index) < p:
        num = p // len(index) + 1
        index = index.append([index] * (num - 1))
    index = index.sort_values()
    values = index.values
    slc = np.linspace(p - 1, len(index) - 1, num=n_inputs - 1, endpoint=False).astype(int)
    return values[slc]


def sort_shuffle_map(partials: Tuple[pd.DataFrame, ...], pivots: np.ndarray, n_partition: int) -> Dict[int, Tuple]:
    """DataFrameGroupbySortShuffle._execute_map (sort.py:113-135): partition i holds the keys in
    (pivot[i-1], pivot[i]] -- label slices of the key-sorted frames, lower pivot dropped"""
    def part(i, df):
        if i == 0:
            return df.loc[: pivots[i]]
        if i == n_partition - 1:
            return df.loc[pivots[i - 1]:].drop(index=pivots[i - 1], errors="ignore")
        return df.loc[pivots[i - 1]: pivots[i]].drop(index=pivots[i - 1], errors="ignore")

    return {i: tuple(part(i, x) for x in partials) for i in range(n_partition)}


def sort_shuffle_reduce(blocks: Sequence[Tuple]) -> Tuple[pd.DataFrame, ...]:
    """DataFrameGroupbySortShuffle._execute_reduce (sort.py:137-151), without the trailing `by`"""
    return tuple(pd.concat([b[k] for b in blocks], axis=0) for k in range(len(blocks[0])))


def range_partition_ids(key_cols: Sequence[np.ndarray], pivots: np.ndarray) -> np.ndarray:
    """partition of every key tuple under the rule above: number of pivots strictly smaller than the key"""
    if len(key_cols) == 1:
        return np.searchsorted(np.asarray(pivots, dtype=np.int64), np.asarray(key_cols[0]), side="left")
    piv = [tuple(int(v) for v in t) for t in pivots]
    keys = list(zip(*[np.asarray(c).tolist() for c in key_cols]))
    import bisect

    return np.array([bisect.bisect_left(piv, k) for k in keys], dtype=np.int64)


# ----------------------------------------------------------------------------------------------------
# whole path
# ----------------------------------------------------------------------------------------------------
def run_groupby_agg(raw: pd.DataFrame, by, func, chunk_size: int, method: str = "tree", sort: bool = True,
                    combine_size: int = 4, selection=None, record: Optional[dict] = None) -> pd.DataFrame:
    """md.DataFrame(raw, chunk_size).groupby(by, sort=sort)[selection].agg(func, method=method) executed stage
    by stage.  method='shuffle': hash partition for sort=False, PSRS range partition for sort=True."""
    by = [by] if not isinstance(by, (list, tuple)) else list(by)
    chunks = [raw.iloc[i:i + chunk_size] for i in range(0, max(len(raw), 1), chunk_size)]
    maps = [agg_map(c, by, func, selection, sort=sort) for c in chunks]
    if record is not None:
        record["map"] = maps
    if method == "shuffle" and sort and len(chunks) > 1:
        R = len(maps)
        samples = [psrs_sample(m[0], R) for m in maps]
        pivots = concat_pivot(samples)
        blocks = [sort_shuffle_map(m, pivots, R) for m in maps]
        if record is not None:
            record["pivots"] = pivots
            record["blocks"] = blocks
        outs = [agg_agg(sort_shuffle_reduce([blocks[i][r] for i in range(R)]), func, sort=True) for r in range(R)]
        if record is not None:
            record["chunk_results"] = outs
        return pd.concat(outs, axis=0)
    if method == "shuffle":
        R = len(maps)
        blocks = [shuffle_map(m, R, (i, 0)) for i, m in enumerate(maps)]
        if record is not None:
            record["blocks"] = blocks
        outs = []
        for r in range(R):
            red = shuffle_reduce([blocks[i][r] for i in range(R)])
            outs.append(agg_agg(red, func, sort=sort))
        if record is not None:
            record["chunk_results"] = outs
        return pd.concat(outs, axis=0) if len(outs) > 1 else outs[0]
    if method != "tree":
        raise NotImplementedError(method)
    level = maps
    while len(level) > combine_size:
        nxt = []
        for i in range(0, len(level), combine_size):
            grp = level[i:i + combine_size]
            cat = grp[0] if len(grp) == 1 else concat_partials(grp)
            nxt.append(agg_combine(cat, func))
        level = nxt
    cat = level[0] if len(level) == 1 else concat_partials(level)
    out = agg_agg(cat, func, sort=sort)
    if record is not None:
        record["chunk_results"] = [out]
    return out
