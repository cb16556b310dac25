"""TEST DOUBLE for mars_b200.kernels (CPU torch tensors, pandas arithmetic through the oracle's call sites).

Used only by the `-m "not gpu"` host-logic tests to exercise tiling / executors / session / exchange plumbing
without a GPU.  It is not importable from the package and is never a fallback: mars_b200.kernels raises
MarsGBError for CPU tensors."""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

import mars_groupby_oracle as orc

F64, I64 = 0, 1
SUM, SUMSQ, COUNT, MIN, MAX = 0, 1, 2, 3, 4
launch_counter = 0
calls = []


def dtype_code(t):
    return F64 if t.dtype == torch.float64 else I64


def _np(t):
    return t.cpu().numpy()


def hash_keys(keys):
    calls.append("hash_keys")
    return torch.from_numpy(orc.hash_keys([_np(k) for k in keys]).view(np.int64))


def partition_ids(keys, n_partitions):
    calls.append("partition_ids")
    return torch.from_numpy(orc.partition_ids([_np(k) for k in keys], n_partitions).astype(np.int32))


def range_partition_ids(keys, pivots):
    calls.append("range_partition_ids")
    piv = list(zip(*[_np(p).tolist() for p in pivots])) if len(pivots) > 1 else _np(pivots[0])
    pid = orc.range_partition_ids([_np(k) for k in keys], piv)
    return torch.from_numpy(np.asarray(pid).astype(np.int32))


def partition_scatter(pid, n_partitions, cols):
    calls.append("partition_scatter")
    p = _np(pid)
    order = np.argsort(p, kind="stable")
    offsets = np.searchsorted(p[order], np.arange(n_partitions + 1)).tolist()
    idx = torch.from_numpy(order)
    return [c[idx] for c in cols], offsets


def groupby_aggregate(keys, vals, accs, sort=False, expect_distinct=False):
    calls.append("groupby_aggregate")
    df = pd.DataFrame({f"k{j}": _np(k) for j, k in enumerate(keys)})
    names = [f"k{j}" for j in range(len(keys))]
    outs = []
    for i, (col, op) in enumerate(accs):
        v = _np(vals[col])
        if op == SUMSQ:
            v = v * v
        df["_v"] = v
        g = df.groupby(names, sort=sort)["_v"]
        r = {SUM: g.sum, SUMSQ: g.sum, COUNT: g.count, MIN: g.min, MAX: g.max}[op]()
        outs.append(r)
    idx = outs[0].index if outs else df.groupby(names, sort=sort).size().index
    karr = [idx.get_level_values(j).to_numpy() for j in range(len(keys))]
    return ([torch.from_numpy(np.ascontiguousarray(k)) for k in karr],
            [torch.from_numpy(np.ascontiguousarray(r.to_numpy())) for r in outs],
            dict(bucket_fallback=0, rows_smem=0, rows_spilled=0, partial_rows=0, table_capacity=0, launches=0))


class Aggregator:
    """double of kernels.Aggregator: collects row batches, aggregates them on result()"""

    def __init__(self, n_keys, val_dtypes, accs, expect_distinct=False):
        self.n_keys, self.accs = n_keys, list(accs)
        self._k = [[] for _ in range(n_keys)]
        self._v = [[] for _ in val_dtypes]
        self.stats = None

    def update(self, keys, vals):
        for j, k in enumerate(keys):
            self._k[j].append(k)
        for j, v in enumerate(vals):
            self._v[j].append(v)
        return self

    def result(self, sort=False, device=None):
        keys = [torch.cat(k) if k else torch.empty(0, dtype=torch.int64) for k in self._k]
        vals = [torch.cat(v) for v in self._v]
        k, a, self.stats = groupby_aggregate(keys, vals, self.accs, sort=sort)
        return k, a

    def close(self):
        pass


def post_mean(s, c):
    calls.append("post_mean")
    return torch.from_numpy(_np(s) / _np(c))


def post_var(s, q, c, ddof=1):
    calls.append("post_var")
    return torch.from_numpy(orc.post_var(pd.Series(_np(q)), pd.Series(_np(s)), pd.Series(_np(c)), ddof=ddof).to_numpy())


def concat_columns(pieces):
    calls.append("concat_columns")
    return pieces[0] if len(pieces) == 1 else torch.cat(list(pieces))


def argsort_keys(keys):
    return torch.from_numpy(np.lexsort(tuple(reversed([_np(k) for k in keys]))).astype(np.int32))


def gather_rows(cols, perm):
    idx = perm.to(torch.int64)
    return [c[idx] for c in cols]
