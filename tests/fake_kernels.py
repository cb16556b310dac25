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


# ----------------------------------------------------------------------------------------------------
# doubles of the asynchronous / pointer-level entry points: the "device" is host memory, so a raw pointer is
# read back as a numpy view.  Layouts (block pair, count row, result block) are the real ones of
# mars_b200.kernels, which the executors rely on.
# ----------------------------------------------------------------------------------------------------
import ctypes as _C

from mars_b200 import kernels as _K

POST_IDENTITY, POST_MEAN, POST_VAR, MAX_POST = _K.POST_IDENTITY, _K.POST_MEAN, _K.POST_VAR, _K.MAX_POST
PostPlan = _K.PostPlan
SMALL_MAX_ROWS, SMALL_MAX_PIECES = _K.SMALL_MAX_ROWS, _K.SMALL_MAX_PIECES
DENSE_GROUPS = 16          # a batch with more groups "is not dense" (device count -1), like a CTA overflowing
_DENSE_CAP = 148 * 16


def _view(ptr, n, is_float):
    if n == 0:
        return np.empty(0, dtype=np.float64 if is_float else np.int64)
    buf = (_C.c_int64 * n).from_address(int(ptr))
    a = np.frombuffer(buf, dtype=np.int64)
    return a.view(np.float64) if is_float else a


def dense_batch_max():
    return 64


def _agg_np(keys, vals, accs, is_float_val, sort=False):
    """(key arrays, accumulator arrays) with pandas (first-appearance order unless sort)"""
    names = [f"k{j}" for j in range(len(keys))]
    df = pd.DataFrame({n: k for n, k in zip(names, keys)})
    outs = []
    for col, op in accs:
        v = vals[col]
        df["_v"] = v * v if op == SUMSQ else v
        g = df.groupby(names, sort=sort)["_v"]
        outs.append({SUM: g.sum, SUMSQ: g.sum, COUNT: g.count, MIN: g.min, MAX: g.max}[op]().to_numpy())
    idx = df.groupby(names, sort=sort).size().index
    return [idx.get_level_values(j).to_numpy() for j in range(len(keys))], outs


def _fill_block(fc, bi, bf, keys, accs_out, n):
    k_views, a_views = fc.views(bi, bf)
    if n >= 0:
        for dst, src in zip(k_views, keys):
            dst[:n] = torch.from_numpy(np.ascontiguousarray(src))
        for dst, src in zip(a_views, accs_out):
            dst[:n] = torch.from_numpy(np.ascontiguousarray(src).astype(np.float64 if dst.dtype == torch.float64 else np.int64))
    bi[fc.n_int, 0] = n


def preagg_dense_batch_async(chunks, accs, val_dtypes, n_keys, device):
    calls.append("preagg_dense_batch_async")
    fc = _K._fast(n_keys, val_dtypes, accs)
    bi, bf = fc.alloc(_DENSE_CAP, device)
    used = {c for c, _ in accs}
    keys = [np.concatenate([_view(kp[j], m, False) for kp, vp, m in chunks]) for j in range(n_keys)]
    vals = []
    for c, dt in enumerate(val_dtypes):
        if c in used and all(vp[c] is not None for _, vp, _ in chunks):
            vals.append(np.concatenate([_view(vp[c], m, dt == F64) for _, vp, m in chunks]))
        else:   # count-only int64 column that was not uploaded: its values do not matter
            vals.append(np.zeros(len(keys[0]), dtype=np.int64))
    k, a = _agg_np(keys, vals, accs, None)
    n = len(k[0]) if len(k[0]) <= DENSE_GROUPS else -1
    _fill_block(fc, bi, bf, k, a, n)
    return bi, bf, fc


def preagg_dense_async(keys, vals, accs, checked=False, val_dtypes=None):
    calls.append("preagg_dense_async")
    dts = val_dtypes if val_dtypes is not None else [dtype_code(v) for v in vals]
    chunk = ([k.data_ptr() for k in keys], [None if v is None else v.data_ptr() for v in vals], int(keys[0].numel()))
    calls.pop()
    out = preagg_dense_batch_async([chunk], accs, dts, len(keys), keys[0].device)
    calls[-1] = "preagg_dense_async"
    return out


def _merge_np(pieces, acc_ops, sort):
    """pieces in pointer form -> (n or -1, key arrays, merged accumulator arrays)"""
    nk, dts = len(pieces[0][0]), pieces[0][4]
    ks, vs = [[] for _ in range(nk)], [[] for _ in dts]
    for kp, ap, n, cptr, _, _ in pieces:
        if n is None:
            n = int(_view(cptr, 1, False)[0])
            if n < 0:
                return -1, None, None
        for j in range(nk):
            ks[j].append(_view(kp[j], n, False).copy())
        for a in range(len(dts)):
            vs[a].append(_view(ap[a], n, dts[a] == F64).copy())
    keys = [np.concatenate(k) for k in ks]
    vals = [np.concatenate(v) for v in vs]
    if len(keys[0]) > SMALL_MAX_ROWS:
        return -1, None, None
    k, a = _agg_np(keys, vals, [(i, op) for i, op in enumerate(acc_ops)], None, sort=sort)
    return len(k[0]), k, a


def merge_small_async(pieces, acc_ops, sort=False, checked=False):
    calls.append("merge_small_async")
    known = sum(p[2] for p in pieces if p[2] is not None)
    live = sum(1 for p in pieces if p[2] is None or p[2] > 0)
    if known > SMALL_MAX_ROWS or live > SMALL_MAX_PIECES or len(pieces[0]) != 6:
        return None
    nk, dts, dev = len(pieces[0][0]), pieces[0][4], pieces[0][5]
    fc = _K._fast(nk, dts, tuple((a, op) for a, op in enumerate(acc_ops)))
    bi, bf = fc.alloc(SMALL_MAX_ROWS, dev)
    n, k, a = _merge_np(pieces, acc_ops, sort)
    _fill_block(fc, bi, bf, k, a, n)
    return bi, bf, fc


def merge_small_post_async(pieces, acc_ops, post, sort=False):
    calls.append("merge_small_post_async")
    known = sum(p[2] for p in pieces if p[2] is not None)
    live = sum(1 for p in pieces if p[2] is None or p[2] > 0)
    if known > SMALL_MAX_ROWS or live > SMALL_MAX_PIECES:
        return None
    nk, dts, dev = len(pieces[0][0]), pieces[0][4], pieces[0][5]
    blk = torch.zeros((1 + nk + post.n + len(acc_ops), SMALL_MAX_ROWS), dtype=torch.int64, device=dev)
    n, k, a = _merge_np(pieces, acc_ops, sort)
    blk[0, 0] = n
    if n >= 0:
        for j in range(nk):
            blk[1 + j, :n] = torch.from_numpy(np.ascontiguousarray(k[j]))
        for j in range(post.n):
            st = post.steps[j]
            f = lambda i: a[i].astype(np.float64)   # noqa: E731
            with np.errstate(all="ignore"):
                if st.kind == POST_IDENTITY:
                    r = a[st.a0]
                elif st.kind == POST_MEAN:
                    r = f(st.a0) / f(st.a1)
                else:
                    r = orc.post_var(pd.Series(f(st.a1)), pd.Series(f(st.a0)), pd.Series(f(st.a2)), ddof=st.ddof).to_numpy()
            r = np.ascontiguousarray(r)
            blk[1 + nk + j, :n] = torch.from_numpy(r.view(np.int64) if r.dtype == np.float64 else r.astype(np.int64))
    return blk, nk, post.n


def result_block_to_host(blk, nk, n_post, is_float):
    calls.append("result_block_to_host")
    n = int(blk[0, 0])
    if n < 0:
        return n, None, None
    hn = blk.numpy()
    keys = [hn[1 + j, :n].copy() for j in range(nk)]
    cols = [hn[1 + nk + j, :n].view(np.float64).copy() if is_float[j] else hn[1 + nk + j, :n].copy() for j in range(n_post)]
    return n, keys, cols


def to_host_counted(columns, count_block, count_row):
    n = int(count_block[count_row, 0])
    if n < 0:
        return n, None
    return n, [c[:n].numpy().copy() for c in columns]


def to_host(columns):
    return [c.numpy().copy() for c in columns]


MID_GROUPS = 2048


def preagg_mid_batch_async(chunks, accs, val_dtypes, n_keys, device):
    """double of the mid-cardinality batch: sum / count over float64 columns, <= MID_GROUPS groups, no NaN"""
    calls.append("preagg_mid_batch_async")
    fc = _K._fast(n_keys, val_dtypes, accs)
    bi, bf = fc.alloc(MID_GROUPS, device)
    used = {c for c, _ in accs}
    keys = [np.concatenate([_view(kp[j], m, False) for kp, vp, m in chunks]) for j in range(n_keys)]
    vals, ok = [], all(op in (SUM, COUNT) for _, op in accs)
    for c, dt in enumerate(val_dtypes):
        if c in used and all(vp[c] is not None for _, vp, _ in chunks):
            v = np.concatenate([_view(vp[c], m, dt == F64) for _, vp, m in chunks])
            ok = ok and dt == F64 and not np.isnan(v).any()
            vals.append(v)
        else:
            vals.append(np.zeros(len(keys[0]), dtype=np.int64))
    k, a = _agg_np(keys, vals, accs, None)
    n = len(k[0]) if ok and len(k[0]) <= MID_GROUPS else -1
    _fill_block(fc, bi, bf, k, a, n)
    return bi, bf, fc


def estimate_distinct(keys):
    calls.append("estimate_distinct")
    n = int(keys[0].numel())
    s = min(n, 16384)
    if s == 0:
        return 0, 0
    stride = n // s
    idx = np.arange(s) * stride
    df = pd.DataFrame({j: _np(k)[idx] for j, k in enumerate(keys)})
    return s, int(len(df.drop_duplicates()))


def hash_partition(keys, n_partitions, cols):
    calls.append("hash_partition")
    pid = partition_ids(keys, n_partitions)
    calls.pop()
    out = partition_scatter(pid, n_partitions, cols)
    calls.pop()
    return out


def partition_ids_frame(keys, n_partitions):
    calls.append("partition_ids_frame")
    h = orc.hash_frame_rows([_np(k) for k in keys])
    return torch.from_numpy((h % np.uint64(n_partitions)).astype(np.int32))


make_row_program = _K.make_row_program


def _cmp(v, op, value):
    with np.errstate(invalid="ignore"):
        return [v <= value, v < value, v >= value, v > value, v == value, v != value][op]


def preagg_dense_batch_fused_async(chunks, accs, val_dtypes, n_keys, device, prog):
    """double of the fused pre-path: the row program is evaluated with numpy, then the dense double aggregates"""
    calls.append("preagg_dense_batch_fused_async")
    used = [c for c, dt in enumerate(val_dtypes) if any(col == c and op != COUNT for col, op in accs) or dt == F64]
    new_chunks, keep_alive = [], []
    for kp, vp, m in chunks:
        phys = [_view(vp[j], m, True).copy() for j in range(prog.n_phys)]
        keep = np.ones(m, dtype=bool)
        if prog.filter_col >= 0:
            raw = _view(vp[prog.filter_col], m, prog.filter_dtype == F64)
            keep = _cmp(raw, prog.filter_op, prog.filter_f64 if prog.filter_dtype == F64 else prog.filter_i64)
        logical = []
        for c in range(len(used)):
            nf = prog.n_factors[c]
            if nf == 0:
                logical.append(phys[c])
                continue
            v = None
            for f in range(nf):
                fac = prog.factors[c][f]
                term = fac.alpha + fac.beta * phys[fac.col]
                v = term if v is None else v * term
            logical.append(v)
        ks = [torch.from_numpy(_view(kp[j], m, False)[keep].copy()) for j in range(n_keys)]
        vs, li = [], 0
        for c, dt in enumerate(val_dtypes):
            if c in used:
                vs.append(torch.from_numpy(np.ascontiguousarray(logical[li][keep])))
                li += 1
            else:
                vs.append(None)
        keep_alive.append((ks, vs))
        new_chunks.append(([k.data_ptr() for k in ks], [None if v is None else v.data_ptr() for v in vs], int(keep.sum())))
    out = preagg_dense_batch_async(new_chunks, accs, val_dtypes, n_keys, device)
    calls.pop()
    return out


def eval_product(phys, factors):
    calls.append("eval_product")
    v = None
    for col, alpha, beta in factors:
        term = alpha + beta * _np(phys[col])
        v = term if v is None else v * term
    return torch.from_numpy(np.ascontiguousarray(v))


def filter_rows(filter_col, op, value, cols):
    calls.append("filter_rows")
    keep = _cmp(_np(filter_col), op, value)
    idx = torch.from_numpy(np.nonzero(keep)[0])
    return [c[idx] for c in cols], int(keep.sum())


# ---- two-pass partition kernel (mgb_split_plan / mgb_split_scatter).  Destination tables hold whatever "addresses"
# the transport double hands out; ``poke(address, values)`` (set by that double) stores 8-byte values there.
poke = None


def split_plan(keys, n_partitions, mode=0):
    calls.append("split_plan")
    pid = (orc.hash_frame_rows([_np(k) for k in keys]) % np.uint64(n_partitions)).astype(np.int64) if (mode & 3) == 1 \
        else orc.partition_ids([_np(k) for k in keys], n_partitions)
    counts = np.bincount(pid, minlength=n_partitions)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return torch.from_numpy(pid), torch.from_numpy(offsets)


def split_scatter(keys, n_partitions, hist, cols, outs=None, table=None, mode=0):
    calls.append("split_scatter")
    pid = _np(hist)
    order = np.argsort(pid, kind="stable")
    bounds = np.searchsorted(pid[order], np.arange(n_partitions + 1))
    for c, col in enumerate(cols):
        src = _np(col)[order]
        if table is None:
            outs[c].copy_(torch.from_numpy(src))
            continue
        tab = _np(table).reshape(len(cols), n_partitions)
        for p_ in range(n_partitions):
            blk = src[bounds[p_]:bounds[p_ + 1]]
            if len(blk):
                poke(int(tab[c, p_]), blk.view(np.int64))


def row_partials(vals, accs):
    """per-row partials of a chunk with ~distinct keys (mgb_row_partial): SUM / MIN / MAX alias the column"""
    calls.append("row_partials")
    out = []
    for col, op in accs:
        v = vals[col]
        if op in (_K.SUM, _K.MIN, _K.MAX):
            out.append(v)
        elif op == _K.COUNT:
            a = _np(v)
            out.append(torch.from_numpy((~np.isnan(a)).astype(np.int64) if a.dtype == np.float64 else np.ones(len(a), dtype=np.int64)))
        else:
            a = _np(v)
            out.append(torch.from_numpy(a * a))
    return out
