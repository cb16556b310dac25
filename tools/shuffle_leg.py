"""Hash-shuffle branch across N GPUs (one process per GPU): whole-path rows/s, NVLink bytes/s of the fused
partition -> peer transfer, parity.  Used by ``bench.py --gpus N`` (the ``shuffle`` object of the bench line) and by
``tools/shuffle_ngpu.py``.

Workload: a BASELINE.json shuffle config (default cfg3: int64 keys drawn from 1e8, one float64 column, sum / count,
``method='shuffle'``, sort=False) at ``rows_per_gpu`` rows per GPU -- weak scaling; 8 x 125 M rows is the config's
full size.  Global chunk c is generated on the host with ``default_rng(SEED + c)`` (tools/baseline_configs) by the
rank that owns it (c % N == rank, the session's placement of data-source chunks) and uploaded before the timed region.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import baseline_configs as bc  # noqa: E402

NVLINK_NOMINAL_GBPS = 900.0      # per GPU per direction (NVLink 5)


def _frames(cfg, rank, world, chunks_per_gpu, chunk_rows, dev, keep_host=True):
    import torch

    from mars_b200.device import DeviceFrame

    cols = cfg.columns()
    probe = cfg.gen(0, 1)
    frames, host = [], {}
    for c in range(chunks_per_gpu * world):
        if c % world == rank:
            data = cfg.gen(c, chunk_rows)
            if keep_host:
                host[c] = data
            ts = [torch.from_numpy(data[k]).to(dev) for k in cols]
            frames.append(DeviceFrame([], None, cols, ts, chunk_rows))
        else:       # same schema, never executed on this rank
            ts = [torch.empty(0, dtype=torch.from_numpy(probe[k]).dtype, device=dev) for k in cols]
            frames.append(DeviceFrame([], None, cols, ts, 0))
    return frames, host


def _lazy(cfg, frames):
    import mars_b200 as md

    by = cfg.by if len(cfg.by) > 1 else cfg.by[0]
    return md.DataFrame.from_chunks(frames).groupby(by, sort=cfg.sort).agg(cfg.func, method="shuffle")


def run(ex, cfg_name: str = "cfg3", rows_per_gpu: int = 125_000_000, chunk_rows: int = 15_625_000, steps: int = 3,
        warmup: int = 2, parity_rows_per_gpu: int = 1_000_000, keep_host: bool = False):
    """-> dict on rank 0 (None elsewhere).  ``ex``: parallel.Exchange of this process."""
    import pandas as pd
    import torch
    import torch.distributed as dist

    import mars_b200 as md

    cfg = bc.CONFIGS[cfg_name]
    rank, world = ex.rank, ex.world
    dev = torch.device("cuda", torch.cuda.current_device())
    t = ex.transport
    group = ex.meta_group

    def barrier():
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def allmax(x):
        v = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX, group=group)
        return float(v.item())

    def allsum(x):
        v = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        return float(v.item())

    # ---- parity at a size the CPU checker finishes in seconds: fetched frame vs the oracle on the same bytes
    pr_chunk = max(parity_rows_per_gpu // 2, 1)
    frames, host = _frames(cfg, rank, world, 2, pr_chunk, dev)
    sess = md.new_session(exchange=ex, default=False)
    r = _lazy(cfg, frames)
    r.execute(session=sess)
    got = sess.fetch(r)
    gathered = [None] * world
    dist.all_gather_object(gathered, host, group=group)
    parity = None
    if rank == 0:
        import mars_groupby_oracle as orc

        allc = {}
        for h in gathered:
            allc.update(h)
        raw = pd.concat([pd.DataFrame(allc[c], copy=False) for c in sorted(allc)], axis=0, ignore_index=True)
        exp = orc.run_groupby_agg(raw, cfg.by, cfg.func, pr_chunk, method="shuffle", sort=cfg.sort)
        g, e = got.sort_index(), exp.sort_index()
        ok = list(g.columns) == list(e.columns) and g.index.equals(e.index)
        if ok:
            for c in e.columns:
                a, b = g[c].to_numpy(), e[c].to_numpy()
                exact = b.dtype.kind in "iu" or c[-1] in ("count", "min", "max")
                ok = ok and (np.array_equal(a, b) if exact else np.allclose(a, b, rtol=1e-9, atol=0, equal_nan=True))
        parity = {"ok": bool(ok), "rows": int(len(raw)), "groups": int(len(e)),
                  "mode": "fetched frame vs the oracle (method='shuffle') on the same host-generated chunks: keys / counts "
                          "bit-exact, float64 within 1e-9"}
    del frames, host, gathered, got, r, sess

    # ---- full size: whole path, then the same steps with the ranks lined up before the transfer (link rate)
    chunks_per_gpu = max(1, rows_per_gpu // chunk_rows)
    rows_rank = chunks_per_gpu * chunk_rows
    t0 = time.perf_counter()
    frames, _ = _frames(cfg, rank, world, chunks_per_gpu, chunk_rows, dev, keep_host=False)
    gen_s = time.perf_counter() - t0
    total_rows = rows_rank * world

    def step():
        sess = md.new_session(exchange=ex, default=False)
        r = _lazy(cfg, frames)
        r.execute(session=sess)
        return sess, r

    times = []
    sess = None
    for it in range(warmup + steps):
        sess = None
        barrier()
        t0 = time.perf_counter()
        sess, r = step()
        barrier()
        if it >= warmup:
            times.append(allmax(time.perf_counter() - t0))
    groups = allsum(sum(len(v) for v in sess._results.values()))
    # invariant at full size: the counts of all groups add up to the number of rows
    # (configs without a count column -- cfg4: min / max / var -- have no such invariant: None)
    cnt_total, has_count = 0, False
    for v in sess._results.values():
        labels = getattr(v, "labels", None)      # executors.DeviceResult: result chunk resident in HBM
        if labels is not None:
            col = next((c for lab, c in zip(labels, v.columns) if isinstance(lab, tuple) and lab[-1] == "count"), None)
            if col is not None:
                has_count = True
                cnt_total += int(col.sum().item())
        else:                                    # small result chunk handed back as pandas
            lab = next((lab for lab in v.columns if isinstance(lab, tuple) and lab[-1] == "count"), None)
            if lab is not None:
                has_count = True
                cnt_total += int(v[lab].sum())
    has_count = allsum(1 if has_count else 0) > 0
    cnt_total = allsum(cnt_total)
    ops = sorted(set(sess.executed_ops))
    del sess, r
    link = None
    if hasattr(t, "push_stats") or hasattr(t, "exchange_stats"):
        t.timed = True
        if hasattr(t, "push_events"):
            t.push_events.clear()
        t._events.clear()
        for _ in range(max(2, steps)):
            sess, r = step()
            barrier()
            del sess, r
        t.timed = False
        ms, nbytes = t.push_stats() if getattr(t, "push_events", None) else t.exchange_stats()
        runs = max(2, steps)
        worst_ms = allmax(ms) / runs
        sent = allmax(nbytes) / runs
        link = {"bytes_sent_per_gpu": sent, "device_ms_max_over_ranks": worst_ms,
                "GBps_per_gpu_per_direction": sent / (worst_ms * 1e-3) / 1e9 if worst_ms > 0 else None,
                "frac_of_900": sent / (worst_ms * 1e-3) / 1e9 / NVLINK_NOMINAL_GBPS if worst_ms > 0 else None,
                "how": ("CUDA events around the partition kernels of a rank's mappers, which store every row straight into "
                        "the receive arena of the GPU that reduces it (ranks lined up first); bytes = rows stored into "
                        "PEER arenas x row bytes; max over ranks" if getattr(t, "push_events", None) else
                        "CUDA events around the transport call (ranks lined up first); max over ranks")}
    # one more step with per-kernel CUDA events (rank 0's kernels; untimed)
    kernel_ms = None
    try:
        import ctypes as C

        from mars_b200 import _lib

        lib = _lib.load()
        lib.mgb_prof_enable(1)
        lib.mgb_prof_reset()
        sess, r = step()
        barrier()
        del sess, r
        kernel_ms = {}
        for kid, nm in enumerate(["preagg_dense", "preagg_smem_or_mid", "merge", "update_global", "table_init", "emit", "hash",
                                  "partition", "sort", "post_or_row_partial", "bucket"]):
            tot, cnt = C.c_double(), C.c_int64()
            lib.mgb_prof_read(kid, C.byref(tot), C.byref(cnt))
            if cnt.value:
                kernel_ms[nm] = {"ms_total": round(tot.value, 3), "launches": cnt.value}
        lib.mgb_prof_enable(0)
    except Exception as e:  # noqa: BLE001
        kernel_ms = {"error": repr(e)[:200]}
    del frames
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    sec = sum(times) / len(times)
    return {"config": cfg_name, "transport": type(t).__name__, "n_gpus": world, "rows": total_rows, "rows_per_gpu": rows_rank,
            "chunk_rows": chunk_rows, "chunks_per_gpu": chunks_per_gpu, "groups": int(groups),
            "value": total_rows / sec, "unit": "rows/s", "ms_per_step": sec * 1e3, "steps": steps, "warmup": warmup,
            "seconds": times, "sum_of_counts_eq_rows": bool(int(cnt_total) == total_rows) if has_count else None,
            "link": link, "parity": parity, "ops": ops, "kernel_time_ms_rank0": kernel_ms,
            "input": f"host numpy default_rng({bc.SEED} + chunk), uploaded before the timed region ({gen_s:.1f} s)",
            "what": "map (pre-aggregation per chunk) -> partition kernel writing into peer arenas -> reduce -> agg; result "
                    "chunks stay in HBM on the rank of their reducer"}
