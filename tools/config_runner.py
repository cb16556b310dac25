"""One BASELINE.json config through the operand API on the current GPU: value, roofline fraction, parity.

Used by ``bench.py`` (sub-records of the bench line) and ``tools/run_configs.py``.  Inputs come from
``tools/baseline_configs`` -- host numpy ``default_rng(SEED + chunk)`` per chunk, uploaded before the timed
region (device-resident in -> device-resident out; the D2H of the final frame is reported
separately) -- so the checker (the oracle, on a sample of the groups at sizes where the full frame would take
minutes on the CPU) sees the same bytes.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import baseline_configs as bc  # noqa: E402


def _peak_gbs() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6650.0


def check_parity(cfg: bc.Config, chunks, got, n_sample=2048, rtol=1e-9):
    """oracle parity on a random sample of the groups (keys / counts / min / max bit-exact, float64 within rtol)
    plus size-independent invariants of the whole frame.  Returns (ok, details)."""
    import pandas as pd

    details = {}
    rows = sum(len(next(iter(c.values()))) for c in chunks)
    details["keys_unique"] = bool(got.index.is_unique)
    vcol = [c for c in chunks[0] if c not in cfg.by][0]
    funcs = cfg.func if isinstance(cfg.func, list) else None
    if funcs and "count" in funcs:
        details["sum_of_counts_eq_rows"] = bool(int(got[(vcol, "count")].sum()) == rows)
    if funcs and "sum" in funcs:
        total = float(sum(c[vcol].sum() for c in chunks))
        details["sum_of_sums_eq_total"] = bool(abs(float(got[(vcol, "sum")].sum()) - total) <= 1e-9 * abs(total))
    if cfg.sort:
        details["sorted"] = bool(got.index.is_monotonic_increasing)
    exp, sample = bc.sampled_key_oracle(cfg, chunks, got, n_sample=n_sample)
    ok = True
    try:
        assert list(sample.columns) == list(exp.columns)
        assert sample.index.equals(exp.index)
        for c in exp.columns:
            g, e = sample[c].to_numpy(), exp[c].to_numpy()
            assert g.dtype == e.dtype
            func = c[-1] if isinstance(c, tuple) else None
            if e.dtype.kind in "iu" or func in ("count", "min", "max"):
                assert np.array_equal(g, e), c
            else:
                assert np.allclose(g, e, rtol=rtol, atol=0, equal_nan=True), c
        details["oracle_sampled_groups"] = int(len(exp))
    except AssertionError as e:
        ok = False
        details["oracle_mismatch"] = repr(e)[:200]
    ok = ok and all(v for k, v in details.items() if isinstance(v, bool))
    del pd
    return ok, details


def run_record(name: str, rows: int, chunk_rows: int = None, repeats: int = 2, parity: bool = True,
               n_sample: int = 2048, fuse: bool = True):
    """-> dict: rows/s over the operand chain (chunks resident in HBM, result chunks left in HBM / tiny results
    read back), fetch time, algorithmic-bytes roofline fraction, parity."""
    import pandas as pd
    import torch

    import mars_b200 as md
    from mars_b200 import kernels
    from mars_b200.device import DeviceFrame

    cfg = bc.CONFIGS[name]
    chunk_rows = chunk_rows or cfg.chunk_rows
    t0 = time.perf_counter()
    chunks = bc.host_chunks(cfg, rows, chunk_rows)
    gen_s = time.perf_counter() - t0
    dev = torch.device("cuda", torch.cuda.current_device())
    cols = list(chunks[0].keys())
    frames = []
    for c in chunks:
        ts = [torch.from_numpy(c[k]).to(dev) for k in cols]
        frames.append(DeviceFrame([], None, cols, ts, len(c[cols[0]])))
    torch.cuda.synchronize()
    by = cfg.by if len(cfg.by) > 1 else cfg.by[0]
    times, fetch_times, out, sess = [], [], None, None
    launches = 0
    import ctypes as C

    from mars_b200 import _lib

    lib = _lib.load()
    kernel_ms = None
    warm = 2        # untimed: the first run learns the shape (cardinality probe, first-chunk paths), the second grows the pools
    for it in range(repeats + warm + 1):
        profiled = it == repeats + warm         # one extra, untimed run with per-kernel CUDA events
        if profiled:
            lib.mgb_prof_enable(1)
            lib.mgb_prof_reset()
        sess = md.new_session(default=False)
        sess.fuse = fuse
        r = md.DataFrame.from_chunks(frames).groupby(by, sort=cfg.sort).agg(cfg.func, method=cfg.method)
        torch.cuda.synchronize()
        l0 = kernels.launch_counter
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        r.execute(session=sess)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        out = r.fetch(session=sess)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if profiled:
            kernel_ms = {}
            names = ["preagg_dense", "preagg_smem_or_mid", "merge", "update_global", "table_init", "emit", "hash",
                     "partition", "sort", "post", "bucket"]
            for kid, nm in enumerate(names):
                tot, cnt = C.c_double(), C.c_int64()
                lib.mgb_prof_read(kid, C.byref(tot), C.byref(cnt))
                if cnt.value:
                    kernel_ms[nm] = {"ms_total": tot.value, "launches": cnt.value}
            lib.mgb_prof_enable(0)
        elif it >= warm:
            times.append(max(t1 - t0, e0.elapsed_time(e1) * 1e-3))
            fetch_times.append(t2 - t1)
            launches = kernels.launch_counter - l0
    sec = sum(times) / len(times)
    groups = int(len(out))
    alg = cfg.algorithmic_bytes(rows, groups)
    peak = _peak_gbs()
    rec = {"config": name, "rows": rows, "chunk_rows": chunk_rows, "chunks": len(frames), "groups": groups,
           "by": cfg.by, "func": cfg.func, "sort": cfg.sort, "method": cfg.method,
           "plan": sorted(set(sess.executed_ops)), "fused_trees": sess.fused_trees,
           "value": rows / sec, "unit": "rows/s", "seconds": times, "fetch_seconds": fetch_times,
           "gpu_launches": int(launches), "kernel_time_ms": kernel_ms,
           "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / sec / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": alg / sec / 1e9 / peak},
           "input": f"host numpy default_rng({bc.SEED} + chunk), uploaded before the timed region; {gen_s:.1f} s to generate"}
    if parity:
        t0 = time.perf_counter()
        ok, details = check_parity(cfg, chunks, out, n_sample=n_sample)
        rec["parity"] = {"ok": bool(ok), "mode": f"oracle on {details.get('oracle_sampled_groups', 0)} sampled groups "
                                                 f"(bit-exact keys/counts/min/max, 1e-9 float) + invariants",
                         "details": details, "seconds": time.perf_counter() - t0}
    del frames, out, chunks
    torch.cuda.empty_cache()
    del pd
    return rec
