#!/usr/bin/env python
"""The five BASELINE.json configs through the operand API on ONE B200, at (scaled) full size.

    python tools/run_configs.py [--scale 1.0] [--configs cfg1,cfg2,cfg3,cfg4,cfg5]

Inputs are generated on the device (torch generators, seeded) chunk by chunk, so that multi-GB configs do not
wait for host random numbers; every run ends with size-independent checks (the oracle would take minutes at
these sizes; exact parity at oracle-sized inputs is what tests/ does):
  sum of counts == rows, keys unique, sum of sums == column total (1e-9), min <= max, groups == distinct keys
  computed independently with torch.unique on the raw keys.
Prints one JSON line per config: rows/s with the chunks resident in HBM -- `rows_per_s` over the whole operand
chain INCLUDING the fetch of the result frame to the host, `execute_rows_per_s` over the operand chain alone
(device-resident in -> result chunks out; large result chunks stay in HBM until the fetch) --, the plan taken,
the number of groups.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mars_b200 as md  # noqa: E402
from mars_b200.device import DeviceFrame  # noqa: E402

SEED = 20240229


def gen(cfg, scale):
    """-> (chunks, by, func, kwargs for groupby/agg, value columns)"""
    g = torch.Generator(device="cuda").manual_seed(SEED)

    def ri(lo, hi, n):
        return torch.randint(lo, hi, (n,), generator=g, device="cuda", dtype=torch.int64)

    def ru(n):
        return torch.rand(n, generator=g, device="cuda", dtype=torch.float64)

    def rn(n, mu, sd):
        return torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * sd + mu

    if cfg == "cfg1":
        n, chunk = int(10_000_000 * scale), 1_000_000
        mk = lambda m: ({"key": ri(0, 1000, m), **{f"v{i}": ru(m) for i in range(4)}})  # noqa: E731
        return n, chunk, mk, "key", ["sum", "mean", "count"], dict(sort=True, method="auto")
    if cfg == "cfg1-shuffle":
        n, chunk = int(10_000_000 * scale), 1_000_000
        mk = lambda m: ({"key": ri(0, 1000, m), **{f"v{i}": ru(m) for i in range(4)}})  # noqa: E731
        return n, chunk, mk, "key", ["sum", "mean", "count"], dict(sort=False, method="shuffle")
    if cfg == "cfg3":
        n, chunk = int(1_000_000_000 * scale), 15_625_000
        keys = int(100_000_000 * scale) or 1
        mk = lambda m: {"key": ri(0, keys, m), "v": ru(m)}  # noqa: E731
        return n, chunk, mk, "key", ["sum", "count"], dict(sort=False, method="shuffle")
    if cfg == "cfg4":
        n, chunk = int(500_000_000 * scale), 7_812_500
        mk = lambda m: {"k1": ri(0, 1000, m), "k2": ri(0, 1000, m), **{f"v{i}": rn(m, 10, 3) for i in range(8)}}  # noqa: E731
        return n, chunk, mk, ["k1", "k2"], ["min", "max", "var"], dict(sort=False, method="shuffle")
    if cfg == "cfg5":
        n, chunk = int(1_000_000_000 * scale), 15_625_000

        def mk(m):   # Zipf(1.2) by inverse transform of the bounded discrete power law (device side)
            u = ru(m)
            kmax = 1e8
            a = 1.2
            x = ((kmax ** (1 - a) - 1) * u + 1) ** (1 / (1 - a))
            return {"key": x.floor().clamp_(1, kmax).to(torch.int64), "v": ru(m)}

        return n, chunk, mk, "key", ["mean", "count"], dict(sort=False, method="shuffle")
    raise SystemExit(cfg)


def run(cfg, scale, repeats):
    n, chunk, mk, by, func, kw = gen(cfg, scale)
    frames = []
    for s in range(0, n, chunk):
        cols = mk(min(chunk, n - s))
        frames.append(DeviceFrame([], None, list(cols.keys()), list(cols.values()), min(chunk, n - s)))
    torch.cuda.synchronize()
    by_list = [by] if isinstance(by, str) else by
    times, exec_times = [], []
    out = None
    for it in range(repeats + 1):
        sess = md.new_session(default=False)
        r = md.DataFrame.from_chunks(frames).groupby(by, sort=kw["sort"]).agg(func, method=kw["method"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r.execute(session=sess)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        out = r.fetch(session=sess)
        torch.cuda.synchronize()
        if it:
            times.append(time.perf_counter() - t0)
            exec_times.append(t1 - t0)
    # size-independent checks
    vcol = [c for c in frames[0].columns if c not in by_list][0]
    checks = {}
    if "count" in func:
        checks["sum_of_counts_eq_rows"] = bool(int(out[(vcol, "count")].sum()) == n)
    checks["keys_unique"] = bool(out.index.is_unique)
    if len(by_list) == 1:
        distinct = sum(1 for _ in [0])  # placeholder replaced below
        allk = torch.cat([f.data[f.columns.index(by_list[0])] for f in frames])
        distinct = int(torch.unique(allk).numel())
        del allk
    else:
        code = torch.cat([f.data[f.columns.index(by_list[0])] * 1_000_003 + f.data[f.columns.index(by_list[1])] for f in frames])
        distinct = int(torch.unique(code).numel())
        del code
    checks["groups_eq_distinct_keys"] = bool(len(out) == distinct)
    if "sum" in func:
        total = float(sum(f.data[f.columns.index(vcol)].sum() for f in frames))
        checks["sum_of_sums_eq_total_1e-9"] = bool(abs(float(out[(vcol, "sum")].sum()) - total) <= 1e-9 * abs(total))
    if "mean" in func and "count" in func:
        total = float(sum(f.data[f.columns.index(vcol)].sum() for f in frames))
        rec = float((out[(vcol, "mean")] * out[(vcol, "count")]).sum())
        checks["mean_times_count_eq_total_1e-9"] = bool(abs(rec - total) <= 1e-9 * abs(total))
    if "min" in func:
        checks["min_le_max"] = bool((out[(vcol, "min")] <= out[(vcol, "max")]).all())
        checks["global_min_max"] = bool(float(out[(vcol, "min")].min()) == float(min(f.data[f.columns.index(vcol)].min() for f in frames))
                                        and float(out[(vcol, "max")].max()) == float(max(f.data[f.columns.index(vcol)].max() for f in frames)))
    if kw["sort"]:
        checks["sorted"] = bool(out.index.is_monotonic_increasing)
    rec = {"config": cfg, "scale": scale, "rows": n, "chunks": len(frames), "groups": int(len(out)),
           "result_chunks": len(r.data.chunks), "plan": sorted(set(sess.executed_ops)),
           "seconds": times, "rows_per_s": n / min(times),
           "execute_seconds": exec_times, "execute_rows_per_s": n / min(exec_times), "checks": checks,
           "all_checks_pass": all(checks.values())}
    print(json.dumps(rec), flush=True)
    del frames, out
    torch.cuda.empty_cache()
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--configs", default="cfg1,cfg1-shuffle,cfg3,cfg4,cfg5")
    ap.add_argument("--repeats", type=int, default=2)
    a = ap.parse_args()
    for c in a.configs.split(","):
        run(c, 1.0 if c.startswith("cfg1") else a.scale, a.repeats)
