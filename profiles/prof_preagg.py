"""Driver for ncu captures of the map-side pre-aggregation kernels on one Q1-shaped chunk (and other shapes).

    python profiles/prof_preagg.py [q1|cfg1|cfg3|cfg4] [rows] [iters]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mars_b200 import kernels as K  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "q1"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(1)
dev = "cuda"


def rnd(n):
    return torch.rand(n, generator=g, device=dev, dtype=torch.float64)


def ri(lo, hi, n):
    return torch.randint(lo, hi, (n,), generator=g, device=dev, dtype=torch.int64)


if shape == "q1":
    keys = [ri(0, 3, rows), ri(0, 2, rows)]
    vals = [rnd(rows) for _ in range(5)] + [ri(1, 1 << 25, rows)]
    accs = [(c, K.SUM) for c in range(5)] + [(0, K.COUNT), (1, K.COUNT), (2, K.COUNT), (5, K.COUNT)]
elif shape == "cfg1":
    keys = [ri(0, 1000, rows)]
    vals = [rnd(rows) for _ in range(4)]
    accs = [(c, K.SUM) for c in range(4)] + [(c, K.COUNT) for c in range(4)]
elif shape == "cfg3":
    keys = [ri(0, 100_000_000, rows)]
    vals = [rnd(rows)]
    accs = [(0, K.SUM), (0, K.COUNT)]
elif shape == "cfg4":
    keys = [ri(0, 1000, rows), ri(0, 1000, rows)]
    vals = [rnd(rows) for _ in range(8)]
    accs = [(c, op) for op in (K.MIN, K.MAX, K.COUNT, K.SUMSQ, K.SUM) for c in range(8)]
else:
    raise SystemExit(shape)
bytes_in = rows * 8 * (len(keys) + len(vals))
import ctypes as C  # noqa: E402
from mars_b200 import _lib  # noqa: E402
lib = _lib.load()
NAMES = ["preagg_dense", "preagg_smem/hash", "merge", "update_global", "table_init", "emit", "hash", "partition", "sort", "post"]
for it in range(iters):
    lib.mgb_prof_enable(1 if it == iters - 1 else 0)
    lib.mgb_prof_reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = K.preagg_dense(keys, vals, accs) if os.environ.get("MGB_NO_DENSE") is None else None
    stats = "dense"
    if res is None:
        k, a, stats = K.groupby_aggregate(keys, vals, accs, sort=False,
                                          expect_distinct=os.environ.get("MGB_DISTINCT") is not None)
    else:
        k, a = res
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{shape} rows={rows} groups={k[0].numel()} {ms:.3f} ms  {bytes_in / ms / 1e6:.1f} GB/s  "
          f"wall {1e3 * (time.perf_counter() - t0):.3f} ms  {stats}", flush=True)
prof = {}
for kid, name in enumerate(NAMES):
    tot, cnt = C.c_double(), C.c_int64()
    lib.mgb_prof_read(kid, C.byref(tot), C.byref(cnt))
    if cnt.value:
        prof[name] = f"{1e3 * tot.value:.1f} us / {cnt.value}"
print("kernel time (CUDA events, last iteration):", prof)
