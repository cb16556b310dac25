"""One chunk through the bucketed aggregation (distinct-keys hint), cfg3- and cfg4-shaped:
python profiles/prof_bucket.py cfg3|cfg4 [rows]   -- prints per-call wall time and the kernel split"""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from mars_b200 import _lib, kernels as K  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
g = torch.Generator(device="cuda").manual_seed(1)
if shape == "cfg3":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 15_625_000
    keys = [torch.randint(0, 25_000_000, (n,), generator=g, device="cuda", dtype=torch.int64)]
    vals = [torch.rand(n, generator=g, device="cuda", dtype=torch.float64)]
    accs = [(0, K.SUM), (0, K.COUNT)]
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 7_812_500
    keys = [torch.randint(0, 1000, (n,), generator=g, device="cuda", dtype=torch.int64) for _ in range(2)]
    vals = [torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 3 + 10 for _ in range(8)]
    accs = [(c, o) for o in (K.MIN, K.MAX, K.SUM, K.SUMSQ, K.COUNT) for c in range(8)]
lib = _lib.load()
for it in range(4):
    if it == 3:
        lib.mgb_prof_enable(1)
        lib.mgb_prof_reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k, a, stats = K.groupby_aggregate(keys, vals, accs, sort=False, expect_distinct=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{shape} {n} rows -> {k[0].numel()} groups: {1e3 * dt:.3f} ms = {n / dt / 1e9:.2f} G rows/s  {stats}")
for kid, name in enumerate(["dense", "preagg", "merge", "update_global", "table_init", "emit", "hash", "partition", "sort",
                            "post", "bucket"]):
    tot, c = C.c_double(), C.c_int64()
    lib.mgb_prof_read(kid, C.byref(tot), C.byref(c))
    if c.value:
        print(f"    {name:14s} {tot.value:8.3f} ms / {c.value}")
if __import__("os").environ.get("TIMELINE"):
    import collections
    import json

    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        K.groupby_aggregate(keys, vals, accs, sort=False, expect_distinct=True)
        torch.cuda.synchronize()
    prof.export_chrome_trace("/tmp/prof_bucket_trace.json")
    ev = json.load(open("/tmp/prof_bucket_trace.json"))["traceEvents"]
    by = collections.defaultdict(lambda: [0.0, 0])
    for e in ev:
        if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
            nm = e["name"].split("(")[0].split("<")[0]
            by[nm][0] += e["dur"]
            by[nm][1] += 1
    for nm, (d, c) in sorted(by.items(), key=lambda kv: -kv[1][0]):
        print(f"    {d:9.1f} us {c:4d} x {nm}")
