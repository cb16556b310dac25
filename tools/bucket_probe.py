"""debug probe: bucketed aggregation on a few input shapes, each under its own timeout"""
import subprocess
import sys
import time

CASES = sys.argv[2:] if len(sys.argv) > 2 and sys.argv[1] == "--cases" else ["uniform", "hot", "hot_small", "two_keys", "wide"]

if len(sys.argv) == 2:
    import numpy as np
    import torch
    from mars_b200 import kernels as K

    case = sys.argv[1]
    rng = np.random.default_rng(1)
    n = 300_000
    keys = [rng.integers(-2 ** 62, 2 ** 62, n, dtype=np.int64)]
    vals = [rng.normal(0, 1, n), rng.random(n)]
    accs = [(0, K.SUM), (1, K.SUM), (0, K.COUNT), (1, K.COUNT)]
    if case == "hot":
        keys[0][::7] = keys[0][0]
    if case == "hot_small":
        n = 20000
        keys = [keys[0][:n].copy()]
        keys[0][::7] = keys[0][0]
        vals = [v[:n] for v in vals]
    if case == "two_keys":
        keys = [rng.integers(0, 300, n, dtype=np.int64), rng.integers(0, 200, n, dtype=np.int64)]
    if case == "wide":
        keys = [rng.integers(0, 300, n, dtype=np.int64), rng.integers(0, 200, n, dtype=np.int64)]
        vals = [rng.normal(10, 3, n) for _ in range(8)]
        accs = [(c, o) for o in (K.MIN, K.MAX, K.SUM, K.SUMSQ, K.COUNT) for c in range(8)]
    dk = [torch.from_numpy(k).cuda() for k in keys]
    dv = [torch.from_numpy(v).cuda() for v in vals]
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k, a, stats = K.groupby_aggregate(dk, dv, accs, sort=False, expect_distinct=True)
        torch.cuda.synchronize()
        print(case, it, f"{1e3 * (time.perf_counter() - t0):.2f} ms", k[0].numel(), stats, flush=True)
else:
    for c in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, c], timeout=40, capture_output=True, text=True)
            print(r.stdout, r.stderr[-2000:], flush=True)
        except subprocess.TimeoutExpired as e:
            print(c, "TIMEOUT", (e.stdout or b"")[-2000:], (e.stderr or b"")[-2000:], flush=True)
