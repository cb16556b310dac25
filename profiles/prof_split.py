"""Partition kernel (mgb_split_plan + mgb_split_scatter) on shuffle-shaped partials: CUDA-event time per pass and
achieved HBM rate.  usage: python profiles/prof_split.py [rows] ; prints one JSON line per shape."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mars_b200 import kernels as K  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 15_625_000
peak = 6454.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, nk, ncols, R in (("cfg3 partial (key,sum,count)", 1, 3, 8), ("cfg3 R=64", 1, 3, 64), ("cfg3 R=256", 1, 3, 256),
                           ("cfg4 partial (2 keys + 40 accs)", 2, 42, 8), ("cfg5 partial (key,sum,count,count)", 1, 4, 8)):
    n = rows if ncols < 20 else rows // 2
    g = torch.Generator(device="cuda").manual_seed(1)
    keys = [torch.randint(0, 1 << 40, (n,), dtype=torch.int64, device="cuda", generator=g) for _ in range(nk)]
    cols = keys + [torch.rand(n, dtype=torch.float64, device="cuda", generator=g) for _ in range(ncols - nk)]
    outs = [torch.empty_like(c) for c in cols]
    state = {}

    def plan():
        state["h"], state["o"] = K.split_plan(keys, R)

    def scatter():
        K.split_scatter(keys, R, state["h"], cols, outs=outs)

    t_plan = timeit(plan)
    t_scatter = timeit(scatter)
    alg = n * ncols * 8 * 2
    print(json.dumps({"shape": name, "rows": n, "cols": ncols, "R": R, "plan_ms": t_plan, "scatter_ms": t_scatter,
                      "scatter_GBps": alg / t_scatter / 1e6, "scatter_frac_of_hbm": alg / t_scatter / 1e6 / peak,
                      "both_GBps": (alg + 8 * nk * n) / (t_plan + t_scatter) / 1e6,
                      "both_frac_of_hbm": (alg + 8 * nk * n) / (t_plan + t_scatter) / 1e6 / peak,
                      "rows_per_s": n / (t_plan + t_scatter) * 1e3}), flush=True)
    del keys, cols, outs, state
    torch.cuda.empty_cache()
