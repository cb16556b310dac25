"""Per-step timing of the end-to-end (host chunks) path: python profiles/e2e_steps.py [steps]"""
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mars_b200 as md  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
sizes = bench.chunk_sizes(bench.ROWS_PER_GPU, bench.CHUNK_ROWS)
host_frames = []
for c, rows in enumerate(sizes):
    cols = bench.gen_chunk(c, rows)
    flat = torch.empty((len(bench.COLUMNS), rows), dtype=torch.int64).pin_memory()
    data = {}
    for j, name in enumerate(bench.COLUMNS):
        view = flat[j].numpy().view(cols[name].dtype)
        view[:] = cols[name]
        data[name] = view
    host_frames.append(pd.DataFrame(data, copy=False))
for it in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sess = md.new_session(default=False)
    r = md.DataFrame.from_chunks(host_frames).to_gpu().groupby(bench.BY).agg(bench.Q1_FUNC, method="auto")
    r.execute(session=sess)
    out = r.fetch(session=sess)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"step {it}: {1e3 * (t1 - t0):.1f} ms  mem allocated {torch.cuda.memory_allocated() / 1e9:.2f} GB reserved "
          f"{torch.cuda.memory_reserved() / 1e9:.2f} GB", flush=True)
