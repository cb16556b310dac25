"""cProfile of one bench step's host side (small chunks so that kernels are negligible)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mars_b200 as md  # noqa: E402
from mars_b200.device import DeviceFrame  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
frames = []
for c in range(15):
    cols = bench.gen_chunk(c, rows)
    frames.append(DeviceFrame([], None, bench.COLUMNS, [torch.from_numpy(cols[n]).cuda() for n in bench.COLUMNS], rows))


def step():
    mdf = md.DataFrame.from_chunks(frames)
    sess = md.new_session(default=False)
    r = mdf.groupby(bench.BY).agg(bench.Q1_FUNC, method="auto")
    r.execute(session=sess)
    return r.fetch(session=sess)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
print(f"step: {1e3 * (time.perf_counter() - t0) / 10:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
