"""Wall-clock share of every operand class in one bench step (no profiler overhead):
python profiles/host_breakdown.py [rows_per_chunk]"""
import collections
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mars_b200 as md  # noqa: E402
from mars_b200 import core  # noqa: E402
from mars_b200.device import DeviceFrame  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
frames = []
for c in range(15):
    cols = bench.gen_chunk(c, rows)
    frames.append(DeviceFrame([], None, bench.COLUMNS, [torch.from_numpy(cols[n]).cuda() for n in bench.COLUMNS], rows))

acc = collections.defaultdict(float)
cnt = collections.Counter()
orig = core.execute


import cProfile  # noqa: E402
import pstats  # noqa: E402
PROF = cProfile.Profile()
PROFILE_STAGE = os.environ.get("PROFILE_STAGE")     # e.g. "agg", "combine", "map"


def timed_execute(results, op):
    t0 = time.perf_counter()
    prof = PROFILE_STAGE is not None and getattr(op.stage, "name", None) == PROFILE_STAGE
    if prof:
        PROF.enable()
    try:
        return orig(results, op)
    finally:
        if prof:
            PROF.disable()
        k = f"{type(op).__name__}:{getattr(op.stage, 'name', None)}"
        acc[k] += time.perf_counter() - t0
        cnt[k] += 1


core.execute = timed_execute
phases = collections.defaultdict(float)


def step():
    t0 = time.perf_counter()
    mdf = md.DataFrame.from_chunks(frames)
    sess = md.new_session(default=False)
    r = mdf.groupby(bench.BY).agg(bench.Q1_FUNC, method="auto")
    t1 = time.perf_counter()
    r.execute(session=sess)
    t2 = time.perf_counter()
    out = r.fetch(session=sess)
    t3 = time.perf_counter()
    phases["build graph"] += t1 - t0
    phases["execute (tile + run chunks)"] += t2 - t1
    phases["fetch"] += t3 - t2
    return out


for _ in range(5):
    step()
acc.clear(); cnt.clear(); phases.clear()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    step()
total = time.perf_counter() - t0
print(f"step: {1e3 * total / N:.3f} ms")
for k, v in phases.items():
    print(f"  {k:32s} {1e3 * v / N:7.3f} ms")
in_ops = sum(acc.values())
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"    {k:36s} {cnt[k] // N:3d} calls  {1e6 * v / cnt[k]:7.1f} us each  {1e3 * v / N:7.3f} ms per step")
print(f"    session loop outside executors: {1e3 * (phases['execute (tile + run chunks)'] - in_ops) / N:.3f} ms per step")

if PROFILE_STAGE:
    pstats.Stats(PROF).sort_stats("tottime").print_stats(28)
