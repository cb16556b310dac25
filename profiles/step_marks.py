"""Bench step (cfg2, resident chunks): host time BEFORE the band kernel is enqueued and AFTER its result has arrived --
the two parts of a step during which the GPU idles.    python profiles/step_marks.py [rows_per_chunk]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mars_b200 as md  # noqa: E402
from mars_b200 import executors as E, kernels as K  # noqa: E402
from mars_b200.device import DeviceFrame  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
frames = []
for c in range(15):
    cols = bench.gen_chunk(c, rows)
    frames.append(DeviceFrame([], None, bench.COLUMNS, [torch.from_numpy(cols[n]).cuda() for n in bench.COLUMNS], rows))
torch.cuda.synchronize()
marks = {}


def wrap(mod, name, tag):
    orig = getattr(mod, name)

    def f(*a, **k):
        marks[tag + ":in"] = time.perf_counter()
        r = orig(*a, **k)
        marks[tag + ":out"] = time.perf_counter()
        return r

    setattr(mod, name, f)


wrap(K, "_preagg_batch", "launch")
wrap(K, "result_block_to_host", "d2h")
wrap(E, "_assemble_result", "assemble")
wrap(K, "merge_small_post_async", "merge")


def step():
    marks.clear()
    marks["t0"] = time.perf_counter()
    mdf = md.DataFrame.from_chunks(frames)
    sess = md.new_session(default=False)
    r = mdf.groupby(bench.BY).agg(bench.Q1_FUNC, method="auto")
    marks["lazy"] = time.perf_counter()
    r.execute(session=sess)
    marks["executed"] = time.perf_counter()
    out = r.fetch(session=sess)
    marks["t1"] = time.perf_counter()
    return out


for _ in range(10):
    step()
acc = {}
N = 50
for _ in range(N):
    torch.cuda.synchronize()
    step()
    m = dict(marks)
    seq = [("build lazy graph", "t0", "lazy"), ("execute() entry -> band launch call", "lazy", "launch:in"),
           ("band launch call (ctypes + cudaLaunch)", "launch:in", "launch:out"),
           ("tile + walk graph until the merge call", "launch:out", "merge:in"), ("merge call", "merge:in", "merge:out"),
           ("until D2H call", "merge:out", "d2h:in"), ("D2H + wait for the GPU", "d2h:in", "d2h:out"),
           ("until assemble", "d2h:out", "assemble:in"), ("assemble pandas frame", "assemble:in", "assemble:out"),
           ("rest of execute()", "assemble:out", "executed"), ("fetch", "executed", "t1"), ("TOTAL", "t0", "t1")]
    for name, a, b in seq:
        acc[name] = acc.get(name, 0.0) + m[b] - m[a]
for name, v in acc.items():
    print(f"{name:45s} {1e3 * v / N:8.3f} ms")
