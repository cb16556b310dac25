"""Where one bench step spends its host time (round 2: band-level batch).
    python profiles/step_phases.py [rows_per_chunk] [--profile]
Phase timers at the given chunk size (default: the bench's 2^22 rows), and with --profile a cProfile of 200
steps at 65536 rows per chunk (GPU time negligible: what remains is the host path)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mars_b200 as md  # noqa: E402
from mars_b200 import session as S  # noqa: E402
from mars_b200.device import DeviceFrame  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rows = int(args[0]) if args else 1 << 22
method = os.environ.get("METHOD", "auto")


def make(rows):
    frames = []
    for c in range(15):
        cols = bench.gen_chunk(c, rows)
        frames.append(DeviceFrame([], None, bench.COLUMNS, [torch.from_numpy(cols[n]).cuda() for n in bench.COLUMNS], rows))
    torch.cuda.synchronize()
    return frames


marks = []
orig_run = S.Session._run_chunks_local


def run_marked(self, result_chunks, ctx, owners):
    t0 = time.perf_counter()
    r = orig_run(self, result_chunks, ctx, owners)
    marks.append(("run_chunks(%d)" % len(result_chunks), t0, time.perf_counter()))
    return r


S.Session._run_chunks_local = run_marked


def step(frames, fuse=True):
    marks.clear()
    t0 = time.perf_counter()
    mdf = md.DataFrame.from_chunks(frames)
    sess = md.new_session(default=False)
    sess.fuse = fuse
    r = mdf.groupby(bench.BY).agg(bench.Q1_FUNC, method=method)
    t1 = time.perf_counter()
    r.execute(session=sess)
    t2 = time.perf_counter()
    out = r.fetch(session=sess)
    t3 = time.perf_counter()
    return out, (t0, t1, t2, t3, list(marks))


frames = make(rows)
for fuse in (True, False):
    for _ in range(5):
        step(frames, fuse)
    N = 30
    acc = {}
    tot = 0.0
    for _ in range(N):
        torch.cuda.synchronize()
        out, (t0, t1, t2, t3, mk) = step(frames, fuse)
        tot += t3 - t0
        acc["build lazy graph"] = acc.get("build lazy graph", 0) + t1 - t0
        last = t1
        for i, (name, a, b) in enumerate(mk):
            acc[f"tile before {name} #{i}"] = acc.get(f"tile before {name} #{i}", 0) + a - last
            acc[f"{name} #{i}"] = acc.get(f"{name} #{i}", 0) + b - a
            last = b
        acc["after last run_chunks (results)"] = acc.get("after last run_chunks (results)", 0) + t2 - last
        acc["fetch"] = acc.get("fetch", 0) + t3 - t2
    print(f"fuse={fuse} method={method} rows/chunk={rows}: step {1e3 * tot / N:.3f} ms")
    for k, v in acc.items():
        print(f"   {k:44s} {1e3 * v / N:7.3f} ms")

if "--profile" in sys.argv:
    small = make(65536)
    for _ in range(20):
        step(small)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        step(small)
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(45)
