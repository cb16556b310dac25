"""GPU timeline of one config run (CUPTI through torch.profiler; nsys is not in the image):
python profiles/timeline.py cfg3 0.25 [trace.json]

Prints, for the profiled run: wall time, GPU busy time (union of kernel / memcpy / memset intervals), the idle
share, device time per kernel / copy kind, and the host-side CUDA API calls by total time (allocation and
synchronisation calls show up here)."""
import collections
import json
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import run_configs  # noqa: E402
import mars_b200 as md  # noqa: E402
from mars_b200.device import DeviceFrame  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
trace_path = sys.argv[3] if len(sys.argv) > 3 else "/tmp/timeline_trace.json"
n, chunk, mk, by, func, kw = run_configs.gen(cfg, scale)
frames = []
for s in range(0, n, chunk):
    cols = mk(min(chunk, n - s))
    frames.append(DeviceFrame([], None, list(cols.keys()), list(cols.values()), min(chunk, n - s)))
torch.cuda.synchronize()


def once():
    sess = md.new_session(default=False)
    r = md.DataFrame.from_chunks(frames).groupby(by, sort=kw["sort"]).agg(func, method=kw["method"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r.execute(session=sess)
    out = r.fetch(session=sess)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, len(out)


for _ in range(2):
    plain, groups = once()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    wall, groups = once()
prof.export_chrome_trace(trace_path)
ev = json.load(open(trace_path))["traceEvents"]
gpu = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
api = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("cuda_runtime", "cuda_driver")]
gpu.sort(key=lambda e: e["ts"])
busy, end = 0.0, -1.0
for e in gpu:
    s, t = e["ts"], e["ts"] + e["dur"]
    if s > end:
        busy += t - s
        end = t
    elif t > end:
        busy += t - end
        end = t
span = (max(e["ts"] + e["dur"] for e in gpu) - gpu[0]["ts"]) if gpu else 0.0
print(f"{cfg} scale {scale}: {n} rows, {groups} groups; unprofiled {1e3 * plain:.1f} ms, profiled {1e3 * wall:.1f} ms")
print(f"GPU span {span / 1e3:.1f} ms, busy {busy / 1e3:.1f} ms ({100 * busy / max(span, 1):.0f} %)")
by_name = collections.defaultdict(lambda: [0.0, 0])
for e in gpu:
    name = e["name"]
    if e["cat"] == "kernel":
        name = name.split("(")[0].split("<")[0]
    by_name[(e["cat"], name)][0] += e["dur"]
    by_name[(e["cat"], name)][1] += 1
print("device time by kernel / copy:")
for (cat, name), (d, c) in sorted(by_name.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  {d / 1e3:9.2f} ms {c:6d} x  {cat:10s} {name[:90]}")
by_api = collections.defaultdict(lambda: [0.0, 0])
for e in api:
    by_api[e["name"]][0] += e["dur"]
    by_api[e["name"]][1] += 1
print("host CUDA API time:")
for name, (d, c) in sorted(by_api.items(), key=lambda kv: -kv[1][0])[:15]:
    print(f"  {d / 1e3:9.2f} ms {c:6d} x  {name}")
if os.environ.get("PYPROF"):
    import cProfile
    import pstats

    pr = cProfile.Profile()
    pr.enable()
    once()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(30)
