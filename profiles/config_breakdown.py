"""Per-operand-class wall time + per-kernel device time of one config run:
python profiles/config_breakdown.py cfg3 0.1"""
import collections
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import run_configs  # noqa: E402
import mars_b200 as md  # noqa: E402
from mars_b200 import _lib, core  # noqa: E402
from mars_b200.device import DeviceFrame  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
n, chunk, mk, by, func, kw = run_configs.gen(cfg, scale)
frames = []
for s in range(0, n, chunk):
    cols = mk(min(chunk, n - s))
    frames.append(DeviceFrame([], None, list(cols.keys()), list(cols.values()), min(chunk, n - s)))
torch.cuda.synchronize()
lib = _lib.load()
acc = collections.defaultdict(float)
cnt = collections.Counter()
orig = core.execute


def timed_execute(results, op):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        return orig(results, op)
    finally:
        torch.cuda.synchronize()
        k = f"{type(op).__name__}:{getattr(op.stage, 'name', None)}"
        acc[k] += time.perf_counter() - t0
        cnt[k] += 1


for it in range(2):
    if it == 1:
        core.execute = timed_execute
        lib.mgb_prof_enable(1)
        lib.mgb_prof_reset()
    sess = md.new_session(default=False)
    r = md.DataFrame.from_chunks(frames).groupby(by, sort=kw["sort"]).agg(func, method=kw["method"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r.execute(session=sess)
    out = r.fetch(session=sess)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
print(f"{cfg} scale {scale}: {n} rows, {len(out)} groups, {1e3 * total:.1f} ms (synchronised per operand)")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:40s} {cnt[k]:3d} calls {1e3 * v:9.2f} ms total {1e3 * v / cnt[k]:8.3f} ms each")
NAMES = ["preagg_dense", "preagg_smem/hash", "merge", "update_global", "table_init", "emit", "hash", "partition", "sort", "post"]
for kid, name in enumerate(NAMES):
    tot, c = C.c_double(), C.c_int64()
    lib.mgb_prof_read(kid, C.byref(tot), C.byref(c))
    if c.value:
        print(f"    kernel {name:18s} {tot.value:9.2f} ms / {c.value} launches")
if os.environ.get("PROFILE_STAGE"):
    import cProfile
    import pstats

    core.execute = orig
    lib.mgb_prof_enable(0)
    stage = os.environ["PROFILE_STAGE"]
    prof = cProfile.Profile()

    def prof_execute(results, op):
        on = getattr(op.stage, "name", None) == stage and type(op).__name__ == "DataFrameGroupByAgg"
        if on:
            prof.enable()
        try:
            return orig(results, op)
        finally:
            if on:
                torch.cuda.synchronize()
                prof.disable()

    core.execute = prof_execute
    sess = md.new_session(default=False)
    r = md.DataFrame.from_chunks(frames).groupby(by, sort=kw["sort"]).agg(func, method=kw["method"])
    r.execute(session=sess)
    pstats.Stats(prof).sort_stats("tottime").print_stats(18)
