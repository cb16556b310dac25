import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mars_b200 import device as D
dev = torch.device('cuda', 0)
cols = [np.random.default_rng(i).random(4_194_304) for i in range(8)]
for _ in range(2):
    ts = [D._to_device(c, dev) for c in cols]; torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    ts = [D._to_device(c, dev) for c in cols]
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(os.environ.get("MGB_STAGE_THREADS"), os.environ.get("MGB_STAGE_SLABS"), os.environ.get("MGB_STAGE_SLAB_MB"), "GB/s %.1f" % (8 * 4_194_304 * 8 / dt / 1e9))
