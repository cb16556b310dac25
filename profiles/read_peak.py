"""Read-only HBM rate a plain streaming reduction reaches on this GPU (reference point next to the copy peak of
MEASURED_PEAKS.json: the pre-aggregation kernels read, they hardly write).  python profiles/read_peak.py"""
import json

import torch

x = torch.rand(1 << 29, device="cuda", dtype=torch.float64)      # 4 GiB
for _ in range(3):
    x.sum()
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    x.sum()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
y = torch.empty_like(x)
bestc = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y.copy_(x)
    e1.record()
    torch.cuda.synchronize()
    bestc = min(bestc, e0.elapsed_time(e1))
print(json.dumps({"read_only_sum_GBps": x.numel() * 8 / best / 1e6, "copy_read_plus_write_GBps": 2 * x.numel() * 8 / bestc / 1e6,
                  "bytes": x.numel() * 8}))
