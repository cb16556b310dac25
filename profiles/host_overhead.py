"""Host-side cost breakdown of one dense map call (python profiles/host_overhead.py)."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mars_b200 import _lib, kernels as K  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
g = torch.Generator(device="cuda").manual_seed(1)
keys = [torch.randint(0, 3, (rows,), generator=g, device="cuda"), torch.randint(0, 2, (rows,), generator=g, device="cuda")]
vals = [torch.rand(rows, generator=g, device="cuda", dtype=torch.float64) for _ in range(5)]
accs = [(c, K.SUM) for c in range(5)] + [(0, K.COUNT), (1, K.COUNT), (2, K.COUNT)]
lib = _lib.load()
for _ in range(3):
    K.preagg_dense(keys, vals, accs)
torch.cuda.synchronize()
N = 50
t0 = time.perf_counter()
for _ in range(N):
    K.preagg_dense(keys, vals, accs)
t1 = time.perf_counter()
print(f"preagg_dense wrapper: {1e6 * (t1 - t0) / N:.1f} us per call")
# raw C call with preallocated outputs
dts = [K.dtype_code(v) for v in vals]
fc = K._fast(2, dts, accs)
spec = fc.spec
cap = C.c_int64()
lib.mgb_dense_capacity(C.byref(cap))
blocks = fc.alloc(cap.value, "cuda")
kp, vp, okp, oap = K._ptrs(keys), K._ptrs(vals), fc.okp, fc.oap
gq = C.c_int64()
st = torch.cuda.current_stream().cuda_stream
t0 = time.perf_counter()
for _ in range(N):
    lib.mgb_preagg_dense(C.byref(spec), kp, vp, rows, okp, oap, cap.value, C.byref(gq), st)
t1 = time.perf_counter()
print(f"raw C call (launch + 8-byte D2H + sync): {1e6 * (t1 - t0) / N:.1f} us per call, groups={gq.value}")
lib.mgb_prof_enable(1)
lib.mgb_prof_reset()
for _ in range(N):
    lib.mgb_preagg_dense(C.byref(spec), kp, vp, rows, okp, oap, cap.value, C.byref(gq), st)
tot, cnt = C.c_double(), C.c_int64()
lib.mgb_prof_read(0, C.byref(tot), C.byref(cnt))
print(f"kernel by CUDA events: {1e3 * tot.value / cnt.value:.1f} us per launch ({cnt.value} launches); "
      f"{rows * 56 / (tot.value / cnt.value * 1e-3) / 1e9:.0f} GB/s of the 56 B/row actually read, "
      f"{rows * 64 / (tot.value / cnt.value * 1e-3) / 1e9:.0f} GB/s algorithmic (64 B/row)")
t0 = time.perf_counter()
for _ in range(N):
    fc.alloc(cap.value, "cuda")
t1 = time.perf_counter()
print(f"output block allocation: {1e6 * (t1 - t0) / N:.1f} us")
t0 = time.perf_counter()
for _ in range(N):
    torch.cuda.current_stream().synchronize()
t1 = time.perf_counter()
print(f"idle stream sync: {1e6 * (t1 - t0) / N:.1f} us")
