"""BASELINE config 1 (10 M rows, 1000 keys, 4 float64 columns, sum / mean / count) through the band-level mid-cardinality
batch: kernel time by CUDA events; run under `ncu -k regex:k_preagg_mid` for the capture.  python profiles/prof_mid.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import baseline_configs as bc  # noqa: E402
from mars_b200 import kernels as K  # noqa: E402

cfg = bc.CONFIGS["cfg1"]
chunks = bc.host_chunks(cfg, cfg.rows, cfg.chunk_rows)
dev = torch.device("cuda", 0)
cols = list(chunks[0])
tens = [[torch.from_numpy(c[k]).to(dev) for k in cols] for c in chunks]
accs = tuple([(i, K.SUM) for i in range(4)] + [(i, K.COUNT) for i in range(4)])
arg = [([t[0].data_ptr()], [x.data_ptr() for x in t[1:]], t[0].numel()) for t in tens]
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bi, bf, fc = K.preagg_mid_batch_async(arg, accs, [K.F64] * 4, 1, dev)
    e1.record()
    torch.cuda.synchronize()
    n = int(bi[fc.n_int, 0])
    print(f"run {it}: {e0.elapsed_time(e1) * 1e3:.1f} us for {cfg.rows} rows -> {n} groups, "
          f"{cfg.rows * 40 / e0.elapsed_time(e1) / 1e6:.0f} GB/s")
