#!/usr/bin/env python
"""NVLink rate of the partition kernel's stores into peer arenas (mgb_split_scatter with a destination table), isolated:
every rank splits one shuffle-shaped partial (rows x cols of 8 bytes) into R blocks, ALL of which live in the arena of
rank (rank + 1) % N ("ring") or are spread over all peers ("all").   torchrun --nproc-per-node N tools/push_microbench.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mars_b200 import kernels as K, parallel  # noqa: E402

rows = int(os.environ.get("ROWS", 15_625_000))
LONG = K.SPLIT_LONG_RUNS if os.environ.get("LONG_RUNS", "1") != "0" else 0    # LONG_RUNS=0: 2048-row tiles
ex = parallel.init_exchange("gloo")
t = ex.transport
rank, world = ex.rank, ex.world
assert getattr(t, "push", False), "needs the push transport"
dev = torch.device("cuda", torch.cuda.current_device())
g = torch.Generator(device="cuda").manual_seed(rank)
out = []
for ncols, R, mode in ((3, 8, "ring"), (3, 64, "ring"), (3, 8, "all"), (3, 64, "all"), (3, 256, "all"), (10, 64, "all"),
                       (3, 64, "local")):
    keys = [torch.randint(0, 1 << 40, (rows,), dtype=torch.int64, device=dev, generator=g)]
    cols = keys + [torch.rand(rows, dtype=torch.float64, device=dev, generator=g) for _ in range(ncols - 1)]
    hist, offsets = K.split_plan(keys, R, LONG)
    off = offsets.cpu().numpy()
    cnt = off[1:] - off[:-1]
    pad = (cnt + 1) & ~1
    # partition p goes to rank dest[p]; in that arena: [source rank][column][partition] with a generous fixed stride
    if mode == "ring":
        dest = np.full(R, (rank + 1) % world)
    elif mode == "local":
        dest = np.full(R, rank)
    else:
        dest = np.arange(R) * world // R
    stride = int(rows // 1 + 2 * R + 64) & ~1            # elements per (source, column) segment: room for every block
    need = 8 * world * ncols * stride
    slots = [1] * world
    t.ensure(slots, [need] * world)
    start = np.concatenate([[0], np.cumsum(pad)])[:-1]
    tab = np.zeros((ncols, R), dtype=np.int64)
    for c in range(ncols):
        for p in range(R):
            tab[c, p] = t.bases[1][int(dest[p])] + 8 * ((rank * ncols + c) * stride + int(start[p]))
    dtab = torch.from_numpy(tab.reshape(-1)).to(dev)
    remote = int(cnt[dest != rank].sum()) * ncols * 8

    def run():
        K.split_scatter(keys, R, hist, cols, table=dtab, mode=LONG)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    v = torch.tensor([ms], dtype=torch.float64)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    dist.barrier()
    out.append({"cols": ncols, "R": R, "mode": mode, "ms": float(v.item()), "remote_bytes": remote,
                "GBps_out_per_gpu": remote / (float(v.item()) * 1e-3) / 1e9,
                "total_GBps": rows * ncols * 8 / (float(v.item()) * 1e-3) / 1e9})
    del keys, cols, hist, dtab
if rank == 0:
    for o in out:
        print(json.dumps(o), flush=True)
t.close()
dist.destroy_process_group()
