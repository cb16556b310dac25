#!/usr/bin/env python
"""Hash-shuffle branch across N GPUs (one process per GPU, torchrun): parity against pandas + exchange bandwidth.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/shuffle_ngpu.py [--rows 4000000] [--keys 1000000] [--chunks-per-gpu 4] [--check]

Every rank generates the same seeded frame (so the pandas check can run on rank 0), owns the map chunks
c % N == rank, and the shuffle moves partial blocks with the grouped ncclSend/ncclRecv all-to-all of
libmarsgb (mgb_comm_alltoallv).  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mars_b200 as md  # noqa: E402
from mars_b200 import parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--keys", type=int, default=1_000_000)
    ap.add_argument("--chunks-per-gpu", type=int, default=4)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--sort", action="store_true", help="groupby(sort=True): PSRS range shuffle instead of the hash shuffle")
    ap.add_argument("--config", default=None, help="run tools/shuffle_leg.py on this BASELINE config (cfg3 / cfg4 / cfg5) "
                                                   "with --rows rows per GPU instead of the uniform-key frame")
    ap.add_argument("--chunk-rows", type=int, default=15_625_000)
    args = ap.parse_args()
    ex = parallel.init_exchange("gloo")          # gloo: host metadata; bytes move through libmarsgb + NCCL
    if args.config:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import shuffle_leg

        rec = shuffle_leg.run(ex, args.config, rows_per_gpu=args.rows, chunk_rows=args.chunk_rows)
        if ex.rank == 0:
            print(json.dumps(rec), flush=True)
        ex.transport.close()
        dist.destroy_process_group()
        return
    rank, world = ex.rank, ex.world
    ex.transport.timed = True
    rng = np.random.default_rng(11)
    raw = pd.DataFrame({"key": rng.integers(0, args.keys, args.rows), "v": rng.random(args.rows)})
    n_chunks = args.chunks_per_gpu * world
    chunk = -(-args.rows // n_chunks)
    # chunks resident in HBM (each rank only materialises the chunks it owns: c % world == rank)
    from mars_b200.device import DeviceFrame
    pieces = []
    for c in range(n_chunks):
        part = raw.iloc[c * chunk:(c + 1) * chunk]
        if c % world == rank:
            pieces.append(DeviceFrame.from_pandas(part, device=torch.device("cuda", torch.cuda.current_device())))
        else:   # placeholder with the same schema: never executed on this rank
            pieces.append(DeviceFrame.from_pandas(part.iloc[:0], device=torch.device("cuda", torch.cuda.current_device())))
    times = []
    for it in range(4):
        sess = md.new_session(exchange=ex, default=False)
        r = md.DataFrame.from_chunks(pieces).groupby("key", sort=args.sort).agg(["sum", "count"], method="shuffle")
        if it == 1:
            ex.transport._events.clear()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        r.execute(session=sess)
        torch.cuda.synchronize()
        dist.barrier()
        times.append(time.perf_counter() - t0)
    local_groups = torch.tensor([float(sum(len(v) for v in sess._results.values()))], dtype=torch.float64)
    dist.all_reduce(local_groups)
    ok = None
    got = sess.fetch(r) if args.check else None      # gathers every result chunk as pandas: small inputs only
    if args.check and rank == 0:
        exp = raw.groupby("key").agg(["sum", "count"])
        g = got if args.sort else got.sort_index()      # sort=True: the fetched frame must already be ordered
        ok = bool(g.index.equals(exp.index) and np.array_equal(g[("v", "count")], exp[("v", "count")])
                  and np.allclose(g[("v", "sum")], exp[("v", "sum")], rtol=1e-9, atol=0))
    ms, nbytes = ex.transport.exchange_stats()          # runs 1..3 (run 0 = warm-up incl. NCCL connection setup)
    stat = torch.tensor([ms, float(nbytes)], dtype=torch.float64)
    stats = [torch.zeros_like(stat) for _ in range(world)]
    dist.all_gather(stats, stat)
    if rank == 0:
        worst_ms = max(float(s_[0]) for s_ in stats)
        per_gpu_bytes = max(float(s_[1]) for s_ in stats)
        print(json.dumps({"n_gpus": world, "sort": bool(args.sort), "transport": type(ex.transport).__name__, "rows": args.rows, "groups": int(local_groups.item()), "chunks": n_chunks,
                          "parity_vs_pandas": ok, "seconds_per_run": times, "rows_per_s": args.rows / min(times[1:]),
                          "exchange": {"runs": 3, "device_ms_max_over_ranks": worst_ms,
                                       "bytes_sent_per_gpu": per_gpu_bytes,
                                       "GBps_per_gpu_per_direction": per_gpu_bytes / (worst_ms * 1e-3) / 1e9,
                                       "nvlink_reference_GBps": 770, "frac_of_measured_peer_copy": per_gpu_bytes / (worst_ms * 1e-3) / 1e9 / 770},
                          "ops": sorted(set(sess.executed_ops))}), flush=True)
    ex.transport.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
