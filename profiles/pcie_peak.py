"""Pinned host -> device copy bandwidth of this box (what bounds bench.py's e2e): one 1 GiB copy, 32 MiB copies
back to back on one stream (the size of one Q1 column of a 2^22-row chunk), and the same on two streams."""
import time

import torch

dev = torch.device("cuda", 0)
big = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)


def timed(fn, nbytes, reps=8):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return nbytes / best / 1e9


print(f"1 GiB in one copy:            {timed(lambda: dst.copy_(big, non_blocking=True), 1 << 30):6.1f} GB/s")
step = 32 << 20


def chunks():
    for o in range(0, 1 << 30, step):
        dst[o:o + step].copy_(big[o:o + step], non_blocking=True)


print(f"32 MiB copies, one stream:    {timed(chunks, 1 << 30):6.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def two():
    for i, o in enumerate(range(0, 1 << 30, step)):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2):
            dst[o:o + step].copy_(big[o:o + step], non_blocking=True)


print(f"32 MiB copies, two streams:   {timed(two, 1 << 30):6.1f} GB/s")
