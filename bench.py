#!/usr/bin/env python
"""bench.py -- groupby-agg rows/s on the TPC-H-Q1-shaped config (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over the whole batch: the Q1-shaped lineitem frame (SF10 cardinality,
59,986,052 rows per GPU, chunk_size 2^22 -> 15 row chunks) goes through
`df.groupby(['returnflag','linestatus']).agg({...q01...})` tiled exactly like the reference tiles it
(DataFrameGroupByAgg map per chunk -> concat/combine tree -> agg) and every chunk operand is executed through
the registered `fn(ctx, op)` executors, which call libmarsgb.so through the C ABI.

  value   rows/s with the chunks already resident in HBM (DeviceFrame data source)
  e2e     rows/s with HOST (pinned) chunks: per step H2D of every chunk (DataFrameToGPU) and D2H of the result
          frame are inside the timed region
  roofline  dominant kernel (map-side pre-aggregation): algorithmic bytes (64 B/row) / its CUDA-event time
  cpu_baseline  the oracle port of the reference's pandas path on a bounded sample, all host cores

N > 1 (torchrun): weak scaling -- every rank owns its own 15 chunks (chunk c of rank r is global chunk
15 r + c); map + local tree are independent per rank; the per-rank partial tuples (4 groups) are gathered to
rank 0 with the NCCL exchange and the final agg stage runs there.  No other data-path collective.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240229
ROWS_PER_GPU = 59_986_052
CHUNK_ROWS = 1 << 22
BY = ["returnflag", "linestatus"]
COLUMNS = ["returnflag", "linestatus", "quantity", "extendedprice", "discount", "disc_price", "charge", "orderkey"]
Q1_FUNC = {"quantity": ["sum", "mean"], "extendedprice": ["sum", "mean"], "disc_price": ["sum"],
           "charge": ["sum"], "discount": ["mean"], "orderkey": ["count"]}
BYTES_PER_ROW = 8 * len(COLUMNS)          # every input column read once: 64 B/row
L2_BYTES = 126 << 20


sys.path.insert(0, os.path.join(ROOT, "tools"))
import baseline_configs as bc  # noqa: E402

gen_chunk = bc.CONFIGS["cfg2"].gen      # Q1-shaped synthetic lineitem chunk


def chunk_sizes(total_rows: int, chunk_rows: int):
    return [min(chunk_rows, total_rows - s) for s in range(0, total_rows, chunk_rows)]


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference's OWN operands (unmodified mars-project/mars built into baseline/_ref by
# oracle/build_ref.sh, driven by oracle/ref_runner.ReferenceTreeRun: its tiling, its compiled reduction steps,
# its pandas-backed execute(ctx, op); stage=map operands fanned out over all host cores, concat / combine / agg
# in the parent).  Falls back to the oracle PORT (same pandas arithmetic at the same call sites) only when
# baseline/_ref is absent, and says which in cpu_baseline.kind.
# ----------------------------------------------------------------------------------------------------
def reference_build_available() -> bool:
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "mars"))


def _host_frame(rows: int, chunk_rows: int):
    """the Q1-shaped frame of `rows` rows, chunk c generated with default_rng(SEED + c) -- the same bytes the
    GPU arm sees -- written straight into full-length columns (no concat copy)"""
    import pandas as pd

    cols = {name: None for name in COLUMNS}
    start = 0
    for c, n in enumerate(chunk_sizes(rows, chunk_rows)):
        chunk = gen_chunk(c, n)
        for name in COLUMNS:
            if cols[name] is None:
                cols[name] = np.empty(rows, dtype=chunk[name].dtype)
            cols[name][start:start + n] = chunk[name]
        start += n
    return pd.DataFrame(cols, copy=False)


_PORT_CHUNKS = None


def _port_map(i):
    import mars_groupby_oracle as orc

    return orc.agg_map(_PORT_CHUNKS[i], BY, Q1_FUNC)


class _PortTreeRun:
    """same structure as ReferenceTreeRun on the oracle's restatement (used only without baseline/_ref)"""

    def __init__(self, raw, chunk_rows, cores):
        global _PORT_CHUNKS
        import multiprocessing as mp

        _PORT_CHUNKS = [raw.iloc[s:s + chunk_rows] for s in range(0, len(raw), chunk_rows)]
        self.n = len(_PORT_CHUNKS)
        self.pool = mp.get_context("fork").Pool(cores)

    def step(self):
        import mars_groupby_oracle as orc

        level = self.pool.map(_port_map, range(self.n), chunksize=1)
        while len(level) > 4:
            level = [orc.agg_combine(orc.concat_partials(level[i:i + 4]), Q1_FUNC) if len(level[i:i + 4]) > 1
                     else orc.agg_combine(level[i], Q1_FUNC) for i in range(0, len(level), 4)]
        cat = level[0] if len(level) == 1 else orc.concat_partials(level)
        return orc.agg_agg(cat, Q1_FUNC, sort=True)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_path(rows: int, chunk_rows: int, steps: int, warmup: int, cores: int):
    """-> (rows/s, seconds per timed step, kind, result frame of the last step)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    raw = _host_frame(rows, chunk_rows)
    if reference_build_available():
        import ref_runner

        run, kind = ref_runner.ReferenceTreeRun(raw, BY, Q1_FUNC, chunk_rows, cores, method="tree", sort=True), "reference"
    else:
        run, kind = _PortTreeRun(raw, chunk_rows, cores), "port"
    times, res = [], None
    try:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            res = run.step()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    finally:
        run.close()
    assert res is not None and len(res) >= 1
    return rows * len(times) / sum(times), times, kind, res


def cpu_baseline_leg(args, n_chunks: int = 8, steps: int = 4, warmup: int = 1):
    """the reference arm on a bounded sample (n_chunks full chunks), run as `bench.py --impl reference` in a
    child process; returns the cpu_baseline object of the bench line"""
    rows = min(args.rows, n_chunks * args.chunk_rows)
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", str(warmup),
           "--rows", str(rows), "--chunk-rows", str(args.chunk_rows)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    t0 = time.perf_counter()
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
        cb = line["cpu_baseline"]
        cb["sample"] = (f"{rows} rows ({n_chunks} chunks) per step, {steps} steps after {warmup} warm-up: " + cb["sample"]
                        + f"; {time.perf_counter() - t0:.1f} s in total incl. input generation")
        return cb
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "rows/s", "cores": os.cpu_count(), "kind": "unavailable",
                "sample": f"cpu baseline leg failed: {e!r}"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    rps, times, kind, res = cpu_path(args.rows, args.chunk_rows, steps, warmup, cores)
    n_chunks = len(chunk_sizes(args.rows, args.chunk_rows))
    what = ("unmodified reference operands (baseline/_ref, oracle/ref_runner.ReferenceTreeRun)" if kind == "reference"
            else "oracle port of the reference operands (baseline/_ref absent)")
    sample = (f"{args.rows} rows = {n_chunks} chunks of <= {args.chunk_rows} rows of the Q1-shaped workload per step "
              f"(one GPU's share of the job), {what}, stage=map operands on {cores} processes")
    cfg = workload_config(args, 1)
    line = {
        "impl": "reference", "metric": "groupby-agg rows/s", "value": rps, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": rps, "unit": "rows/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rps, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "result_groups": int(len(res)),
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "TPC-H Q1-shaped synthetic lineitem (BASELINE.json configs[1]): groupby(returnflag, "
                        "linestatus) sum/mean/count, 2 int64 keys + 5 float64 + 1 int64 value columns",
            "rows_per_gpu": args.rows, "chunk_rows": args.chunk_rows, "chunks_per_gpu": len(chunk_sizes(args.rows, args.chunk_rows)),
            "groups": 4, "method": "tree (auto)", "bytes_per_row": BYTES_PER_ROW,
            "l2": "inputs larger than L2 (%.2f GB per step per GPU)" % (args.rows * BYTES_PER_ROW / 1e9),
            "parallelism": f"dp{world} (row chunks sharded, weak)"}


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons under the bench's load, sampled by a separate `nvidia-smi -lms` process (the recipe
    of /opt/skills/guides/B200_PROFILING.md: start before, kill after).  In-process NVML queries were measured to
    slow the step down by up to 2x (they contend with the kernel launches for the driver), so nothing inside this
    process touches NVML: the sampler is started and given time to initialise BEFORE the warm-up steps, it keeps
    running through the timed steps, and the same step keeps running for a short window afterwards so that at
    least two samples are taken under the identical, continuous load.  Samples are time-stamped; only those taken
    between the first warm-up step and the end of that window are used."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int, period_ms: int = 200):
        self.idx = device_index
        self.period_ms = period_ms
        self.proc, self.path = None, None
        self.t_load0 = self.t_load1 = None

    def _physical_index(self):
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            parts = visible.split(",")
            if self.idx < len(parts) and parts[self.idx].strip().isdigit():
                return int(parts[self.idx])
        return self.idx

    def start(self, wait_s: float = 4.0):
        try:
            fd, self.path = tempfile.mkstemp(prefix="mgb_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self._physical_index()),
                 "-lms", str(self.period_ms), "-f", self.path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            t_end = time.time() + wait_s
            while time.time() < t_end and os.path.getsize(self.path) == 0:     # its start-up is over once it has written
                time.sleep(0.02)
        except Exception:  # noqa: BLE001
            self.proc = None

    def load_begins(self):
        self.t_load0 = time.time()

    def hold_load(self, run_step, n_steps: int):
        """keep the same load running (untimed; a step count that every rank computes alike) so that the sampler
        sees it at least twice"""
        for _ in range(n_steps):
            run_step()
        self.t_load1 = time.time()

    def stop(self):
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        samples, mx, reasons, total = [], None, set(), 0
        if self.proc is not None:
            try:
                self.proc.terminate()
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                pass
            try:
                names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                for line in open(self.path):
                    f = [x.strip() for x in line.split(",")]
                    if len(f) < 7:
                        continue
                    total += 1
                    try:
                        ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                        mhz = float(f[1])
                    except ValueError:
                        continue
                    slack = self.period_ms / 1000.0
                    if self.t_load0 is not None and not (self.t_load0 <= ts <= (self.t_load1 or time.time()) + slack):
                        continue
                    samples.append(mhz)
                    mx = float(f[2])
                    reasons.update(n for n, v in zip(names, f[3:7]) if v.lower().startswith("active"))
                os.unlink(self.path)
            except Exception:  # noqa: BLE001
                pass
        if samples:
            out.update(sm_mhz=statistics.median(samples), sm_max_mhz=mx, samples=len(samples), reasons=sorted(reasons),
                       source=f"nvidia-smi -lms {self.period_ms} in a separate process; samples taken between the first warm-up "
                              f"step and the end of a {0.5} s window of the same step after the timed region "
                              f"({total} samples in the file)")
        return out


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
# bench sizes of the other BASELINE.json configs (sub-records of the N=1 line): exact chunk shape, fewer chunks
CONFIG_BENCH = {
    "cfg1": (10_000_000, 1_000_000),         # full size, default sort=True / method='auto' (tree)
    "cfg1s": (10_000_000, 1_000_000),        # same data, sort=False + method='shuffle' (the named hash path, R = 10)
    "cfg3": (8 * 15_625_000, 15_625_000),    # 1/8 of the rows, R = 8
    "cfg4": (8 * 7_812_500, 7_812_500),
    "cfg5": (8 * 15_625_000, 15_625_000),
}


def main_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (rank 0, N=1 only), in a process of its own: it imports the reference behind its
    # compatibility shim and forks worker processes -- neither belongs in the process that drives CUDA
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(args)

    import pandas as pd
    import torch
    import torch.distributed as dist

    import mars_b200 as md
    from mars_b200 import _lib, kernels
    from mars_b200.device import DeviceFrame

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from mars_b200.parallel import NcclTransport

        comm = NcclTransport(rank, world)       # libmarsgb's own communicator (dlopen()ed NCCL)
    lib = _lib.load()

    sizes = chunk_sizes(args.rows, args.chunk_rows)
    n_chunks = len(sizes)
    rows_rank = sum(sizes)

    # host chunks: pinned (one flat pinned tensor per chunk, columns are views) and pageable pandas frames
    host_frames, host_pageable, dev_frames = [], [], []
    for c, rows in enumerate(sizes):
        cols = gen_chunk(rank * n_chunks + c, rows)
        flat = torch.empty((len(COLUMNS), rows), dtype=torch.int64).pin_memory()
        data = {}
        for j, name in enumerate(COLUMNS):
            view = flat[j].numpy().view(cols[name].dtype)
            view[:] = cols[name]
            data[name] = view
        host_frames.append(pd.DataFrame(data, copy=False))
        host_pageable.append(pd.DataFrame(cols, copy=False))
        dev_frames.append(DeviceFrame([], None, COLUMNS, [
            flat[j].to(dev, non_blocking=True).view(torch.float64 if cols[n_].dtype == np.float64 else torch.int64)
            for j, n_ in enumerate(COLUMNS)], rows))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(frames, fuse=True):
        mdf = md.DataFrame.from_chunks(frames)
        if frames is not dev_frames:
            mdf = mdf.to_gpu()
        sess = md.new_session(default=False)
        sess.fuse = fuse
        r = mdf.groupby(BY).agg(Q1_FUNC, method="auto")
        if world == 1:
            r.execute(session=sess)
            out = r.fetch(session=sess)
        else:
            out = sess.execute_sharded_tree(r, comm=comm)
        return out

    step_ms = {}

    def timed(frames, steps, warmup, sampler=None, fuse=True):
        for _ in range(warmup):
            step(frames, fuse)
        # no collector pause inside the timed region: what the earlier phases of this process left behind (host
        # frames, e2e runs) is collected now and parked; the session itself pauses the collector per graph
        gc.collect()
        gc.freeze()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        per_step = []
        for i in range(steps):
            ts = time.perf_counter()
            out = step(frames, fuse)
            per_step.append((time.perf_counter() - ts) * 1e3)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        step_ms[id(frames), fuse] = per_step
        ms = e0.elapsed_time(e1)
        # the result frame is fetched to the host inside every step, so device time ~ wall time; use the
        # larger of the two so that host-side work after the last kernel is not hidden
        ms = max(ms, wall * 1e3)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # ---- e2e (host chunks: pinned, then pageable), then the device-resident headline with clocks + per-kernel timing
    from mars_b200 import device as _dev

    e2e_steps = max(1, min(args.steps, 3))
    step(host_frames)                                   # warm-up (allocator, plans)
    h2d0, nar0, skip0 = _dev.h2d_bytes, _dev.narrowed_columns, _dev.skipped_columns
    ms_e2e, out_e2e = timed(host_frames, e2e_steps, 0)
    h2d_bytes = (_dev.h2d_bytes - h2d0) // e2e_steps   # bytes really copied per step (counted at the copy call)
    e2e_narrowed = (_dev.narrowed_columns - nar0) / e2e_steps, (_dev.skipped_columns - skip0) / e2e_steps
    ms_page = None
    e2e_variants = {}
    if world == 1:
        step(host_pageable)
        ms_page, _ = timed(host_pageable, 2, 0)
        ms_page /= 2
        # the same pinned step with less of the narrow transfer encoding (device._to_device)
        saved = _dev.NARROW, _dev.NARROW_FLOATS, _dev.NARROW_DEFER, _dev.NARROW_OPPORTUNISTIC
        for name, flags, note in (
                ("all_deferred", (True, True, True, True), "int64 columns and integer-valued float64 columns cross PCIe narrow; "
                 "the host pass runs in the background of the band's other uploads; a pass that has not started when the "
                 "copy engine runs dry is skipped (that column crosses at 8 bytes)"),
                ("all_deferred_wait", (True, True, True, False), "MGB_NARROW_OPPORTUNISTIC=0: every host pass is waited for"),
                ("all_sync", (True, True, False, False), "MGB_NARROW_DEFER=0: the host pass of a column completes before the "
                 "next column is requested"),
                ("ints_deferred", (True, False, True, True), "MGB_NARROW_FLOATS=0: only int64 columns cross PCIe narrow"),
                ("ints_sync", (True, False, False, False), "MGB_NARROW_FLOATS=0 MGB_NARROW_DEFER=0"),
                ("full_width", (False, False, False, False), "MGB_NARROW=0: every uploaded column crosses PCIe at 8 bytes per "
                 "element")):
            if flags == saved:
                continue
            _dev.NARROW, _dev.NARROW_FLOATS, _dev.NARROW_DEFER, _dev.NARROW_OPPORTUNISTIC = flags
            _dev._no_narrow.clear()
            _dev._narrow_hint.clear()
            try:
                step(host_frames)
                h2d0, nar0, skip0 = _dev.h2d_bytes, _dev.narrowed_columns, _dev.skipped_columns
                ms_v, _ = timed(host_frames, e2e_steps, 0)
                e2e_variants[name] = {"value": rows_rank * e2e_steps / (ms_v * 1e-3), "ms_per_step": ms_v / e2e_steps,
                                      "h2d_bytes_per_step": int((_dev.h2d_bytes - h2d0) // e2e_steps),
                                      "narrowed_columns_per_step": (_dev.narrowed_columns - nar0) / e2e_steps,
                                      "skipped_columns_per_step": (_dev.skipped_columns - skip0) / e2e_steps, "note": note}
            finally:
                _dev.NARROW, _dev.NARROW_FLOATS, _dev.NARROW_DEFER, _dev.NARROW_OPPORTUNISTIC = saved
                _dev._no_narrow.clear()
                _dev._narrow_hint.clear()
    del host_pageable

    sampler = ClockSampler(local)
    if not args.no_clocks:
        sampler.start()
    lib.mgb_prof_enable(1)
    sampler.load_begins()
    for _ in range(args.warmup):
        step(dev_frames)
    barrier()
    lib.mgb_prof_reset()
    launches0 = kernels.launch_counter
    ms, out = timed(dev_frames, args.steps, 0)
    launches = kernels.launch_counter - launches0
    import ctypes as C
    prof = {}
    for kid, name in enumerate(["preagg_dense", "preagg_smem", "merge", "update_global", "table_init", "emit", "hash",
                                "partition", "sort", "post"]):
        tot, cnt = C.c_double(), C.c_int64()
        lib.mgb_prof_read(kid, C.byref(tot), C.byref(cnt))
        if cnt.value:
            prof[name] = {"ms_total": tot.value, "launches": cnt.value}
    lib.mgb_prof_enable(0)
    if not args.no_clocks:
        sampler.hold_load(lambda: step(dev_frames), min(2000, max(5, int(500.0 / max(ms / args.steps, 0.05)))))
    clocks = sampler.stop()

    # the same step without the band-level batch: every chunk operand on its own, what a real Mars worker does
    po_steps = max(2, min(args.steps, 5))
    ms_po, out_po = timed(dev_frames, po_steps, 2, fuse=False)

    parity = None
    if rank == 0 and world == 1 and cpu_baseline and cpu_baseline.get("kind") != "unavailable":
        # the result of this run against the CPU arm's on the SAME chunks (the cpu leg's bounded sample)
        try:
            n_cpu = min(8, n_chunks)
            sess = md.new_session(default=False)
            r = md.DataFrame.from_chunks(dev_frames[:n_cpu]).groupby(BY).agg(Q1_FUNC, method="auto")
            got = r.execute(session=sess).fetch(session=sess)
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import mars_groupby_oracle as orc

            raw = pd.concat(host_frames[:n_cpu], axis=0, ignore_index=True)
            exp = orc.run_groupby_agg(raw, BY, Q1_FUNC, args.chunk_rows, method="tree", sort=True)
            ok = list(got.columns) == list(exp.columns) and got.index.equals(exp.index)
            for c in exp.columns:
                g, e = got[c].to_numpy(), exp[c].to_numpy()
                ok = ok and (np.array_equal(g, e) if c[1] == "count" else np.allclose(g, e, rtol=1e-9, atol=0))
            parity = {"ok": bool(ok), "mode": f"full frame vs the oracle on the first {n_cpu} chunks (counts bit-exact, float64 1e-9)"}
        except Exception as e:  # noqa: BLE001
            parity = {"ok": False, "mode": f"check failed to run: {e!r}"}

    dev_step_ms = step_ms[id(dev_frames), True]
    # N > 1: the hash-shuffle branch across the ranks (cfg3-shaped, weak scaling: 125 M rows per GPU = the config's
    # full size at 8 GPUs) -- whole-path rows/s, NVLink rate of the fused partition -> peer transfer, parity
    shuffle = None
    shuffle_configs = {}
    if world > 1 and not args.no_shuffle:
        del host_frames
        torch.cuda.empty_cache()
        from mars_b200 import parallel
        import shuffle_leg

        ex = None
        try:
            ex = parallel.init_exchange()
            shuffle = shuffle_leg.run(ex, "cfg3", rows_per_gpu=args.shuffle_rows, chunk_rows=15_625_000,
                                      steps=max(2, min(args.steps, 3)), warmup=2)
        except Exception as e:  # noqa: BLE001
            import traceback

            shuffle = {"error": repr(e)[:300], "trace": traceback.format_exc()[-600:]}
        # the other two shuffle-bound BASELINE configs at their per-GPU share of the full size (weak scaling: 8 GPUs
        # hold the whole config): cfg4 62.5 M rows per GPU, cfg5 125 M rows per GPU
        if ex is not None and not (shuffle or {}).get("error") and not args.no_shuffle_configs:
            for name, rows_gpu, chunk in (("cfg4", 62_500_000, 7_812_500), ("cfg5", 125_000_000, 15_625_000)):
                try:
                    rec = shuffle_leg.run(ex, name, rows_per_gpu=min(rows_gpu, args.shuffle_rows), chunk_rows=chunk, steps=2,
                                          warmup=2)
                except Exception as e:  # noqa: BLE001
                    rec = {"error": repr(e)[:300]}
                if rank == 0:
                    shuffle_configs[name] = rec
        if ex is not None:
            ex.transport.close()
        host_frames = []
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        del host_frames, dev_frames
        torch.cuda.empty_cache()
        import config_runner

        configs = {}
        for name, (rows, chunk_rows) in CONFIG_BENCH.items():
            try:
                configs[name] = config_runner.run_record(name, rows, chunk_rows, repeats=2)
            except Exception as e:  # noqa: BLE001
                configs[name] = {"config": name, "error": repr(e)[:300]}

    if rank == 0:
        # sanity: the paths agree and the counts add up
        assert out.shape == out_e2e.shape == out_po.shape
        total_rows = rows_rank * world
        cnt_col = ("orderkey", "count")
        assert int(out[cnt_col].sum()) == total_rows, (int(out[cnt_col].sum()), total_rows)
        assert int(out_po[cnt_col].sum()) == total_rows
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        dom = prof.get("preagg_dense") or prof.get("preagg_smem") or {"ms_total": float("nan"), "launches": 1}
        per_launch_ms = dom["ms_total"] / dom["launches"]
        rows_per_launch = rows_rank * args.steps / dom["launches"]
        achieved = rows_per_launch * BYTES_PER_ROW / (per_launch_ms * 1e-3) / 1e9
        traffic, traffic_per_row = None, None   # dram__bytes_read.sum + dram__bytes_write.sum per launch (committed ncu capture)
        tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic_per_row = json.load(open(tpath))["dram_bytes_per_row"]
                traffic = traffic_per_row * rows_per_launch
            except Exception:  # noqa: BLE001
                traffic = None
        value = total_rows * args.steps / (ms * 1e-3)
        e2e_value = total_rows * e2e_steps / (ms_e2e * 1e-3)
        line = {
            "metric": "groupby-agg rows/s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": int(h2d_bytes) * world,
                    "h2d_note": "columns are uploaded on first use: the int64 column that is only counted never crosses "
                                "PCIe (chunk bytes in host memory: %d per step per GPU)" % (rows_rank * BYTES_PER_ROW)
                                + ("; int64 columns%s whose values fit 1 / 2 / 4 bytes (here: the two key columns%s) are "
                                   "narrowed on the host inside the timed region (checked on every value: lossless or not "
                                   "done), cross PCIe narrow and are restored to 8-byte columns in HBM"
                                   % ((" and float64 columns holding integers", " and quantity") if _dev.NARROW_FLOATS
                                      else ("", "")) if _dev.NARROW else ""),
                    "transfer_encoding": {"narrow_ints": bool(_dev.NARROW), "narrow_integer_valued_floats":
                                          bool(_dev.NARROW and _dev.NARROW_FLOATS),
                                          "deferred": bool(_dev.NARROW and _dev.NARROW_DEFER),
                                          "narrowed_columns_per_step": e2e_narrowed[0],
                                          "skipped_columns_per_step": e2e_narrowed[1],
                                          "variants": e2e_variants or None},
                    "d2h_bytes_per_step": int(out.shape[0] * (out.shape[1] + len(BY)) * 8),
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "host_memory": "pinned",
                    "pageable": None if ms_page is None else {"value": total_rows / (ms_page * 1e-3), "ms_per_step": ms_page,
                                                               "note": "same step from pageable pandas chunks (what a Mars chunk store hands over)"}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": ("k_preagg_tiny (<= 4 groups per CTA in registers, TMA-staged tiles) + the immediate exit of "
                                    "k_preagg_dense enqueued behind it" if "preagg_dense" in prof else "k_preagg_hash/k_preagg_smem"),
                         "kernel_ms_per_launch": per_launch_ms, "bytes_per_launch": rows_per_launch * BYTES_PER_ROW,
                         "launches": dom["launches"],
                         "frac_on_traffic": None if not traffic_per_row else achieved * traffic_per_row / BYTES_PER_ROW / peak,
                         "note": "achieved counts the 64 algorithmic B/row (SURVEY 8d); the kernel does not read the count-only "
                                 "int64 column, frac_on_traffic counts the DRAM bytes it really moves (ncu capture)",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                         "step_hbm_frac": (rows_rank * BYTES_PER_ROW / (ms / args.steps * 1e-3) / 1e9) / peak},
            "kernel_time_ms": prof,
            "step_ms": {"min": min(dev_step_ms), "median": statistics.median(dev_step_ms), "max": max(dev_step_ms),
                        "note": "host wall time of every timed step; value uses the whole region (device events / wall, max)"},
            "per_operand": {"value": total_rows * po_steps / (ms_po * 1e-3), "unit": "rows/s", "ms_per_step": ms_po / po_steps,
                            "steps": po_steps,
                            "note": "same step with the band-level batch off (MGB_FUSE=0): every chunk operand executed on "
                                    "its own through execute(ctx, op), as under a real Mars worker"},
            "parity": parity,
            "clocks": clocks,
            "cpu_baseline": cpu_baseline,
            "configs": configs,
            "shuffle": shuffle,
            "shuffle_configs": shuffle_configs or None,
        }
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="rows per GPU (default: SF10 lineitem cardinality)")
    ap.add_argument("--chunk-rows", type=int, default=CHUNK_ROWS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE.json configs")
    ap.add_argument("--no-shuffle", action="store_true", help="N > 1: skip the hash-shuffle leg")
    ap.add_argument("--no-shuffle-configs", action="store_true", help="N > 1: only the cfg3 shuffle leg, not cfg4 / cfg5")
    ap.add_argument("--shuffle-rows", type=int, default=125_000_000, help="rows per GPU of the shuffle leg (cfg3-shaped)")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostics: do not start the nvidia-smi clock sampler")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
