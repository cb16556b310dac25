"""Multi-rank host logic on CPU: two gloo ranks run the hash-shuffle branch and the sharded tree branch with
the kernels module replaced by the CPU double and the NCCL byte mover replaced by a gloo all-to-all.  What is
checked is everything above the C ABI on the N > 1 path: block packing by destination rank, the count
exchange, unpacking under the reference's ctx keys, reducer placement, result gathering."""
from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class GlooTransport:
    """stand-in for parallel.NcclTransport: same alltoallv contract over gloo send/recv of CPU tensors"""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def alltoallv(self, send_cols, send_off, recv_off):
        out = []
        for c in send_cols:
            r = torch.empty(recv_off[-1], dtype=c.dtype)
            reqs = []
            for p in range(self.world):
                if p == self.rank:
                    continue
                ns, nr = send_off[p + 1] - send_off[p], recv_off[p + 1] - recv_off[p]
                if ns:
                    reqs.append(dist.isend(c[send_off[p]:send_off[p + 1]].contiguous(), p))
                if nr:
                    reqs.append(dist.irecv(r[recv_off[p]:recv_off[p + 1]], p))
            for q in reqs:
                q.wait()
            own = slice(send_off[self.rank], send_off[self.rank + 1])
            r[recv_off[self.rank]:recv_off[self.rank + 1]] = c[own]
            out.append(r)
        return out


class FakePushTransport(GlooTransport):
    """stand-in for parallel.PushTransport: receive arenas are files mapped by both processes, an "address" is
    (owner rank + 1) << 48 | slot << 44 | byte offset, the count exchange is a gloo all_gather, the fence a barrier.
    Everything above the C ABI of the push shuffle -- layout of the arenas, destination tables, slot rotation,
    views handed to the reducers, fallback to the packed exchange -- runs unchanged."""

    push = True
    SLOTS = (1, 2, 3)

    def __init__(self, rank, world, tag, deny_slot=False):
        super().__init__(rank, world)
        self.tag, self.deny_slot = tag, deny_slot
        self.maps = {}                       # (owner, slot) -> np.memmap
        self.caps = {s_: [0] * world for s_ in self.SLOTS}
        self.bases = {s_: [((q + 1) << 48) | (s_ << 44) for q in range(world)] for s_ in self.SLOTS}
        self.next = 0
        self.pushes = 0
        self.bytes_sent = 0
        import fake_kernels

        fake_kernels.poke = self.poke

    def _path(self, owner, slot):
        return f"/tmp/mgb_fake_arena_{self.tag}_{owner}_{slot}"

    def poke(self, address, values):
        owner, slot, off = (address >> 48) - 1, (address >> 44) & 15, address & ((1 << 44) - 1)
        assert off % 8 == 0
        m = self.maps[owner, slot]
        m[off // 8: off // 8 + len(values)] = values

    def allgather_host(self, vec):
        out = [torch.empty_like(vec) for _ in range(self.world)]
        dist.all_gather(out, vec)
        return torch.stack(out).numpy()

    def free_slot(self):
        if self.deny_slot and self.rank == 1:
            return -1
        self.next = self.next % len(self.SLOTS) + 1
        return self.next

    def ensure(self, slots, need_bytes):
        growing = [q for q in range(self.world) if need_bytes[q] > self.caps[slots[q]][q]]
        if not growing:
            return
        for q in growing:
            self.maps.pop((q, slots[q]), None)
        dist.barrier()
        if self.rank in growing:
            n = max(need_bytes[self.rank] // 8, 1) * 2
            np.memmap(self._path(self.rank, slots[self.rank]), dtype=np.int64, mode="w+", shape=(n,)).flush()
        dist.barrier()
        for q in growing:
            m = np.memmap(self._path(q, slots[q]), dtype=np.int64, mode="r+")
            self.maps[q, slots[q]] = m
            self.caps[slots[q]][q] = m.nbytes

    def local_view(self, slot, n_elements):
        return torch.from_numpy(self.maps[self.rank, slot])[:max(n_elements, 1)]

    def upload(self, array):
        return torch.from_numpy(array)

    def fence(self):
        for m in self.maps.values():
            m.flush()
        dist.barrier()

    def timed_begin(self):
        return None

    def timed_end(self, e0, nbytes):
        self.pushes += 1
        self.bytes_sent += nbytes


def _worker(rank, world, port, case, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_kernels
        import mars_b200 as md
        from mars_b200 import executors, parallel

        executors._kernels = fake_kernels
        executors._device = lambda: torch.device("cpu")
        rng = np.random.default_rng(5)
        n = 4000
        raw = pd.DataFrame({"key": rng.integers(0, 300, n), "k2": rng.integers(0, 3, n), "v0": rng.normal(10, 3, n),
                            "v1": rng.random(n)})
        if case == "shuffle":
            ex = parallel.Exchange(GlooTransport(rank, world), rank, world)
            sess = md.new_session(exchange=ex, default=False)
            r = md.DataFrame(raw, chunk_size=500).to_gpu().groupby(["key", "k2"], sort=False) \
                .agg(["sum", "mean", "count", "min", "max", "var"], method="shuffle")
            r.execute(session=sess)
            got = sess.fetch(r)
            exp = raw.groupby(["key", "k2"]).agg(["sum", "mean", "count", "min", "max", "var"])
            q.put((rank, got.sort_index().equals(exp) or _close(got.sort_index(), exp), ex.shuffle_rows_sent,
                   sorted(set(sess.executed_ops))))
        elif case in ("push", "push_denied"):
            # fused partition -> peer transfer: mappers stop after pass 1, the exchange lays out the arenas and
            # lets pass 2 write into them; three shuffles in a row rotate the arena slots.  "push_denied": one rank
            # has no free arena -> every rank takes the packed exchange for that shuffle
            tr = FakePushTransport(rank, world, port, deny_slot=(case == "push_denied"))
            ex = parallel.Exchange(tr, rank, world)
            funcs = ["sum", "mean", "count", "min", "max", "var"]
            exp = raw.groupby(["key", "k2"]).agg(funcs)
            ok = True
            seen_ops = set()
            for rep in range(4):
                sess = md.new_session(exchange=ex, default=False)
                # rep 3: ONE chunk -- rank 1 owns neither a mapper nor a reducer and learns the schema from rank 0
                r = md.DataFrame(raw, chunk_size=(500, 500, 1300, 5000)[rep]).to_gpu().groupby(["key", "k2"], sort=False) \
                    .agg(funcs, method="shuffle")
                r.execute(session=sess)
                got = sess.fetch(r)
                ok = ok and _close(got.sort_index(), exp)
                seen_ops.update(sess.executed_ops)
            q.put((rank, ok, ex.shuffle_rows_sent, sorted(seen_ops), tr.pushes,
                   fake_kernels.calls.count("split_scatter")))
        elif case == "psrs":   # sort=True shuffle: samples -> pivots -> range shuffle across ranks
            ex = parallel.Exchange(GlooTransport(rank, world), rank, world)
            sess = md.new_session(exchange=ex, default=False)
            r = md.DataFrame(raw, chunk_size=500).to_gpu().groupby(["key", "k2"], sort=True) \
                .agg(["sum", "mean", "count", "min", "max"], method="shuffle")
            r.execute(session=sess)
            got = sess.fetch(r)
            exp = raw.groupby(["key", "k2"]).agg(["sum", "mean", "count", "min", "max"])
            q.put((rank, _close(got, exp), ex.shuffle_rows_sent, sorted(set(sess.executed_ops))))
        else:   # sharded tree: each rank owns half of the rows
            half = raw.iloc[rank * (n // 2):(rank + 1) * (n // 2)]
            sess = md.new_session(default=False)
            r = md.DataFrame(half, chunk_size=400).to_gpu().groupby("k2").agg({"v0": ["sum", "mean"], "v1": ["count"]},
                                                                             method="tree")
            got = sess.execute_sharded_tree(r)
            exp = raw.groupby("k2").agg({"v0": ["sum", "mean"], "v1": ["count"]})
            ok = True if rank != 0 else _close(got, exp)
            q.put((rank, ok, 0, sorted(set(sess.executed_ops)), got is None))
    finally:
        dist.destroy_process_group()


def _close(a, b):
    try:
        pd.testing.assert_frame_equal(a, b, rtol=1e-9, atol=0, check_exact=False)
        return True
    except AssertionError as e:  # pragma: no cover
        print(e)
        return False


def _run(case, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


@pytest.mark.timeout(180)
def test_hash_shuffle_across_two_ranks():
    out = _run("shuffle", 29641)
    for rank, ok, rows_sent, ops in out:
        assert ok, f"rank {rank}: result differs from pandas"
        assert rows_sent > 0                      # blocks really crossed ranks
        assert "DataFrameGroupByOperand:map" in ops and "DataFrameGroupByOperand:reduce" in ops


@pytest.mark.timeout(180)
def test_push_shuffle_across_two_ranks():
    """deferred split + destination tables + arena views (parallel.Exchange._shuffle_push) against pandas"""
    out = _run("push", 29644)
    for rank, ok, rows_sent, ops, pushes, scatters in out:
        assert ok, f"rank {rank}: result differs from pandas"
        assert rows_sent > 0 and pushes == 4 and scatters > 0
        assert "DataFrameGroupByOperand:map" in ops and "DataFrameGroupByOperand:reduce" in ops


@pytest.mark.timeout(180)
def test_push_shuffle_falls_back_on_every_rank_when_one_has_no_arena():
    out = _run("push_denied", 29645)
    for rank, ok, rows_sent, ops, pushes, scatters in out:
        assert ok, f"rank {rank}: result differs from pandas"
        assert rows_sent > 0 and pushes == 0


@pytest.mark.timeout(180)
def test_psrs_sort_shuffle_across_two_ranks():
    """groupby(sort=True) + shuffle over two ranks: the sample / pivot / range-split operands of
    mars/dataframe/groupby/sort.py; the fetched frame is globally sorted (no sort_index before comparing)"""
    out = _run("psrs", 29643)
    for rank, ok, rows_sent, ops in out:
        assert ok, f"rank {rank}: result differs from pandas"
        assert rows_sent > 0
    assert any("DataFrameGroupbySortShuffle:map" in ops for _, _, _, ops in out)


@pytest.mark.timeout(180)
def test_sharded_tree_across_two_ranks():
    out = _run("tree", 29642)
    assert out[0][1] and out[0][4] is False       # rank 0 owns the (correct) result
    assert out[1][4] is True                      # rank 1 returns None
    for rank, _, _, ops, _ in out:
        assert "DataFrameGroupByAgg:map" in ops
