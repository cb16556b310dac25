"""Multi-GPU plumbing for the shuffle: one process per GPU, partials stay in HBM.

``Exchange.shuffle`` replaces the storage-service transport between DataFrameGroupByOperand map and reduce
(mapper blocks `ctx[(mapper_key, reducer_index)]`, mars/core/operand/shuffle.py:111-130): every rank packs the
blocks of its mapper chunks by destination rank and one grouped ncclSend/ncclRecv all-to-all
(``mgb_comm_alltoallv``) moves all columns; the receiver re-inserts the blocks under the same ctx keys, so
the reduce executor is unchanged.  Block sizes travel through a small host-side all-gather (the 8x8 count
matrix of section 8e).  ``torch.distributed`` is used for rendezvous and host metadata only.

The byte mover is a *transport* object: ``NcclTransport`` (product) drives libmarsgb's NCCL exchange; tests
substitute a gloo transport to exercise the host logic with world_size 2 on CPU.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import MarsGBError, check
from .device import DeviceFrame


def _find_libnccl() -> Optional[str]:
    """The NCCL torch was built against (nvidia/nccl wheel), else the system one."""
    try:
        import nvidia.nccl  # type: ignore

        for root in list(nvidia.nccl.__path__):
            hits = sorted(glob.glob(os.path.join(root, "lib", "libnccl.so*")))
            if hits:
                return hits[0]
    except Exception:  # noqa: BLE001
        pass
    return None


class NcclTransport:
    """grouped ncclSend/ncclRecv all-to-all through libmarsgb (mgb_comm_alltoallv)."""

    def __init__(self, rank: int, world: int, meta_group=None):
        self.rank, self.world = rank, world
        self._lib = _lib.load()
        path = _find_libnccl()
        check(self._lib.mgb_nccl_load(path.encode() if path else None))
        uid = (C.c_uint8 * 128)()
        if rank == 0:
            check(self._lib.mgb_comm_unique_id(uid))
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=0, group=meta_group)
        uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
        self._comm = C.c_void_p()
        check(self._lib.mgb_comm_init(C.byref(self._comm), uid, rank, world))
        self.bytes_sent = 0
        self.timed = False          # set True to bracket every exchange with CUDA events (bandwidth reports)
        self._events = []           # (start, end, bytes sent to peers)

    def exchange_stats(self):
        """(total device milliseconds, bytes sent to peers) over the timed exchanges so far"""
        ms = 0.0
        nbytes = 0
        for e0, e1, b in self._events:
            e1.synchronize()
            ms += e0.elapsed_time(e1)
            nbytes += b
        return ms, nbytes

    def alltoallv(self, send_cols: Sequence[torch.Tensor], send_off: Sequence[int], recv_off: Sequence[int]):
        dev = torch.device("cuda", torch.cuda.current_device())
        recv_cols = []
        for c in send_cols:
            if not c.is_cuda:
                raise MarsGBError(2, "NCCL exchange needs CUDA tensors")
            recv_cols.append(torch.empty(recv_off[-1], dtype=c.dtype, device=dev))
        so = (C.c_int64 * (self.world + 1))(*send_off)
        ro = (C.c_int64 * (self.world + 1))(*recv_off)
        sp = _lib.ptr_array([c.data_ptr() for c in send_cols])
        rp = _lib.ptr_array([c.data_ptr() for c in recv_cols])
        own = send_off[self.rank + 1] - send_off[self.rank]
        nbytes = 8 * len(send_cols) * (send_off[-1] - own)
        if self.timed:
            # bandwidth measurement: line the ranks up first, otherwise the bracket includes waiting for a
            # peer that is still packing its blocks
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(self._lib.mgb_comm_alltoallv(self._comm, sp, so, rp, ro, len(send_cols),
                                           torch.cuda.current_stream().cuda_stream))
        if self.timed:
            e1.record()
            self._events.append((e0, e1, nbytes))
        self.bytes_sent += nbytes
        return recv_cols

    def allgather_i64(self, send: torch.Tensor, recv: torch.Tensor):
        """recv[r * n : (r + 1) * n] = rank r's ``send`` (8-byte elements; mgb_comm_allgather_i64, ncclAllGather)"""
        if not (send.is_cuda and recv.is_cuda and send.is_contiguous() and recv.is_contiguous()):
            raise MarsGBError(2, "NCCL all-gather needs contiguous CUDA tensors")
        if send.element_size() != 8 or recv.numel() != send.numel() * self.world:
            raise MarsGBError(1, "all-gather buffers do not match")
        check(self._lib.mgb_comm_allgather_i64(self._comm, send.data_ptr(), recv.data_ptr(), send.numel(),
                                               torch.cuda.current_stream().cuda_stream))

    def close(self):
        if self._comm:
            self._lib.mgb_comm_destroy(self._comm)
            self._comm = C.c_void_p()


class IpcPullTransport:
    """All-to-all on one NVSwitch box without NCCL: the send buffer is packed into a library-owned arena whose
    CUDA IPC handle travels with the block counts (same metadata all-gather); every rank then PULLS its slices
    out of the peers' arenas with cudaMemcpyAsync -- copy engines over NVLink, no SM involved (mgb_ipc_*).
    Needs all GPUs of the box visible to every process (torchrun); the NCCL transport has no such need."""

    pull = True

    def __init__(self, rank: int, world: int, meta_group=None):
        self.rank, self.world, self.meta_group = rank, world, meta_group
        self._lib = _lib.load()
        self.bytes_sent = 0
        self.timed = False
        self._events = []

    exchange_stats = NcclTransport.exchange_stats

    def reserve(self, n_elements: int):
        """-> (device address of the send arena, IPC handle bytes).  Not collective: only valid while the arena does
        not have to grow (first call / capacity probe); shuffles use ``reserve_collective``."""
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        check(self._lib.mgb_ipc_arena(int(n_elements) * 8, C.byref(ptr), handle))
        return int(ptr.value or 0), bytes(handle)

    def reserve_collective(self, n_elements: int):
        """Collective: every rank's send arena (slot 0) holds at least its ``n_elements``.  An exported allocation is
        never freed while a peer maps it: ranks whose arena must grow are announced first, every peer closes its
        mapping, and only after a barrier the arena is reallocated and its new handle published.  Returns this rank's
        arena address; ``self.peer_base[q]`` is rank q's arena as mapped here."""
        if not hasattr(self, "peer_base"):
            self.peer_base = [0] * self.world
            self._cap0 = [0] * self.world
        need = int(n_elements) * 8
        needs = [None] * self.world
        dist.all_gather_object(needs, need, group=self.meta_group)
        growing = [q for q in range(self.world) if needs[q] > self._cap0[q] or not self.peer_base[q]]
        if growing:
            for q in growing:
                if q != self.rank:
                    check(self._lib.mgb_ipc_close_slot(q, 0))
            torch.cuda.synchronize()
            dist.barrier(group=self.meta_group)
            info = None
            if self.rank in growing:
                ptr, cap = C.c_void_p(), C.c_int64()
                handle = (C.c_uint8 * 64)()
                check(self._lib.mgb_ipc_arena_slot(0, need, C.byref(ptr), handle, C.byref(cap), None))
                info = (bytes(handle), int(cap.value), int(ptr.value or 0))
            infos = [None] * self.world
            dist.all_gather_object(infos, info, group=self.meta_group)
            for q in growing:
                handle, cap, ptr = infos[q]
                self._cap0[q] = cap
                if q == self.rank:
                    self.peer_base[q] = ptr
                else:
                    mapped = C.c_void_p()
                    check(self._lib.mgb_ipc_open_slot(q, 0, (C.c_uint8 * 64).from_buffer_copy(handle), C.byref(mapped)))
                    self.peer_base[q] = int(mapped.value or 0)
        return self.peer_base[self.rank]

    def pull(self, recvbuf: torch.Tensor, recv_off, peers):
        """peers[q] = element offset of my slice in q's arena, or None"""
        stream = torch.cuda.current_stream().cuda_stream
        nbytes = 0
        if self.timed:
            torch.cuda.synchronize()
            dist.barrier(group=self.meta_group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for step in range(1, self.world):
            q = (self.rank + step) % self.world     # staggered: at every step each arena serves one puller
            off = peers[q]
            n = recv_off[q + 1] - recv_off[q]
            if off is None or n == 0:
                continue
            check(self._lib.mgb_ipc_pull(recvbuf.data_ptr() + recv_off[q] * 8, self.peer_base[q] + off * 8, n * 8, stream))
            nbytes += n * 8
        if self.timed:
            e1.record()
            self._events.append((e0, e1, nbytes))
        self.bytes_sent += nbytes        # symmetric exchange: bytes pulled == bytes served on average
        # nobody may repack its arena before every peer has finished pulling from it
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.meta_group)

    def close(self):
        pass


class _ArenaOwner:
    """exposes a library-owned receive arena to torch (``torch.as_tensor`` keeps this object alive for as long as
    any view of the arena lives: its death is the signal that the slot may be written again)"""

    def __init__(self, ptr: int, n_elements: int):
        self.__cuda_array_interface__ = {"shape": (int(n_elements),), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 2}


class PushTransport(IpcPullTransport):
    """Fused partition -> peer transfer on one NVSwitch box.  Every rank owns receive arenas that all peers map
    through CUDA IPC; after ONE fixed-size device-side exchange of the block sizes (ncclAllGather through
    libmarsgb, mgb_comm_allgather_i64) every mapper's partition kernel (mgb_split_scatter) stores its rows
    straight into the arenas of the ranks that reduce them -- SM stores over NVLink, no pack pass, no copy
    engine, no pickled metadata; a stream-ordered all-gather of one word is the fence after which every rank's
    blocks are complete.  Shuffles that cannot be deferred (the PSRS range shuffle) use the inherited pull path."""

    push = True
    SLOTS = (1, 2, 3)

    def __init__(self, rank: int, world: int, meta_group=None):
        super().__init__(rank, world, meta_group)
        self.nccl = NcclTransport(rank, world, meta_group)
        self.caps = {s_: [0] * world for s_ in self.SLOTS}     # every rank's capacity (bytes), known to all ranks
        self.bases = {s_: [0] * world for s_ in self.SLOTS}    # arenas as mapped into this process
        self._lease = {s_: None for s_ in self.SLOTS}
        self._fence = None
        self.push_events = []                                  # (e0, e1, bytes stored into peer arenas)

    # -- collective helpers -----------------------------------------------------------------------------
    def device(self):
        return torch.device("cuda", torch.cuda.current_device())

    def allgather_host(self, vec: torch.Tensor):
        """vec: device int64 [n] -> numpy int64 [world, n] on the host (the one host wait of a push shuffle)"""
        from . import kernels

        out = torch.empty(self.world * vec.numel(), dtype=torch.int64, device=vec.device)
        self.nccl.allgather_i64(vec, out)
        return kernels.to_host([out])[0].reshape(self.world, vec.numel())

    def free_slot(self) -> int:
        for s_ in self.SLOTS:
            ref = self._lease[s_]
            if ref is None or ref() is None:
                return s_
        return -1

    def ensure(self, slots: Sequence[int], need_bytes: Sequence[int]):
        """collective and deterministic: rank q's arena ``slots[q]`` holds at least need_bytes[q].  A growing arena
        is first unmapped by every peer (the exported allocation is never freed while mapped), then reallocated and
        its handle published."""
        growing = [q for q in range(self.world) if need_bytes[q] > self.caps[slots[q]][q] or not self.bases[slots[q]][q]]
        if not growing:
            return
        for q in growing:
            if q != self.rank:
                check(self._lib.mgb_ipc_close_slot(q, slots[q]))
        torch.cuda.synchronize()
        dist.barrier(group=self.meta_group)
        info = None
        if self.rank in growing:
            ptr, cap = C.c_void_p(), C.c_int64()
            handle = (C.c_uint8 * 64)()
            check(self._lib.mgb_ipc_arena_slot(slots[self.rank], int(need_bytes[self.rank]), C.byref(ptr), handle,
                                               C.byref(cap), None))
            info = (bytes(handle), int(cap.value), int(ptr.value or 0))
        infos = [None] * self.world
        dist.all_gather_object(infos, info, group=self.meta_group)
        for q in growing:
            handle, cap, ptr = infos[q]
            self.caps[slots[q]][q] = cap
            if q == self.rank:
                self.bases[slots[q]][q] = ptr
            else:
                mapped = C.c_void_p()
                check(self._lib.mgb_ipc_open_slot(q, slots[q], (C.c_uint8 * 64).from_buffer_copy(handle), C.byref(mapped)))
                self.bases[slots[q]][q] = int(mapped.value or 0)

    def local_view(self, slot: int, n_elements: int) -> torch.Tensor:
        import weakref

        owner = _ArenaOwner(self.bases[slot][self.rank], max(int(n_elements), 1))
        self._lease[slot] = weakref.ref(owner)
        return torch.as_tensor(owner, device=self.device())

    def upload(self, array) -> torch.Tensor:
        return torch.from_numpy(array).to(self.device(), non_blocking=True)

    def fence(self):
        """stream-ordered: when it completes on a rank, the partition kernels of ALL ranks have finished"""
        if self._fence is None:
            self._fence = (torch.zeros(1, dtype=torch.int64, device=self.device()),
                           torch.zeros(self.world, dtype=torch.int64, device=self.device()))
        self.nccl.allgather_i64(*self._fence)

    def timed_begin(self):
        if not self.timed:
            return None
        torch.cuda.synchronize()
        dist.barrier(group=self.meta_group)
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return e0

    def timed_end(self, e0, nbytes):
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.push_events.append((e0, e1, nbytes))
        self.bytes_sent += nbytes

    def push_stats(self):
        ms, nbytes = 0.0, 0
        for e0, e1, b in self.push_events:
            e1.synchronize()
            ms += e0.elapsed_time(e1)
            nbytes += b
        return ms, nbytes

    def close(self):
        self.nccl.close()


def _concat(pieces: List[torch.Tensor], like: torch.Tensor) -> torch.Tensor:
    pieces = [p for p in pieces if p.numel()]
    if not pieces:
        return like[:0]
    if like.is_cuda:
        from . import kernels

        return kernels.concat_columns(pieces)
    return torch.cat(pieces)  # CPU tensors only occur in host-logic tests (gloo transport)


def reducer_rank(r: int, n_reducers: int, world: int) -> int:
    """Rank that runs reducer r: contiguous ranges of reducers per rank (the reference lets its scheduler place
    reducers, mars/services/task/analyzer/analyzer.py:382-413 -- any fixed map is valid).  Ranges make the
    blocks a mapper sends to one peer a single row range of its partitioned columns."""
    per = -(-n_reducers // world)
    return min(r // per, world - 1)


class Exchange:
    def __init__(self, transport, rank: int, world: int, meta_group=None):
        self.transport = transport
        self.rank, self.world = rank, world
        self.meta_group = meta_group
        self.shuffle_rows_sent = 0

    # ------------------------------------------------------------------------------------------ helpers
    def _gather_meta(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.meta_group)
        return out

    @staticmethod
    def _unwrap(entry):
        """shuffle block in ctx -> (payload, wrapped): the hash shuffle stores ``(mapper_index, payload)``
        (groupby/core.py:389-435), the range shuffle of sort=True the bare tuple of frames (sort.py:113-135)"""
        if (isinstance(entry, tuple) and len(entry) == 2 and not isinstance(entry[0], DeviceFrame)
                and isinstance(entry[1], (tuple, DeviceFrame))):
            return entry[1], True
        return entry, False

    @staticmethod
    def _payload_frames(payload) -> Tuple[List[DeviceFrame], bool]:
        if isinstance(payload, tuple):
            return [f for f in payload if isinstance(f, DeviceFrame)], True
        return [payload], False

    @staticmethod
    def _schema(frames: List[DeviceFrame]):
        return dict(index_names=frames[0].index_names, nk=len(frames[0].index),
                    slots=[(f.columns, [str(t.dtype) for t in f.data]) for f in frames])

    @staticmethod
    def _flatten(frames: List[DeviceFrame]) -> List[torch.Tensor]:
        for f in frames[1:]:
            if not frames[0].same_index(f):
                raise NotImplementedError("exchange expects the partial frames of one block to share their index")
        cols = list(frames[0].index)
        for f in frames:
            cols.extend(f.data)
        return cols

    @staticmethod
    def _rebuild(schema, cols: List[torch.Tensor], a: int, b: int, as_tuple: bool):
        nk = schema["nk"]
        keys = [t[a:b] for t in cols[:nk]]
        frames, pos = [], nk
        for columns, _ in schema["slots"]:
            frames.append(DeviceFrame(schema["index_names"], keys, columns, [t[a:b] for t in cols[pos:pos + len(columns)]],
                                      b - a))
            pos += len(columns)
        return tuple(frames) + (False,) if as_tuple else frames[0]

    # ------------------------------------------------------------------------------------------ shuffle
    def shuffle(self, ctx, proxy_chunk, owners: Dict[str, int]):
        """Moves every mapper block to the rank of its reducer.  Reducers are placed in contiguous ranges
        (``reducer_rank``), so the blocks one mapper sends to one peer are ONE row range of its scattered
        columns; the send buffer holds, per peer, all columns of all local mappers ([column][mapper][rows]) and
        travels as a single message per peer.  The receiver's blocks are views into the receive buffer."""
        mappers = list(proxy_chunk.op.inputs)
        R = int(mappers[0].op.shuffle_size)
        P = self.world
        local = [i for i, m in enumerate(mappers) if owners[m.key] == self.rank]
        if getattr(ctx, "defer_split", False) and getattr(self.transport, "push", False):
            if self._shuffle_push(ctx, mappers, R, local, owners):
                return
        for i in local:      # mappers that stopped after pass 1: blocks of local columns for the packed exchange
            d = ctx.pop((mappers[i].key, "deferred"), None)
            if d is not None:
                d.emit(ctx)

        def key_of(m, r):
            return (m.key, (r, m.index[1] if len(m.index) > 1 else 0))

        counts, schema, as_tuple, wrapped = {}, None, True, True
        for i in local:
            m = mappers[i]
            for r in range(R):
                payload, wrapped = self._unwrap(ctx[key_of(m, r)])
                frames, as_tuple = self._payload_frames(payload)
                counts[(i, r)] = len(frames[0])
                if schema is None:
                    schema = self._schema(frames)
        reducers_of = [[r for r in range(R) if reducer_rank(r, R, P) == p] for p in range(P)]
        on_gpu = torch.cuda.is_available()
        dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
        pull = bool(getattr(self.transport, "pull", False))
        # ---- pack (needs local information only): peer p, column c, local mappers ascending, reducers of p
        send_rows = [0 if p == self.rank else sum(counts[(i, r)] for i in local for r in reducers_of[p]) for p in range(P)]
        ncols_local = 0 if schema is None else schema["nk"] + sum(len(c) for c, _ in schema["slots"])
        send_off = [0]
        for p in range(P):
            send_off.append(send_off[-1] + ncols_local * send_rows[p])
        if pull:
            send_ptr = self.transport.reserve_collective(max(send_off[-1], 1))
            sendbuf = None
        else:
            sendbuf = torch.empty(max(send_off[-1], 1), dtype=torch.int64, device=dev)
        segments = []
        for p in range(P):
            if p == self.rank:
                continue
            row0 = 0
            for i in local:
                m = mappers[i]
                for r in reducers_of[p]:
                    key = key_of(m, r)
                    frames, _ = self._payload_frames(self._unwrap(ctx[key])[0])
                    n = len(frames[0])
                    if n:
                        for c, t in enumerate(self._flatten(frames)):
                            start = send_off[p] + c * send_rows[p] + row0
                            if pull:
                                segments.append((t, send_ptr + start * 8))
                            else:
                                segments.append((t, sendbuf[start:start + n].view(t.dtype)))
                    row0 += n
                    del ctx[key]
        if segments:
            if on_gpu:
                from . import kernels

                kernels.copy_segments(segments)
            else:   # CPU tensors only occur in host-logic tests (gloo transport)
                for src, dst in segments:
                    dst.copy_(src)
        if pull:
            torch.cuda.current_stream().synchronize()   # the arena is complete before its handle is published
        # ---- one metadata all-gather: block counts (+ the arena handle and slice offsets for pull transports)
        metas = self._gather_meta(dict(counts=counts, schema=schema, as_tuple=as_tuple, wrapped=wrapped,
                                       send_off=send_off if pull else None))
        all_counts = {}
        for m_ in metas:
            all_counts.update(m_["counts"])
            if schema is None and m_["schema"] is not None:
                schema, as_tuple, wrapped = m_["schema"], m_["as_tuple"], m_["wrapped"]

        def entry(m, block):   # hash shuffle blocks carry the mapper index, range shuffle blocks do not
            return (m.index, block) if wrapped else block

        if schema is None:
            if pull:
                dist.barrier(group=self.meta_group)
            return
        owner_of_mapper = [owners[m.key] for m in mappers]
        ncols = schema["nk"] + sum(len(c) for c, _ in schema["slots"])
        recv_rows = [0 if q == self.rank else
                     sum(all_counts[(i, r)] for i, o in enumerate(owner_of_mapper) if o == q for r in reducers_of[self.rank])
                     for q in range(P)]
        recv_off = [0]
        for q in range(P):
            recv_off.append(recv_off[-1] + ncols * recv_rows[q])
        if pull:
            recvbuf = torch.empty(max(recv_off[-1], 1), dtype=torch.int64, device=dev)
            peers = [None if q == self.rank or metas[q]["send_off"] is None else metas[q]["send_off"][self.rank]
                     for q in range(P)]
            self.transport.pull(recvbuf, recv_off, peers)
        else:
            recvbuf = self.transport.alltoallv([sendbuf], send_off, recv_off)[0]
        self.shuffle_rows_sent += sum(send_rows)
        # unpack: views into the receive buffer
        dtypes = [torch.int64] * schema["nk"]
        for _, dts in schema["slots"]:
            dtypes += [getattr(torch, d.replace("torch.", "")) for d in dts]
        for q in range(P):
            if q == self.rank or recv_rows[q] == 0:
                if q != self.rank:
                    for i, m in enumerate(mappers):
                        if owner_of_mapper[i] == q:
                            for r in reducers_of[self.rank]:
                                ctx[key_of(m, r)] = entry(m, self._rebuild(
                                    schema, [recvbuf[:0].view(dt) for dt in dtypes], 0, 0, as_tuple))
                continue
            cols = [recvbuf[recv_off[q] + c * recv_rows[q]: recv_off[q] + (c + 1) * recv_rows[q]].view(dtypes[c])
                    for c in range(ncols)]
            pos = 0
            for i, m in enumerate(mappers):
                if owner_of_mapper[i] != q:
                    continue
                for r in reducers_of[self.rank]:
                    n = all_counts[(i, r)]
                    ctx[key_of(m, r)] = entry(m, self._rebuild(schema, cols, pos, pos + n, as_tuple))
                    pos += n

    # ------------------------------------------------------------------------------------ push shuffle
    def _shuffle_push(self, ctx, mappers, R: int, local: List[int], owners: Dict[str, int]) -> bool:
        """Hash shuffle with the fused partition -> peer transfer (PushTransport).  Collective; returns False --
        on every rank alike -- when some rank could not defer a mapper or has no free receive arena; the caller
        then runs the packed exchange.

        Layout of rank p's arena: for every source rank q, for every column c, the blocks (mapper of q ascending,
        reducer of p ascending) back to back, each padded to an even number of rows so that every block starts on a
        16-byte boundary (16-byte stores in the partition kernel); the reducers' blocks are views into it."""
        import numpy as np

        t = self.transport
        P = self.world
        per_rank = [[i for i, m in enumerate(mappers) if owners[m.key] == q] for q in range(P)]
        M = max(1, max(len(x) for x in per_rank))
        splits = [ctx.get((mappers[i].key, "deferred")) for i in local]
        ok = all(d is not None and d.R == R for d in splits)
        slot = t.free_slot() if ok else -1
        ncols = len(splits[0].cols) if (ok and splits) else 0
        head = np.zeros(4 + M * (R + 1), dtype=np.int64)
        head[:3] = (1 if (ok and slot > 0) else 0, slot, ncols)
        vec = t.upload(head)
        if ok:
            for j, d in enumerate(splits):
                vec[4 + j * (R + 1): 4 + (j + 1) * (R + 1)].copy_(d.offsets)
        allv = t.allgather_host(vec)                                   # [P, 4 + M (R + 1)]: the one host wait
        if not all(int(allv[q, 0]) == 1 for q in range(P)):
            return False
        slots = [int(allv[q, 1]) for q in range(P)]
        ncols = max(int(allv[q, 2]) for q in range(P))
        if ncols == 0:
            for i in local:
                ctx.pop((mappers[i].key, "deferred"), None)
            return True
        offs = allv[:, 4:].reshape(P, M, R + 1)
        cnt = offs[:, :, 1:] - offs[:, :, :-1]                          # [source rank, its j-th mapper, reducer]
        for q in range(P):
            cnt[q, len(per_rank[q]):, :] = 0
        pad = (cnt + 1) & ~1
        red_lo = [min((r for r in range(R) if reducer_rank(r, R, P) == p), default=0) for p in range(P)]
        red_hi = [max((r + 1 for r in range(R) if reducer_rank(r, R, P) == p), default=0) for p in range(P)]
        seg_rows = np.zeros((P, P), dtype=np.int64)                     # [receiver p, source q]: padded rows per column
        within = [None] * P                                             # [p][q, j, r - lo]: row of the block in its segment
        for p in range(P):
            sub = pad[:, :, red_lo[p]:red_hi[p]]
            flat = sub.reshape(P, -1)
            csum = np.cumsum(flat, axis=1)
            within[p] = (csum - flat).reshape(sub.shape)
            seg_rows[p] = flat.sum(axis=1)
        seg_base = np.zeros((P, P), dtype=np.int64)                     # first element of segment (p, q) in p's arena
        seg_base[:, 1:] = np.cumsum(ncols * seg_rows, axis=1)[:, :-1]
        need = [int(8 * ncols * seg_rows[p].sum()) for p in range(P)]
        t.ensure(slots, need)
        # ---- destination tables of the local mappers: [mapper, column, reducer] -> address
        q = self.rank
        sent = 0
        if local:
            tab = np.zeros((len(local), ncols, R), dtype=np.int64)
            cix = np.arange(ncols, dtype=np.int64)[:, None]
            for p in range(P):
                lo, hi = red_lo[p], red_hi[p]
                if hi <= lo:
                    continue
                base = t.bases[slots[p]][p]
                for j in range(len(local)):
                    tab[j, :, lo:hi] = base + 8 * (seg_base[p, q] + cix * seg_rows[p, q] + within[p][q, j][None, :])
                if p != q:
                    sent += int(cnt[q, :len(local), lo:hi].sum())
            dtab = t.upload(tab.reshape(len(local), -1))
            e0 = t.timed_begin()
            from . import executors      # the kernels module the executors are bound to

            for j, d in enumerate(splits):
                executors._kernels.split_scatter(d.keys, R, d.hist, d.cols, table=dtab[j], mode=d.MODE)
            t.timed_end(e0, sent * ncols * 8)
        else:
            t.timed_end(t.timed_begin(), 0)
        t.fence()
        self.shuffle_rows_sent += sent
        # ---- the blocks of this rank's reducers: views into its arena
        schema = splits[0].schema() if splits else None
        if any(not x for x in per_rank):
            metas = self._gather_meta(schema)
            schema = next(m_ for m_ in metas if m_ is not None)
        dtypes = [torch.int64] * schema["nk"]
        for _, dts in schema["slots"]:
            dtypes += [getattr(torch, d_.replace("torch.", "")) for d_ in dts]
        arena = t.local_view(slots[q], need[q] // 8)
        p = self.rank
        lo, hi = red_lo[p], red_hi[p]
        template = splits[0] if splits else None
        is_tuple = template.is_tuple if template is not None else True
        from .device import ColumnSet

        nk = schema["nk"]
        spans, pos = [], nk          # (columns, first data column, one past the last) per partial frame of a block
        for columns, _ in schema["slots"]:
            spans.append((columns, pos, pos + len(columns)))
            pos += len(columns)
        for src in range(P):
            rows = int(seg_rows[p, src])
            base = int(seg_base[p, src])
            # one view per column of the segment; the blocks are ROW RANGES of these (nothing is sliced per block:
            # with 64 mappers x 8 reducers x 42 columns the slicing alone was most of a step's host time)
            cset = ColumnSet([arena[base + c * rows: base + (c + 1) * rows].view(dtypes[c]) for c in range(ncols)])
            for j, i in enumerate(per_rank[src]):
                m = mappers[i]
                for r in range(lo, hi):
                    a = int(within[p][src, j, r - lo])
                    b = a + int(cnt[src, j, r])
                    frames = [DeviceFrame.from_range(schema["index_names"], columns, cset, 0, nk, d0, d1, a, b)
                              for columns, d0, d1 in spans]
                    block = tuple(frames) + (False,) if is_tuple else frames[0]
                    ctx[(m.key, (r, m.index[1] if len(m.index) > 1 else 0))] = (m.index, block)
        for i in local:
            ctx.pop((mappers[i].key, "deferred"), None)
        return True

    # ------------------------------------------------------------------------------- point-to-point fetch
    def transfer(self, ctx, key, src: int, dst: int):
        """Move ctx[key] (a tuple of partial frames) from rank src to rank dst (tree branch gather)."""
        box = [None]
        if self.rank == src:
            value = ctx[key]
            frames, as_tuple = self._payload_frames(value)
            if frames and all(isinstance(f, DeviceFrame) for f in frames):
                box = [dict(schema=self._schema(frames), n=len(frames[0]), as_tuple=as_tuple)]
            else:   # small host objects (PSRS samples / pivots: pandas index frames, numpy arrays)
                box = [dict(obj=value)]
        dist.broadcast_object_list(box, src=src, group=self.meta_group)
        meta = box[0]
        if "obj" in meta:
            if self.rank == dst:
                ctx[key] = meta["obj"]
            return
        if self.rank not in (src, dst):
            return
        n = meta["n"]
        send_off = [0] * (self.world + 1)
        recv_off = [0] * (self.world + 1)
        if self.rank == src:
            cols = self._flatten(self._payload_frames(ctx[key])[0])
            for p in range(dst + 1, self.world + 1):
                send_off[p] = n
        else:
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
            cols = [torch.empty(0, dtype=torch.int64, device=dev) for _ in range(meta["schema"]["nk"])]
            for columns, dts in meta["schema"]["slots"]:
                cols += [torch.empty(0, dtype=getattr(torch, d.replace("torch.", "")), device=dev) for d in dts]
            for p in range(src + 1, self.world + 1):
                recv_off[p] = n
        recv = self.transport.alltoallv(cols, send_off, recv_off)
        if self.rank == dst:
            v = self._rebuild(meta["schema"], recv, 0, n, meta["as_tuple"])
            if meta["as_tuple"]:
                v = v[:-1]  # partial tuples outside the shuffle carry no deliver_by flag
            ctx[key] = v

    def fetch_result(self, obj, owner: int):
        box = [obj if self.rank == owner else None]
        dist.broadcast_object_list(box, src=owner, group=self.meta_group)
        return box[0]

    def sync_size_recorders(self, chunks):
        from .dataframe import SizeRecorder

        names = sorted({c.op.size_recorder_name for c in chunks if getattr(c.op, "size_recorder_name", None)})
        for name in names:
            rec = SizeRecorder.lookup(name)
            rec.get(exact=True)   # evaluates deferred sizes (every rank needs real numbers to agree on the plan)
            rec._bound = [None] * len(rec._agg)
            rec.bounded = False
            metas = self._gather_meta((list(rec._raw), list(rec._agg)))
            rec._raw = [x for m in metas for x in m[0]]
            rec._agg = [x for m in metas for x in m[1]]


def init_exchange(backend: Optional[str] = None) -> Exchange:
    """Create the exchange for this process from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK,
    MASTER_ADDR / MASTER_PORT).  One process per GPU."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if torch.cuda.is_available():
        torch.cuda.set_device(local % torch.cuda.device_count())
    if not dist.is_initialized():
        dist.init_process_group(backend=backend or "gloo", rank=rank, world_size=world)
    # host metadata travels over gloo (the default group, or a gloo side group when the caller initialised NCCL);
    # bytes move through libmarsgb
    meta_group = None
    if world > 1 and dist.get_backend() != "gloo":
        meta_group = dist.new_group(backend="gloo")
    # transport: MGB_EXCHANGE = nccl | ipc | auto (default).  auto = the peer-to-peer pull transport when every
    # rank can map every peer's arena (one box, all GPUs visible, P2P capable), else grouped ncclSend/ncclRecv
    kind = os.environ.get("MGB_EXCHANGE", "auto")     # push (default where possible) | ipc (pull only) | nccl
    transport = None
    if kind in ("push", "ipc", "auto") and world > 1 and torch.cuda.is_available():
        ok = False
        try:
            t = (IpcPullTransport if kind == "ipc" else PushTransport)(rank, world, meta_group)
            _, handle = t.reserve(1)
            handles = [None] * world
            dist.all_gather_object(handles, handle, group=meta_group)
            lib = _lib.load()
            for q, h in enumerate(handles):
                if q != rank:
                    src = C.c_void_p()
                    check(lib.mgb_ipc_open(q, (C.c_uint8 * 64).from_buffer_copy(h), C.byref(src)))
            ok = True
        except Exception:  # noqa: BLE001
            if kind in ("ipc", "push"):
                raise
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=meta_group)
        if all(flags):
            transport = t
    if transport is None:
        transport = NcclTransport(rank, world, meta_group)
    return Exchange(transport, rank, world, meta_group)
