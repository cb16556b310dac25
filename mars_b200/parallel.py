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
        """-> (device address of the send arena, IPC handle bytes)"""
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        check(self._lib.mgb_ipc_arena(int(n_elements) * 8, C.byref(ptr), handle))
        return int(ptr.value or 0), bytes(handle)

    def pull(self, recvbuf: torch.Tensor, recv_off, peers):
        """peers[q] = (handle bytes, element offset of my slice in q's arena) or None"""
        stream = torch.cuda.current_stream().cuda_stream
        nbytes = 0
        if self.timed:
            torch.cuda.synchronize()
            dist.barrier(group=self.meta_group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for step in range(1, self.world):
            q = (self.rank + step) % self.world     # staggered: at every step each arena serves one puller
            info = peers[q]
            n = recv_off[q + 1] - recv_off[q]
            if info is None or n == 0:
                continue
            handle, off = info
            src = C.c_void_p()
            hb = (C.c_uint8 * 64).from_buffer_copy(handle)
            check(self._lib.mgb_ipc_open(q, hb, C.byref(src)))
            check(self._lib.mgb_ipc_pull(recvbuf.data_ptr() + recv_off[q] * 8, (src.value or 0) + off * 8, n * 8, stream))
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
        handle = None
        if pull:
            send_ptr, handle = self.transport.reserve(max(send_off[-1], 1))
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
        metas = self._gather_meta(dict(counts=counts, schema=schema, as_tuple=as_tuple, wrapped=wrapped, handle=handle,
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
            peers = [None if q == self.rank or metas[q]["handle"] is None else
                     (metas[q]["handle"], metas[q]["send_off"][self.rank]) for q in range(P)]
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
    meta_group = None  # default group is gloo: host metadata only; bytes move through libmarsgb
    # transport: MGB_EXCHANGE = nccl | ipc | auto (default).  auto = the peer-to-peer pull transport when every
    # rank can map every peer's arena (one box, all GPUs visible, P2P capable), else grouped ncclSend/ncclRecv
    kind = os.environ.get("MGB_EXCHANGE", "auto")
    transport = None
    if kind in ("ipc", "auto") and world > 1 and torch.cuda.is_available():
        ok = False
        try:
            t = IpcPullTransport(rank, world, meta_group)
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
            if kind == "ipc":
                raise
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=meta_group)
        if all(flags):
            transport = t
    if transport is None:
        transport = NcclTransport(rank, world, meta_group)
    return Exchange(transport, rank, world, meta_group)
