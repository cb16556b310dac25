"""GPU-resident chunk store: intermediate partials of the groupby path stay in HBM across subtask boundaries.

The reference keeps chunk data in a ``StorageBackend`` per storage level (mars/storage/base.py:94-293); its GPU
level is ``CudaStorage`` (mars/storage/cuda.py:212-311), which serialises cudf / cupy objects into device buffers.
With cudf absent, the values the GPU executors exchange -- ``DeviceFrame`` / ``PiecewiseFrame`` partial tuples,
``DeviceResult`` result chunks -- need a home of their own:

  ``DeviceChunkStore``      the backend: same methods, argument meaning and results as ``StorageBackend`` (``put`` /
                            ``get`` / ``delete`` / ``object_info`` / ``list`` / ``open_writer`` / ``open_reader``,
                            ``level`` GPU, ``ObjectInfo`` sizes for the quota).  ``put`` keeps the object BY HANDLE --
                            no copy, no serialisation; its size is the bytes it occupies in HBM.  The file-object
                            pair moves bytes between workers: ``open_reader`` serialises a held object once (one D2H
                            per column), ``open_writer`` collects bytes and ``get`` materialises them (H2D) on first use.
  ``dumps`` / ``loads``     the wire format of the device types: a pickled header + the raw 8-byte columns.
  ``register_on_mars()``    (needs the real ``mars``) registers ``MgbGpuStorage`` -- the same store as a subclass of the
                            reference's ``StorageBackend`` -- with ``register_storage_backend``, and a ``Serializer``
                            for the device types (mars/serialization/core.pyx:165-181) so they can be fetched across
                            processes by the reference's own transfer code (mars/services/storage/transfer.py:63-206).
"""
from __future__ import annotations

import io
import os
import pickle
import struct
import sys
import uuid
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .device import DeviceFrame, PiecewiseFrame

_MAGIC = b"MGBDEV1\0"


@dataclass
class ObjectInfo:
    """mirror of mars.storage.base.ObjectInfo (base.py:87-91)"""
    size: int = None
    device: int = None
    object_id: Any = None


# ----------------------------------------------------------------------------------------------------
# wire format of the device types
# ----------------------------------------------------------------------------------------------------
def _is_device_value(obj) -> bool:
    from .executors import DeviceResult

    if isinstance(obj, (DeviceFrame, PiecewiseFrame, DeviceResult)):
        return True
    return isinstance(obj, tuple) and len(obj) > 0 and any(
        isinstance(x, (DeviceFrame, PiecewiseFrame, DeviceResult)) for x in obj)


def device_nbytes(obj) -> int:
    """bytes the value occupies in HBM (what the quota of the GPU level accounts)"""
    if isinstance(obj, tuple):
        return sum(device_nbytes(x) for x in obj)
    n = getattr(obj, "nbytes", None)
    if n is not None and not isinstance(obj, (np.ndarray,)):
        return int(n)
    if isinstance(obj, np.ndarray):
        return int(obj.nbytes)
    return int(sys.getsizeof(obj))


def _frame_parts(f: DeviceFrame):
    keys = list(f.index) if f.index is not None else []
    return dict(kind="frame", index_names=f.index_names, has_index=f.index is not None, columns=f.columns, n=len(f),
                dtypes=[str(t.dtype) for t in keys + list(f.data)], range_start=f.range_start), keys + list(f.data)


def _encode(obj, cols: List[torch.Tensor]):
    """-> header object; appends the value's columns to ``cols``"""
    from .executors import DeviceResult

    if isinstance(obj, PiecewiseFrame):
        return dict(kind="pieces", pieces=[_encode(p, cols) for p in obj.pieces])
    if isinstance(obj, DeviceFrame):
        h, ts = _frame_parts(obj)
        cols.extend(ts)
        return h
    if isinstance(obj, DeviceResult):
        ts = list(obj.columns) + list(obj.index)
        cols.extend(ts)
        return dict(kind="result", labels=obj.labels, ncol=len(obj.columns), meta=obj.meta,
                    dtypes=[str(t.dtype) for t in ts], n=len(obj))
    if isinstance(obj, tuple):
        return dict(kind="tuple", items=[_encode(x, cols) for x in obj])
    return dict(kind="host", value=obj)


def _decode(h, cols: List[torch.Tensor]):
    from .executors import DeviceResult

    k = h["kind"]
    if k == "pieces":
        return PiecewiseFrame([_decode(p, cols) for p in h["pieces"]])
    if k == "tuple":
        return tuple(_decode(x, cols) for x in h["items"])
    if k == "host":
        return h["value"]
    nt = len(h["dtypes"])
    ts = [cols.pop(0) for _ in range(nt)]
    if k == "result":
        return DeviceResult(h["labels"], ts[:h["ncol"]], ts[h["ncol"]:], h["meta"])
    nk = len(h["index_names"]) if h["has_index"] else 0
    f = DeviceFrame(h["index_names"], ts[:nk] if h["has_index"] else None, h["columns"], ts[nk:], h["n"])
    f.range_start = h.get("range_start")
    return f


def dumps(obj) -> bytes:
    """header (pickle) + raw columns, each column read from the device with one copy"""
    cols: List[torch.Tensor] = []
    header = _encode(obj, cols)
    hb = pickle.dumps(dict(header=header, lengths=[int(t.numel()) for t in cols]), protocol=pickle.HIGHEST_PROTOCOL)
    out = io.BytesIO()
    out.write(_MAGIC)
    out.write(struct.pack("<Q", len(hb)))
    out.write(hb)
    for t in cols:
        out.write(t.detach().contiguous().cpu().numpy().tobytes())
    return out.getvalue()


def loads(data, device=None):
    """inverse of ``dumps``; the columns are uploaded to ``device`` (default: the current CUDA device)"""
    mv = memoryview(data)
    if bytes(mv[:8]) != _MAGIC:
        raise ValueError("not a device-chunk byte stream")
    (hl,) = struct.unpack("<Q", mv[8:16])
    meta = pickle.loads(mv[16:16 + hl])
    pos = 16 + hl
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    dtypes = []

    def walk(h):
        if h["kind"] in ("pieces",):
            for p in h["pieces"]:
                walk(p)
        elif h["kind"] == "tuple":
            for x in h["items"]:
                walk(x)
        elif h["kind"] != "host":
            dtypes.extend(h["dtypes"])

    walk(meta["header"])
    cols = []
    for n, dt in zip(meta["lengths"], dtypes):
        npdt = np.float64 if "float64" in dt else np.int64
        arr = np.frombuffer(mv, dtype=npdt, count=n, offset=pos).copy()
        pos += n * 8
        cols.append(torch.from_numpy(arr).to(device))
    return _decode(meta["header"], cols)


# ----------------------------------------------------------------------------------------------------
# file objects (mars/storage/core.py StorageFileObject protocol: read / write / seek / tell / close)
# ----------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self, store: "DeviceChunkStore", object_id: str, size: Optional[int]):
        self._store, self.object_id, self._buf, self._closed = store, object_id, io.BytesIO(), False

    def write(self, data):
        return self._buf.write(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data)

    def tell(self):
        return self._buf.tell()

    def close(self):
        if not self._closed:
            self._closed = True
            self._store._bytes[self.object_id] = self._buf.getvalue()

    async def __aenter__(self):
        return self

    async def __aexit__(self, *exc):
        self.close()


class _Reader:
    def __init__(self, data: bytes, object_id: str):
        self._buf, self.object_id = io.BytesIO(data), object_id

    def read(self, size=-1):
        return self._buf.read(size)

    def seek(self, offset, whence=os.SEEK_SET):
        if whence == os.SEEK_SET and offset < 0:
            raise ValueError("negative seek position")
        return self._buf.seek(offset, whence)

    def tell(self):
        return self._buf.tell()

    def close(self):
        pass

    async def __aenter__(self):
        return self

    async def __aexit__(self, *exc):
        self.close()


class DeviceChunkStore:
    """GPU-level chunk store (mirror of mars.storage.base.StorageBackend; template mars/storage/cuda.py:212-311)"""

    name = "mgb_gpu"
    is_seekable = True

    def __init__(self, size: Optional[int] = None, device: Optional[int] = None):
        self._size = size
        self._device = device
        self._objects: Dict[str, Any] = {}
        self._bytes: Dict[str, bytes] = {}        # written through open_writer, not materialised yet
        self._sizes: Dict[str, int] = {}

    # -- lifecycle ------------------------------------------------------------------------------------
    @classmethod
    async def setup(cls, **kwargs) -> Tuple[Dict, Dict]:
        size = kwargs.pop("size", None)
        device = kwargs.pop("device", None)
        if kwargs:
            raise TypeError(f'DeviceChunkStore got unexpected config: {",".join(kwargs)}')
        return dict(size=size, device=device), dict()

    @staticmethod
    async def teardown(**kwargs):
        pass

    @property
    def level(self):
        return "GPU"

    @property
    def size(self):
        return self._size

    @property
    def used(self) -> int:
        return sum(self._sizes.values())

    @property
    def backend_info(self) -> dict:
        return {"name": self.name}

    # -- objects --------------------------------------------------------------------------------------
    async def put(self, obj, importance: int = 0) -> ObjectInfo:
        object_id = str(uuid.uuid4())
        size = device_nbytes(obj)
        if self._size is not None and _is_device_value(obj) and self.used + size > self._size:
            raise MemoryError(f"GPU chunk store quota exceeded: {self.used} + {size} > {self._size}")
        self._objects[object_id] = obj            # by handle: the partials never leave HBM
        self._sizes[object_id] = size
        return ObjectInfo(size=size, device=self._device, object_id=object_id)

    async def get(self, object_id, **kwargs):
        if object_id in self._objects:
            return self._objects[object_id]
        data = self._bytes[object_id]             # arrived through open_writer: materialise on first use
        obj = self._materialise(data)
        self._objects[object_id] = obj
        self._sizes[object_id] = device_nbytes(obj)
        del self._bytes[object_id]
        return obj

    def _materialise(self, data: bytes):
        if data[:8] == _MAGIC:
            dev = torch.device("cuda", self._device) if (self._device is not None and torch.cuda.is_available()) else None
            return loads(data, dev)
        return data

    async def delete(self, object_id):
        self._objects.pop(object_id, None)
        self._bytes.pop(object_id, None)
        self._sizes.pop(object_id, None)

    async def object_info(self, object_id) -> ObjectInfo:
        if object_id in self._sizes:
            return ObjectInfo(size=self._sizes[object_id], device=self._device, object_id=object_id)
        return ObjectInfo(size=len(self._bytes[object_id]), device=self._device, object_id=object_id)

    async def list(self) -> List:
        return list(self._objects) + list(self._bytes)

    async def open_writer(self, size=None):
        return _Writer(self, str(uuid.uuid4()), size)

    async def open_reader(self, object_id):
        if object_id in self._bytes:
            return _Reader(self._bytes[object_id], object_id)
        return _Reader(self._serialise(self._objects[object_id]), object_id)

    def _serialise(self, obj) -> bytes:
        if isinstance(obj, (bytes, bytearray)):
            return bytes(obj)
        return dumps(obj)

    async def fetch(self, object_id):
        return None

    async def pin(self, object_id):
        return None

    async def unpin(self, object_id):
        return None


# ----------------------------------------------------------------------------------------------------
# the same store behind the reference's own interfaces
# ----------------------------------------------------------------------------------------------------
_mars_backend = None


def register_on_mars():
    """Registers (i) ``MgbGpuStorage``, a ``mars.storage.base.StorageBackend`` at ``StorageLevel.GPU`` that holds
    the GPU executors' values by handle, under the backend name "mgb_gpu"; (ii) a ``Serializer`` for the device
    types, so that ``mars.serialization.serialize`` / ``AioSerializer`` -- what the reference's storage handler and
    transfer code call -- can move them between processes.  Returns the backend class."""
    global _mars_backend
    if _mars_backend is not None:       # idempotent: the serializer registry asserts one class per serializer id
        return _mars_backend
    from mars.serialization.core import Serializer, buffered
    from mars.serialization import serialize as mars_serialize, deserialize as mars_deserialize
    from mars.storage.base import ObjectInfo as MarsObjectInfo, StorageBackend, StorageLevel, register_storage_backend
    from mars.storage.core import StorageFileObject
    from mars.utils import implements

    from .executors import DeviceResult

    class DeviceValueSerializer(Serializer):
        """header = our pickled header; buffers = the raw columns as host byte buffers"""

        @buffered
        def serial(self, obj, context):
            cols: List[torch.Tensor] = []
            header = _encode(obj, cols)
            bufs = [memoryview(t.detach().contiguous().cpu().numpy()).cast("B") for t in cols]
            return (pickle.dumps(header), [int(t.numel()) for t in cols]), bufs, True

        def deserial(self, serialized, context, subs):
            hb, lengths = serialized[:2]
            header = pickle.loads(hb)
            dtypes = []

            def walk(h):
                if h["kind"] == "pieces":
                    for p in h["pieces"]:
                        walk(p)
                elif h["kind"] == "tuple":
                    for x in h["items"]:
                        walk(x)
                elif h["kind"] != "host":
                    dtypes.extend(h["dtypes"])

            walk(header)
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
            cols = []
            for n, dt, buf in zip(lengths, dtypes, subs):
                arr = np.frombuffer(buf, dtype=np.float64 if "float64" in dt else np.int64, count=n).copy()
                cols.append(torch.from_numpy(arr).to(dev))
            return _decode(header, cols)

    for t in (DeviceFrame, PiecewiseFrame, DeviceResult):
        DeviceValueSerializer.register(t)

    @register_storage_backend
    class MgbGpuStorage(StorageBackend):
        name = "mgb_gpu"
        is_seekable = True

        def __init__(self, size=None, device=None):
            self._store = DeviceChunkStore(size=size, device=device)

        @classmethod
        @implements(StorageBackend.setup)
        async def setup(cls, **kwargs):
            return await DeviceChunkStore.setup(**kwargs)

        @staticmethod
        @implements(StorageBackend.teardown)
        async def teardown(**kwargs):
            pass

        @property
        @implements(StorageBackend.level)
        def level(self):
            return StorageLevel.GPU

        @property
        @implements(StorageBackend.size)
        def size(self):
            return self._store.size

        @implements(StorageBackend.get)
        async def get(self, object_id, **kwargs):
            if object_id in self._store._bytes:     # came through open_writer: the reference's wire format
                from mars.serialization import AioDeserializer

                blob = self._store._bytes[object_id]
                obj = await AioDeserializer(StorageFileObject(_Reader(blob, object_id), object_id=object_id)).run()
                self._store._objects[object_id] = obj
                self._store._sizes[object_id] = device_nbytes(obj)
                del self._store._bytes[object_id]
                return obj
            return await self._store.get(object_id)

        @implements(StorageBackend.put)
        async def put(self, obj, importance=0):
            info = await self._store.put(obj, importance)
            return MarsObjectInfo(size=info.size, device=info.device, object_id=info.object_id)

        @implements(StorageBackend.delete)
        async def delete(self, object_id):
            await self._store.delete(object_id)

        @implements(StorageBackend.object_info)
        async def object_info(self, object_id):
            info = await self._store.object_info(object_id)
            return MarsObjectInfo(size=info.size, device=info.device, object_id=info.object_id)

        @implements(StorageBackend.open_writer)
        async def open_writer(self, size=None):
            w = await self._store.open_writer(size)
            return StorageFileObject(w, object_id=w.object_id)

        @implements(StorageBackend.open_reader)
        async def open_reader(self, object_id):
            if object_id in self._store._objects:
                from mars.serialization import AioSerializer

                bufs = await AioSerializer(self._store._objects[object_id]).run()
                data = b"".join(bytes(b) for b in bufs)
                return StorageFileObject(_Reader(data, object_id), object_id=object_id)
            r = await self._store.open_reader(object_id)
            return StorageFileObject(r, object_id=object_id)

        @implements(StorageBackend.list)
        async def list(self):
            return await self._store.list()

    del mars_serialize, mars_deserialize
    _mars_backend = MgbGpuStorage
    return MgbGpuStorage
