"""GPU-resident chunk store: the backend honours the reference's storage contract
(mars/storage/base.py:94-293; test contract mars/storage/tests/test_libs.py:142-233) and keeps the GPU executors'
values by handle.  CPU tests use frames over CPU tensors (host logic); the real-Mars backend is exercised when the
reference build (baseline/_ref) is present; -m gpu tests hold real HBM columns."""
from __future__ import annotations

import asyncio
import os

import numpy as np
import pandas as pd
import pytest
import torch

import ref_runner


def _frame(device, n=1000, seed=0, with_index=True):
    from mars_b200.device import DeviceFrame

    rng = np.random.default_rng(seed)
    idx = [torch.from_numpy(rng.integers(0, 50, n)).to(device)] if with_index else None
    data = [torch.from_numpy(rng.random(n)).to(device), torch.from_numpy(rng.integers(0, 9, n)).to(device)]
    return DeviceFrame(["key"] if with_index else [], idx, ["s", "c"], data, n)


def _same(a, b):
    pd.testing.assert_frame_equal(a.to_pandas(), b.to_pandas())


def _contract(store_cls, device, by_handle_check=True):
    from mars_b200.device import PiecewiseFrame
    from mars_b200.executors import DeviceResult
    from mars_b200 import storage

    async def run():
        params, teardown = await store_cls.setup(size=1 << 30)
        st = store_cls(**params)
        # base operations of the reference's contract: numpy + pandas (incl. a string column) round trip, sizes agree
        a = np.random.rand(10, 10)
        ia = await st.put(a)
        np.testing.assert_array_equal(a, await st.get(ia.object_id))
        assert (await st.object_info(ia.object_id)).size == ia.size
        df = pd.DataFrame({"col1": np.arange(10), "col2": [f"str{i}" for i in range(10)], "col3": np.random.rand(10)})
        idf = await st.put(df)
        pd.testing.assert_frame_equal(df, await st.get(idf.object_id))
        assert len(await st.list()) == 2
        await st.delete(idf.object_id)
        assert len(await st.list()) == 1
        # the GPU executors' values: held by handle, size == bytes in HBM
        f1, f2 = _frame(device, 1000, 1), _frame(device, 10, 2)
        tup = (f1, PiecewiseFrame([f1, f2]), False)
        it = await st.put(tup)
        got = await st.get(it.object_id)
        if by_handle_check:
            assert got is tup and got[0].data[0].data_ptr() == f1.data[0].data_ptr()      # no copy
        assert it.size == f1.nbytes + (f1.nbytes + f2.nbytes) + storage.device_nbytes(False)
        # cross-worker path: reader on the sender's store -> writer on the receiver's store -> get
        st2 = store_cls(**params)
        r = await st.open_reader(it.object_id)
        async with r as reader:
            blob = await _maybe(reader.read())
        async with await st2.open_writer(size=len(blob)) as w:
            await _maybe(w.write(blob[:100]))
            await _maybe(w.write(blob[100:]))
        back = await st2.get(w.object_id)
        assert isinstance(back, tuple) and back[2] is False
        _same(back[0], f1)
        pd.testing.assert_frame_equal(back[1].to_pandas(), tup[1].to_pandas())
        assert (await st2.object_info(w.object_id)).size == it.size
        # result chunks
        res = DeviceResult([("v", "sum")], [f1.data[0]], [f1.index[0]],
                           dict(names=["key"], col_value=pd.MultiIndex.from_tuples([("v", "sum")]), two_level=True,
                                as_index=True, dtypes=None))
        ir = await st.put(res)
        assert ir.size == res.nbytes
        async with await st.open_reader(ir.object_id) as reader:
            blob = await _maybe(reader.read())
        async with await st2.open_writer() as w:
            await _maybe(w.write(blob))
        back_res = await st2.get(w.object_id)
        assert back_res.labels == res.labels and len(back_res) == len(res)
        assert all(torch.equal(x.cpu(), y.cpu()) for x, y in zip(back_res.columns + back_res.index, res.columns + res.index))
        if device.type == "cuda":
            pd.testing.assert_frame_equal(back_res.to_pandas(), res.to_pandas())
        # seek / tell of the reader
        async with await st.open_reader(ir.object_id) as reader:
            with pytest.raises((OSError, ValueError)):
                await _maybe(reader.seek(-1))
            assert 5 == await _maybe(reader.seek(5))
            assert 10 == await _maybe(reader.seek(5, os.SEEK_CUR))
            assert 10 == await _maybe(reader.tell())
        await store_cls.teardown(**teardown)

    asyncio.run(run())


async def _maybe(x):
    return await x if asyncio.iscoroutine(x) or isinstance(x, asyncio.Future) else x


def test_device_chunk_store_contract_on_cpu_tensors():
    from mars_b200.storage import DeviceChunkStore

    _contract(DeviceChunkStore, torch.device("cpu"))


def test_quota_counts_device_bytes():
    from mars_b200.storage import DeviceChunkStore

    async def run():
        st = DeviceChunkStore(size=10_000)
        f = _frame(torch.device("cpu"), 300)          # 300 rows x 3 columns x 8 B = 7200 B
        await st.put(f)
        assert st.used == 7200
        with pytest.raises(MemoryError):
            await st.put(_frame(torch.device("cpu"), 300, seed=3))

    asyncio.run(run())


@pytest.mark.skipif(not ref_runner.reference_available(), reason="reference build (baseline/_ref) not present")
def test_backend_registered_on_the_real_mars_passes_the_reference_contract():
    """same sequence through `MgbGpuStorage(StorageBackend)`: registered under its name, level GPU, and the device
    types travel through the reference's own AioSerializer / AioDeserializer (our Serializer is registered)"""
    ref_runner._import_reference()
    from mars.serialization import AioDeserializer, AioSerializer
    from mars.storage.base import StorageLevel, get_storage_backend
    from mars_b200 import storage

    cls = storage.register_on_mars()
    assert get_storage_backend("mgb_gpu") is cls
    st = cls(size=None)
    assert st.level == StorageLevel.GPU
    _contract(cls, torch.device("cpu"))

    async def run():
        import io

        f = _frame(torch.device("cpu"), 500, 4)
        from mars.storage.core import StorageFileObject
        from mars_b200.storage import _Reader

        bufs = await AioSerializer((f, False)).run()
        back = await AioDeserializer(StorageFileObject(_Reader(b"".join(bytes(b) for b in bufs), "x"), object_id="x")).run()
        assert back[1] is False
        _same(back[0], f)
        # reference test_reader_and_writer: raw AioSerializer buffers of a numpy array through writer / reader
        t = np.random.random(10)
        buffers = await AioSerializer(t).run()
        size = sum(getattr(buf, "nbytes", len(buf)) for buf in buffers)
        async with await st.open_writer(size=size) as writer:
            for buf in buffers:
                await writer.write(buf)
        async with await st.open_reader(writer.object_id) as reader:
            r = await AioDeserializer(reader).run()
        np.testing.assert_array_equal(t, r)
        np.testing.assert_array_equal(t, await st.get(writer.object_id))

    asyncio.run(run())


@pytest.mark.gpu
def test_device_chunk_store_holds_hbm_columns_by_handle(cuda_device):
    from mars_b200.storage import DeviceChunkStore

    _contract(DeviceChunkStore, cuda_device)


@pytest.mark.gpu
def test_partials_cross_a_subtask_boundary_through_the_store(cuda_device):
    """map operands in one 'subtask', combine + agg in another: the partial tuples are parked in the chunk store
    in between (put -> ctx cleared -> get), never leaving HBM; result equals the oracle"""
    import mars_b200 as md
    import mars_groupby_oracle as orc
    from conftest import assert_frames_match
    from mars_b200.core import ProcessorContext, build_chunk_graph, execute
    from mars_b200.storage import DeviceChunkStore

    rng = np.random.default_rng(5)
    n = 200_000
    raw = pd.DataFrame({"k": rng.integers(0, 300, n), "v0": rng.normal(10, 3, n), "v1": rng.random(n)})
    func = ["sum", "mean", "count", "min", "max", "var"]
    r = md.DataFrame(raw, chunk_size=25_000).to_gpu().groupby("k").agg(func, method="tree")
    t = r.data
    gen = type(t.op).tile(t.op)
    try:
        next(gen)
    except StopIteration as stop:
        t = stop.value
    order = build_chunk_graph(t.chunks)
    store = DeviceChunkStore()
    ids = {}

    async def run():
        ctx = ProcessorContext()
        for c in order:                                   # subtask 1: data sources, H2D, map operands
            if type(c.op).__name__ == "DataFrameGroupByAgg" and c.op.stage.name != "map":
                continue
            if type(c.op).__name__ == "DataFrameConcat":
                continue
            ctx.set_current_chunk(c)
            execute(ctx, c.op)
            if type(c.op).__name__ == "DataFrameGroupByAgg":
                info = await store.put(ctx[c.key])
                assert info.size > 0
                ids[c.key] = info.object_id
        ptrs = {k: ctx[k][0].ptrs_raw()[1][0] for k in ids}
        ctx.clear()
        ctx2 = ProcessorContext()                         # subtask 2: everything downstream of the map stage
        for k, oid in ids.items():
            ctx2[k] = await store.get(oid)
            assert ctx2[k][0].ptrs_raw()[1][0] == ptrs[k]     # same HBM columns: nothing was copied
        for c in order:
            if c.key in ctx2 or type(c.op).__name__ in ("DataFrameDataSource", "DataFrameToGPU"):
                continue
            if type(c.op).__name__ == "DataFrameGroupByAgg" and c.op.stage.name == "map":
                continue
            ctx2.set_current_chunk(c)
            execute(ctx2, c.op)
        return ctx2[t.chunks[0].key]

    got = asyncio.run(run())
    exp = orc.run_groupby_agg(raw, "k", func, 25_000, method="tree", sort=True)
    assert_frames_match(got, exp, sort_index=False)
