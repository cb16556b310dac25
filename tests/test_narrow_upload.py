"""int64 columns cross PCIe narrow (DataFrameToGPU, mars/dataframe/base/to_gpu.py:28-35): the host pass is
range-checked (lossless or refused), the device pass restores the 8-byte column every kernel of the path reads."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pandas as pd
import pytest

import mars_groupby_oracle as orc
from conftest import assert_frames_match

WIDTHS = {1: np.int8, 2: np.int16, 4: np.int32}


@pytest.fixture(scope="module")
def lib():
    from mars_b200 import _lib

    return _lib.load()


def _narrow(lib, src, width):
    dst = np.full(len(src), -1, dtype=WIDTHS[width])
    fits = C.c_int32(-1)
    assert lib.mgb_host_narrow_i64(src.ctypes.data, len(src), width, dst.ctypes.data, C.byref(fits)) == 0
    return fits.value, dst


@pytest.fixture(params=["avx2-if-present", "baseline"])
def isa(request, lib):
    """both bodies of the host pass: the ones picked at run time, and the SSE2 / scalar fallback"""
    got = lib.mgb_host_narrow_isa(1 if request.param == "baseline" else 0)
    assert got == 1 if request.param == "baseline" else got in (1, 2)
    yield got
    lib.mgb_host_narrow_isa(0)


@pytest.mark.parametrize("width", [1, 2, 4])
def test_host_narrowing_is_lossless_or_refused(lib, isa, width):
    # host code of the library: no device is touched
    info = np.iinfo(WIDTHS[width])
    rng = np.random.default_rng(width)
    for n in (0, 1, 7, 15, 16, 17, 33, 4095, 4096, 4097, 100_003):
        src = rng.integers(info.min, info.max + 1, n, dtype=np.int64)
        if n > 2:
            src[0], src[-1] = info.min, info.max
        fits, dst = _narrow(lib, src, width)
        assert fits == 1
        np.testing.assert_array_equal(dst.astype(np.int64), src)
    for bad in (info.max + 1, info.min - 1, np.iinfo(np.int64).max, np.iinfo(np.int64).min, 1 << 32, -(1 << 32) + 5,
                (1 << 32) + 3, (1 << 40) | 7, -(1 << 63) + 1):
        for pos in (0, 3, 4095, 4096, 4100, 99_990, 99_999):
            src = rng.integers(0, 3, 100_000, dtype=np.int64)
            src[pos] = bad
            assert _narrow(lib, src, width)[0] == 0, (bad, pos)


@pytest.mark.parametrize("width", [1, 2, 4])
def test_host_narrowing_of_integer_valued_floats(lib, isa, width):
    info = np.iinfo(WIDTHS[width])
    rng = np.random.default_rng(10 + width)

    def narrow(src):
        dst = np.full(len(src), -1, dtype=WIDTHS[width])
        fits = C.c_int32(-1)
        assert lib.mgb_host_narrow_f64(src.ctypes.data, len(src), width, dst.ctypes.data, C.byref(fits)) == 0
        return fits.value, dst

    for n in (0, 1, 7, 8, 15, 16, 17, 33, 4097, 100_003):
        src = rng.integers(info.min, info.max + 1, n).astype(np.float64)
        if n > 2:
            src[0], src[-1] = info.min, info.max
        fits, dst = narrow(src)
        assert fits == 1
        back = dst.astype(np.float64)
        np.testing.assert_array_equal(back.view(np.int64), src.view(np.int64))      # the same bits
    for bad in (0.5, -0.0, np.nan, np.inf, -np.inf, float(info.max) + 1, float(info.min) - 1, 1e300, 5e-324, 3.0000000000000004):
        for pos in (0, 5, 4096, 4111, 9_990, 9_999):
            src = rng.integers(0, 50, 10_000).astype(np.float64)
            src[pos] = bad
            assert narrow(src)[0] == 0, (bad, pos)


def test_host_narrowing_rejects_bad_arguments(lib):
    src = np.zeros(4, dtype=np.int64)
    fits = C.c_int32(0)
    assert lib.mgb_host_narrow_i64(src.ctypes.data, 4, 3, src.ctypes.data, C.byref(fits)) == 1
    assert b"width" in lib.mgb_last_error()
    assert lib.mgb_host_narrow_i64(src.ctypes.data, 4, 1, src.ctypes.data, None) == 1
    assert lib.mgb_host_narrow_i64(None, 4, 1, src.ctypes.data, C.byref(fits)) == 1


def test_width_is_chosen_from_a_sample_and_checked_on_the_column():
    from mars_b200 import device

    assert device._narrow_width(np.arange(-128, 128, dtype=np.int64)) == 1
    assert device._narrow_width(np.arange(0, 300_000, 7, dtype=np.int64)) == 4
    assert device._narrow_width(np.array([0, 1 << 40], dtype=np.int64)) == 8
    assert device._narrow_width(np.full(5000, -32768, dtype=np.int64)) == 2
    assert device._narrow_width(np.arange(1, 51, dtype=np.float64)) == 1
    assert device._narrow_width(np.array([1.0, 2.5])) == 8
    assert device._narrow_width(np.array([1.0, np.nan])) == 8
    assert device._narrow_width(np.array([1.0, np.inf])) == 8
    assert device._narrow_width(np.array([1.0, 3e9])) == 8
    assert device._narrow_width(np.array([-40000.0, 7.0])) == 4


class _HostStandIn:
    """the library with the two device entries replaced by numpy on host memory: lets the host logic of the narrow
    upload (slab ring, deferred block, width hints) run without a GPU"""

    def __init__(self, lib):
        self._lib = lib
        self.widened = 0
        self.stream_idle = False

    def __getattr__(self, name):
        return getattr(self._lib, name)

    def _widen(self, src, width, n, dst, out_dtype):
        a = np.ctypeslib.as_array(C.cast(src, C.POINTER(C.c_int8)), shape=(n * width,)).view(WIDTHS[width])
        o = np.ctypeslib.as_array(C.cast(dst, C.POINTER(C.c_int64)), shape=(n,)).view(out_dtype)
        o[:] = a
        self.widened += 1
        return 0

    def mgb_widen_i64(self, src, width, n, dst, stream):
        return self._widen(src, width, n, dst, np.int64)

    def mgb_widen_f64(self, src, width, n, dst, stream):
        return self._widen(src, width, n, dst, np.float64)


@pytest.fixture
def host_device(monkeypatch, lib):
    import torch
    from mars_b200 import _lib, device

    class Event:
        def record(self):
            pass

        def synchronize(self):
            pass

    class Stream:
        cuda_stream = 0

        def query(self):
            return stand_in.stream_idle

    stand_in = _HostStandIn(lib)
    monkeypatch.setattr(_lib, "_lib", stand_in)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "Event", Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: Stream())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch.Tensor, "is_pinned", lambda self: True)
    monkeypatch.setattr(device, "NARROW_MIN_BYTES", 1 << 20)
    monkeypatch.setattr(device, "NARROW", True)
    monkeypatch.setattr(device, "NARROW_FLOATS", True)
    monkeypatch.setattr(device, "_staging", device._StagingRing())
    monkeypatch.setattr(device, "_pinned", device._PinnedPool())
    for learned in (device._no_narrow, device._narrow_hint):
        learned.clear()
    yield stand_in
    for learned in (device._no_narrow, device._narrow_hint):
        learned.clear()


@pytest.mark.parametrize("dtype", [np.int64, np.float64])
@pytest.mark.parametrize("width,n", [(1, 140_000), (2, 9_000_001), (4, 5_000_000)])
def test_slab_ring_pieces_and_threads(host_device, width, n, dtype):
    from mars_b200 import device

    info = np.iinfo(WIDTHS[width])
    src = np.random.default_rng(n).integers(info.min, info.max + 1, n, dtype=np.int64)
    src[0], src[-1] = info.min, info.max
    src = src.astype(dtype)
    before, h2d = device.narrowed_columns, device.h2d_bytes
    out = device._to_device(src, "cpu", "col")
    assert device.narrowed_columns == before + 1 and device.h2d_bytes - h2d == n * width and host_device.widened == 1
    np.testing.assert_array_equal(out.numpy().view(np.int64), src.view(np.int64))
    assert device._narrow_hint["col"] == width


def test_deferred_block_hints_and_refusals(host_device):
    from mars_b200 import device

    n = 200_000
    rng = np.random.default_rng(3)
    small = [rng.integers(0, 100, n) for _ in range(3)]
    grow = [rng.integers(0, 100, n), rng.integers(1000, 1100, n)]
    holed = rng.integers(0, 100, n).astype(np.float64)
    holed[1000] = np.nan                              # not in the sample: refused by the pass over the column
    price = rng.uniform(1, 2, n)
    before, h2d = device.narrowed_columns, device.h2d_bytes
    with device.deferred_uploads():
        outs = [device._to_device(a, "cpu", "small") for a in small]
        outs += [device._to_device(a, "cpu", "grow") for a in grow]
        o_holed = device._to_device(holed, "cpu", "holed")
        o_price = device._to_device(price, "cpu", "price")
        assert device.narrowed_columns == before                        # nothing is accounted before the block ends
        with device.deferred_uploads():                                 # nested: the outer block flushes
            outs.append(device._to_device(small[0], "cpu", "small"))
        assert device.narrowed_columns == before
    for o, a in zip(outs, small + grow + [small[0]]):
        np.testing.assert_array_equal(o.numpy(), a)
    np.testing.assert_array_equal(o_holed.numpy().view(np.int64), holed.view(np.int64))
    np.testing.assert_array_equal(o_price.numpy(), price)
    assert device.narrowed_columns - before == 6 and device.deferred_columns >= 6
    assert device.h2d_bytes - h2d == n * (4 * 1 + 1 + 2 + 8 + 8)
    assert device._narrow_hint == {"small": 1, "grow": 2} and device._no_narrow == {"price", "holed"}
    # outside a block: synchronous, the remembered width is used and dropped when it stops fitting
    wide = rng.integers(0, 1 << 40, n)
    np.testing.assert_array_equal(device._to_device(wide, "cpu", "small").numpy(), wide)
    assert "small" not in device._narrow_hint and "small" not in device._no_narrow
    np.testing.assert_array_equal(device._to_device(wide, "cpu", "small").numpy(), wide)
    assert "small" in device._no_narrow                                  # a freshly sampled width that fails: not again


def test_deferred_block_belongs_to_its_thread(host_device):
    import threading
    from mars_b200 import device

    n = 200_000
    a = np.random.default_rng(8).integers(0, 100, n)
    got = {}
    before = device.narrowed_columns
    with device.deferred_uploads():
        mine = device._to_device(a, "cpu", "mine")
        t = threading.Thread(target=lambda: got.setdefault("out", device._to_device(a, "cpu", "theirs")))
        t.start()
        t.join()
        # the other thread's column is complete when its call returns (its consumer does not know about this block)
        np.testing.assert_array_equal(got["out"].numpy(), a)
        assert device.narrowed_columns == before + 1
    np.testing.assert_array_equal(mine.numpy(), a)
    assert device.narrowed_columns == before + 2


def test_copy_engine_never_waits_for_a_host_pass_that_has_not_started(host_device, monkeypatch):
    import threading
    from concurrent.futures import ThreadPoolExecutor
    from mars_b200 import device

    n = 200_000
    rng = np.random.default_rng(4)
    cols = [rng.integers(0, 100, n) for _ in range(3)]
    device._staging._init()
    pool = ThreadPoolExecutor(1)
    monkeypatch.setattr(device._staging, "npool", pool)
    gate = threading.Event()
    pool.submit(gate.wait)                                 # the only worker is busy: the passes below cannot start
    before, skipped = device.narrowed_columns, device.skipped_columns
    host_device.stream_idle = True                         # ... and the stream has nothing left to do
    try:
        with device.deferred_uploads():
            outs = [device._to_device(a, "cpu", "busy") for a in cols]
    finally:
        gate.set()
    for o, a in zip(outs, cols):
        np.testing.assert_array_equal(o.numpy(), a)
    assert device.skipped_columns - skipped == 3 and device.narrowed_columns == before
    assert "busy" not in device._no_narrow and "busy" not in device._narrow_hint      # not a refusal
    host_device.stream_idle = False                        # copies in flight: the passes are waited for
    with device.deferred_uploads():
        outs = [device._to_device(a, "cpu", "busy") for a in cols]
    for o, a in zip(outs, cols):
        np.testing.assert_array_equal(o.numpy(), a)
    assert device.narrowed_columns - before == 3 and device.skipped_columns - skipped == 3
    pool.shutdown()


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("dtype", [np.int64, np.float64])
@pytest.mark.parametrize("width,n", [(1, 131_072), (1, 4_194_304 + 5), (2, 9_000_001), (4, 300_007), (4, 5_000_000)])
def test_narrow_upload_restores_the_column(cuda_device, monkeypatch, width, n, dtype, pinned):
    import torch
    from mars_b200 import device

    monkeypatch.setattr(device, "NARROW_MIN_BYTES", 1 << 20)

    info = np.iinfo(WIDTHS[width])
    rng = np.random.default_rng(n)
    src = rng.integers(info.min, info.max + 1, n, dtype=np.int64)
    src[::997] = info.min
    src[5::997] = info.max
    src = src.astype(dtype)
    if pinned:
        host = torch.empty(n, dtype=torch.int64).pin_memory()
        arr = host.numpy().view(dtype)
        arr[:] = src
    else:
        arr = src
    before, h2d = device.narrowed_columns, device.h2d_bytes
    out = device._to_device(arr, cuda_device, f"c{width}{np.dtype(dtype).name}")
    assert device.narrowed_columns == before + 1 and device.h2d_bytes - h2d == n * width
    assert out.dtype == (torch.int64 if dtype == np.int64 else torch.float64) and out.shape == (n,)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.int64), src.view(np.int64))


@pytest.mark.gpu
def test_column_that_does_not_fit_is_uploaded_as_it_is(cuda_device, monkeypatch):
    from mars_b200 import device

    monkeypatch.setattr(device, "NARROW_MIN_BYTES", 1 << 20)
    n = 400_000
    src = np.arange(n, dtype=np.int64) % 100
    src[n - 3] = 1 << 45                      # the sample says int8; the range check of the full column refuses
    device._no_narrow.discard("wide")
    before, h2d = device.narrowed_columns, device.h2d_bytes
    out = device._to_device(src, cuda_device, "wide")
    np.testing.assert_array_equal(out.cpu().numpy(), src)
    assert device.narrowed_columns == before and device.h2d_bytes - h2d == n * 8
    assert "wide" in device._no_narrow        # later chunks of that column skip the attempt
    device._no_narrow.discard("wide")


@pytest.mark.gpu
@pytest.mark.parametrize("method,sort", [("tree", True), ("shuffle", False)])
def test_host_chunks_with_narrow_keys_match_the_oracle(cuda_device, monkeypatch, method, sort):
    import mars_b200 as md
    from mars_b200 import device

    monkeypatch.setattr(device, "NARROW_MIN_BYTES", 1 << 20)
    rng = np.random.default_rng(5)
    n, chunk = 900_000, 300_000
    raw = pd.DataFrame({"nflag": rng.integers(-2, 3, n), "ncode": rng.integers(0, 40_000, n),
                        "nqty": rng.integers(1, 51, n), "nprice": rng.normal(100, 20, n),
                        "nunits": rng.integers(0, 300, n).astype(np.float64)})
    raw.loc[rng.integers(0, chunk, 50), "nunits"] = np.nan          # chunk 0 of that column is refused (NaN), like nprice
    funcs = {"nqty": ["sum", "min", "max", "count"], "nprice": ["sum", "mean", "var"], "nunits": ["sum", "mean", "count"]}
    before = device.narrowed_columns
    sess = md.new_session(default=False)
    got = md.DataFrame(raw, chunk_size=chunk).to_gpu().groupby(["nflag", "ncode"], sort=sort).agg(funcs, method=method) \
        .execute(session=sess).fetch(session=sess)
    exp = orc.run_groupby_agg(raw, ["nflag", "ncode"], funcs, chunk, method=method, sort=sort)
    assert_frames_match(got, exp, sort_index=not sort)
    # nflag (1 byte), ncode (4 bytes), nqty (1 byte) of 3 chunks; nprice never; nunits: NaN in its first chunk -> the label
    # is not tried again
    assert device.narrowed_columns - before == 9
    assert {"nprice", "nunits"} <= device._no_narrow


@pytest.mark.gpu
@pytest.mark.parametrize("defer", [True, False, "opportunistic"])
def test_band_launch_narrows_in_the_background(cuda_device, monkeypatch, defer):
    """low-cardinality band: the batch launch collects the columns of all host chunks first, so the narrowing of the
    key columns runs in worker threads while the other columns are enqueued (device.deferred_uploads); a column with a
    value that does not fit is uploaded at 8 bytes when the block ends, and a width remembered from an earlier chunk
    that stops fitting is sampled again"""
    import mars_b200 as md
    from mars_b200 import device

    monkeypatch.setattr(device, "NARROW_MIN_BYTES", 1 << 20)
    monkeypatch.setattr(device, "NARROW_DEFER", bool(defer))
    monkeypatch.setattr(device, "NARROW_OPPORTUNISTIC", defer == "opportunistic")
    for learned in (device._no_narrow, device._narrow_hint):
        learned.clear()
    rng = np.random.default_rng(11)
    n, chunk = 1_200_000, 300_000
    raw = pd.DataFrame({"bflag": rng.integers(0, 3, n), "bstat": rng.integers(-1, 1, n),
                        "bqty": rng.integers(1, 51, n).astype(np.float64), "bprice": rng.uniform(900, 105000, n),
                        "bunits": rng.integers(0, 100, n).astype(np.float64), "bgrow": rng.integers(0, 100, n)})
    raw.loc[2 * chunk + 1000, "bunits"] = np.nan           # chunk 2: not sampled, found by the pass over the column
    raw.loc[chunk:, "bgrow"] += 1000                       # chunk 0 fits one byte, the later ones need two
    funcs = {"bqty": ["sum", "mean"], "bprice": ["sum", "mean"], "bunits": ["sum", "count"], "bgrow": ["sum", "count"]}
    before, dbefore, sbefore = device.narrowed_columns, device.deferred_columns, device.skipped_columns
    sess = md.new_session(default=False)
    got = md.DataFrame(raw, chunk_size=chunk).to_gpu().groupby(["bflag", "bstat"]).agg(funcs, method="tree") \
        .execute(session=sess).fetch(session=sess)
    exp = orc.run_groupby_agg(raw, ["bflag", "bstat"], funcs, chunk, method="tree", sort=True)
    assert_frames_match(got, exp, sort_index=False)
    # bflag, bstat, bqty: 4 chunks each; bunits: chunks 0, 1, 3 (chunk 2 is refused by the pass, the label stays);
    # bgrow: every chunk is sampled when the block collects them all, else chunk 1 is refused with chunk 0's width
    if defer == "opportunistic":      # small chunks: the copies finish early, passes that have not started are skipped
        # 20 columns are candidates; the one with the NaN is refused if its pass ran, skipped if it did not
        assert device.narrowed_columns - before + device.skipped_columns - sbefore in (19, 20)
        return
    assert device.narrowed_columns - before == 12 + 3 + (4 if defer else 3) and device.skipped_columns == sbefore
    assert (device.deferred_columns - dbefore > 0) == defer
    assert "bprice" in device._no_narrow and device._narrow_hint.get("bgrow") == 2
