"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/mgb.h declares
(no compute calls here)."""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mgb.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from mars_b200 import _lib

    return _lib.load()


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("mgb_hash_keys", "mgb_partition_ids", "mgb_partition_scatter", "mgb_agg_create", "mgb_agg_update",
                 "mgb_agg_finalize", "mgb_agg_emit", "mgb_post_mean", "mgb_post_var", "mgb_comm_alltoallv"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    from mars_b200 import _lib

    for name in declared_symbols():
        assert hasattr(lib, name), f"libmarsgb.so does not export {name}"
        assert name in _lib.SIGNATURES, f"ctypes binding lacks {name}"
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert lib.mgb_version() == 100


def test_library_is_sm100a_only():
    from mars_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_errors_are_statuses_not_aborts(lib):
    # invalid arguments come back as a status + message; nothing touches the device
    from mars_b200 import _lib

    st = lib.mgb_hash_keys(None, 1, 5, None, None)
    assert st == 1 and b"null" in lib.mgb_last_error()
    spec = _lib.make_spec(1, [_lib.F64], [(0, _lib.SUM)])
    spec.n_keys = 7
    h = ctypes.c_void_p()
    st = lib.mgb_agg_create(ctypes.byref(h), ctypes.byref(spec), None)
    assert st == 3
    with pytest.raises(_lib.MarsGBError):
        _lib.check(st)


def test_cpu_tensors_fail_loudly():
    import torch
    from mars_b200 import MarsGBError, kernels

    with pytest.raises(MarsGBError):
        kernels.partition_ids([torch.arange(3, dtype=torch.int64)], 4)
    with pytest.raises(MarsGBError):
        kernels.groupby_aggregate([torch.arange(3, dtype=torch.int64)], [torch.zeros(3, dtype=torch.float64)],
                                  [(0, kernels.SUM)])
