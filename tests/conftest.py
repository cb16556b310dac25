"""Shared fixtures.  `-m gpu` tests need a B200 and call through the C ABI of libmarsgb.so; everything else
runs on CPU.  The oracle (oracle/) is imported here as the checker only."""
from __future__ import annotations

import glob
import json
import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _frame(z, prefix) -> pd.DataFrame:
    meta = json.loads(str(z[f"{prefix}_meta"]))
    nlev = int(z[f"{prefix}_nlevels"])
    levels = [z[f"{prefix}_index{i}"] for i in range(nlev)]
    ncols = int(z[f"{prefix}_ncols"])
    data = {i: z[f"{prefix}_col{i}"] for i in range(ncols)}
    if nlev == 1:
        index = pd.Index(levels[0], name=meta["index_names"][0])
    else:
        index = pd.MultiIndex.from_arrays(levels, names=meta["index_names"])
    df = pd.DataFrame(data, index=index)
    cols = meta["columns"]
    if cols and isinstance(cols[0], list):
        df.columns = pd.MultiIndex.from_tuples([tuple(c) for c in cols])
    else:
        df.columns = pd.Index(cols)
    return df


class Golden:
    """One tests/golden/*.npz fixture: inputs + outputs of the unmodified reference (oracle/make_golden.py)."""

    def __init__(self, path):
        self.name = os.path.splitext(os.path.basename(path))[0]
        z = np.load(path, allow_pickle=False)
        self.spec = json.loads(str(z["spec"]))
        self.raw = pd.DataFrame({c: z[f"in_{c}"] for c in self.spec["in_columns"]})
        self.result = _frame(z, "result")
        self.chunk_results = [_frame(z, f"chunk{i}") for i in range(int(z["n_chunk_results"]))]
        self.maps = []
        for i in range(int(z["n_maps"])):
            steps = json.loads(str(z[f"map{i}_steps"]))
            self.maps.append({k: _frame(z, f"map{i}_{k}") for k in steps})

    @property
    def by(self):
        return self.spec["by"]

    @property
    def func(self):
        return self.spec["func"]


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name) -> Golden:
    return Golden(os.path.join(GOLDEN_DIR, name + ".npz"))


def assert_frames_match(got: pd.DataFrame, exp: pd.DataFrame, sort_index: bool, rtol=1e-9):
    """Parity bar of BASELINE.json: group keys / counts / min / max bit-exact, float64 sum / mean / var within
    1e-9 relative; same column labels, order and dtypes."""
    if not isinstance(got, pd.DataFrame):
        got = got.to_pandas()                 # device-resident result chunk (executors.DeviceResult)
    if sort_index:
        got, exp = got.sort_index(), exp.sort_index()
    assert list(got.columns) == list(exp.columns), (list(got.columns), list(exp.columns))
    assert got.index.nlevels == exp.index.nlevels
    for i in range(exp.index.nlevels):
        np.testing.assert_array_equal(got.index.get_level_values(i).to_numpy(),
                                      exp.index.get_level_values(i).to_numpy(), err_msg=f"group keys level {i}")
    assert list(got.index.names) == list(exp.index.names)
    for c in exp.columns:
        g, e = got[c].to_numpy(), exp[c].to_numpy()
        assert g.dtype == e.dtype, (c, g.dtype, e.dtype)
        func = c[-1] if isinstance(c, tuple) else None
        if e.dtype.kind in "iu" or func in ("count", "min", "max"):
            np.testing.assert_array_equal(g, e, err_msg=f"column {c} must be bit-exact")
        else:
            np.testing.assert_allclose(g, e, rtol=rtol, atol=0, equal_nan=True, err_msg=f"column {c}")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
