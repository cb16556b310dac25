"""Drop-in check against the REAL reference operands (build container only: needs the shimmed reference
build of oracle/build_ref.sh; skipped on the GPU box, where /root/reference does not exist).

The reference's own chunk graph (its tiling, its ReductionCompiler output, its chunk metas) is executed with
the executors of this package registered through ``Operand.register_executor`` -- the seam of SURVEY.md
section 8b -- and the result is compared with the reference's own pandas path.  Runs on CPU with the test
double of the kernels module, so what is verified here is the operand-protocol side of the executors: that
they understand the real ``op.pre_funcs / agg_funcs / post_funcs`` callables, the real shuffle keys and chunk
metas.  The CUDA side is verified by the -m gpu tests against the same golden outputs."""
from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

import fake_kernels
from conftest import assert_frames_match

import ref_runner as _rr

REF_DIR = _rr.REF_DIR
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_DIR, "mars")),
                                reason="shimmed reference build not present (build container only)")


@pytest.fixture()
def ref(monkeypatch):
    import ref_runner

    ref_runner._import_reference()
    from mars_b200 import executors, plugin

    monkeypatch.setattr(executors, "_kernels", fake_kernels)
    monkeypatch.setattr(executors, "_device", lambda: torch.device("cpu"))
    plugin.register_on_mars(gpu_only=False)
    yield ref_runner
    plugin.unregister_on_mars()


CASES = [
    ("key", ["sum", "mean", "count", "min", "max", "var"], "tree", True),
    ("key", ["sum", "mean", "count", "min", "max", "var"], "shuffle", False),
    (["k1", "k2"], ["min", "max", "var"], "shuffle", False),
    ("key", ["sum", "mean", "count", "min", "max", "var"], "shuffle", True),       # PSRS range shuffle
    (["k1", "k2"], ["sum", "count"], "shuffle", True),
    (["k1", "k2"], {"v0": ["sum", "mean"], "v1": ["count"]}, "tree", True),
    ("key", "sum", "tree", True),
]


@pytest.mark.parametrize("by,func,method,sort", CASES)
def test_real_reference_graph_runs_on_our_executors(ref, by, func, method, sort):
    rng = np.random.default_rng(7)
    n = 2000
    raw = pd.DataFrame({"key": rng.integers(-20, 20, n), "k1": rng.integers(0, 6, n), "k2": rng.integers(0, 4, n),
                        "v0": rng.normal(10, 3, n), "v1": rng.random(n)})
    cols = (["key"] if by == "key" else by) + ["v0", "v1"]
    raw = raw[cols]
    fake_kernels.calls.clear()
    out = ref.run_reference(raw, by, func, 250, method=method, sort=sort)
    assert "groupby_aggregate" in fake_kernels.calls, "the registered executors did not run"
    exp = raw.groupby(by, sort=True).agg(func)
    got = out["result"]
    if isinstance(exp, pd.Series):
        exp = exp.to_frame()
    assert_frames_match(got.sort_index(), exp, sort_index=False)
    ops = {(r["op"], r["stage"]) for r in out["records"]}
    assert ("DataFrameGroupByAgg", "map") in ops and ("DataFrameGroupByAgg", "agg") in ops
    if method == "shuffle" and not sort:
        assert ("DataFrameGroupByOperand", "map") in ops and ("DataFrameGroupByOperand", "reduce") in ops
    if method == "shuffle" and sort:
        assert ("DataFrameGroupbySortShuffle", "map") in ops and ("DataFrameGroupbySortShuffle", "reduce") in ops
        assert "range_partition_ids" in fake_kernels.calls
        assert got.index.is_monotonic_increasing          # globally sorted without a final sort_index


def test_unsupported_inputs_fall_back_to_the_reference_path(ref):
    # object (string) keys are outside the GPU envelope: the plugin must hand the op to the reference executor
    raw = pd.DataFrame({"key": list("abcab") * 40, "v": np.arange(200.0)})
    out = ref.run_reference(raw, "key", ["sum", "count"], 50, method="tree", sort=True)
    exp = raw.groupby("key").agg(["sum", "count"])
    pd.testing.assert_frame_equal(out["result"], exp)


def test_three_keys_fall_back_to_the_reference_path(ref):
    """more group keys than the library holds (MGB_MAX_KEYS): NotImplementedError before any data is touched
    -> the reference's own executor"""
    rng = np.random.default_rng(3)
    n = 600
    raw = pd.DataFrame({"a": rng.integers(0, 3, n), "b": rng.integers(0, 3, n), "c": rng.integers(0, 2, n),
                        "v": rng.random(n)})
    fake_kernels.calls.clear()
    out = ref.run_reference(raw, ["a", "b", "c"], ["sum", "count"], 150, method="tree", sort=True)
    assert "groupby_aggregate" not in fake_kernels.calls and "preagg_dense_async" not in fake_kernels.calls
    pd.testing.assert_frame_equal(out["result"], raw.groupby(["a", "b", "c"]).agg(["sum", "count"]))


def test_more_than_64_accumulators_fall_back_to_the_reference_path(ref):
    rng = np.random.default_rng(4)
    n = 400
    cols = {"key": rng.integers(0, 5, n)}
    for i in range(23):                      # min / max / var = 5 atomic steps x 23 columns = 115 accumulators
        cols[f"v{i}"] = rng.normal(10, 3, n)
    raw = pd.DataFrame(cols)
    fake_kernels.calls.clear()
    out = ref.run_reference(raw, "key", ["min", "max", "var"], 100, method="tree", sort=True)
    # the map stage (115 accumulators in one call) went to the reference executor: its outputs are pandas
    # partial frames; the merges split the slots into groups of <= 64 accumulators and do run on our executors
    maps = [r for r in out["records"] if r["op"] == "DataFrameGroupByAgg" and r["stage"] == "map"]
    assert maps and all(isinstance(f, pd.DataFrame) for r in maps for f in out["ctx"][r["key"]])
    pd.testing.assert_frame_equal(out["result"], raw.groupby("key").agg(["min", "max", "var"]), rtol=1e-9)


def test_custom_reduction_nunique_falls_back_to_the_reference_path(ref):
    """a12: custom reductions (only `nunique` uses the registry, groupby/custom_aggregation.py:20-86) are outside
    the six functions: every stage of the operand goes to the reference executor"""
    rng = np.random.default_rng(5)
    n = 800
    raw = pd.DataFrame({"key": rng.integers(0, 9, n), "v": rng.integers(0, 20, n)})
    fake_kernels.calls.clear()
    out = ref.run_reference(raw, "key", ["nunique"], 200, method="tree", sort=True)
    assert "groupby_aggregate" not in fake_kernels.calls
    exp = raw.groupby("key").agg(["nunique"])
    pd.testing.assert_frame_equal(out["result"].sort_index(), exp, check_dtype=False)
