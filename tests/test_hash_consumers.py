"""The other consumers of `hash_dataframe_on` that share the hash + partition kernels:
plain groupby shuffles on `by` columns (mars/dataframe/groupby/core.py:353-485) and the merge shuffle
(mars/dataframe/merge/merge.py:93-147).  Every reducer's frame must hold exactly the rows -- labels, order and
values -- that the reference's `df.iloc[hash_dataframe_on(df, on, size)[i]]` blocks give after the row concat of
the reduce stage.  The same test body runs on the CPU double of the kernels (host logic) and, marked gpu, on
libmarsgb."""
from __future__ import annotations

import numpy as np
import pandas as pd
import pytest
import torch

import fake_kernels
import mars_groupby_oracle as orc


def _expected(chunks, on, R):
    """reference semantics: per reducer the concat (mapper order) of df.iloc[filter_i]"""
    out = []
    for r in range(R):
        parts = [c.iloc[orc.hash_dataframe_on_columns(c, on, R)[r]] for c in chunks]
        out.append(pd.concat(parts, axis=0))
    return out


def _graph(md, chunks, kind, on, R):
    from mars_b200.core import OperandStage
    from mars_b200.dataframe import DataFrameGroupByOperand, DataFrameMergeAlign, DataFrameShuffleProxy

    src = md.DataFrame.from_chunks(chunks).to_gpu()
    maps = []
    for c in src.chunks:
        if kind == "groupby":
            op = DataFrameGroupByOperand(stage=OperandStage.map, by=list(on), shuffle_size=R, gpu=True)
        else:
            op = DataFrameMergeAlign(stage=OperandStage.map, shuffle_on=list(on), index_shuffle_size=R, mapper_id=0, gpu=True)
        maps.append(op.new_chunk([c], shape=(np.nan, np.nan), index=c.index))
    proxy = DataFrameShuffleProxy(gpu=True).new_chunk(maps, shape=())
    reds = []
    for r in range(R):
        if kind == "groupby":
            op = DataFrameGroupByOperand(stage=OperandStage.reduce, n_reducers=R, gpu=True)
        else:
            op = DataFrameMergeAlign(stage=OperandStage.reduce, n_reducers=R, gpu=True)
        reds.append(op.new_chunk([proxy], shape=(np.nan, np.nan), index=(r, 0)))
    return reds


def _run(md, chunks, kind, on, R):
    from mars_b200.core import ProcessorContext

    sess = md.new_session(default=False)
    reds = _graph(md, chunks, kind, on, R)
    ctx = ProcessorContext()
    sess._run_chunks_local(reds, ctx, {})
    return [ctx[c.key].to_pandas() for c in reds]


def _check(md, kind, on):
    rng = np.random.default_rng(len(on) + (kind == "merge"))
    n, R = 4000, 5
    raw = pd.DataFrame({"a": rng.integers(-50, 50, n), "b": rng.integers(0, 7, n), "v": rng.random(n),
                        "w": rng.integers(0, 10 ** 6, n)})
    chunks = [raw.iloc[i:i + 900] for i in range(0, n, 900)]
    got = _run(md, chunks, kind, on, R)
    exp = _expected(chunks, on, R)
    assert sum(len(g) for g in got) == n
    for g, e in zip(got, exp):
        pd.testing.assert_frame_equal(g, e)        # same rows, same labels, same order inside every block


@pytest.fixture()
def md_cpu(monkeypatch):
    import mars_b200 as m
    from mars_b200 import executors

    monkeypatch.setattr(executors, "_kernels", fake_kernels)
    monkeypatch.setattr(executors, "_device", lambda: torch.device("cpu"))
    return m


@pytest.mark.parametrize("kind", ["groupby", "merge"])
@pytest.mark.parametrize("on", [["a"], ["a", "b"]])
def test_column_shuffles_host_logic(md_cpu, kind, on):
    _check(md_cpu, kind, on)


def test_frame_row_hash_known_answers():
    """pandas: hash_pandas_object(df[on], index=False) combines the column hashes also for ONE column"""
    k = np.array([1, 2, 3, -5, 2 ** 40], dtype=np.int64)
    exp = pd.util.hash_pandas_object(pd.DataFrame({"a": k}), index=False, categorize=False).to_numpy()
    np.testing.assert_array_equal(orc.hash_frame_rows([k]), exp)
    assert not np.array_equal(orc.hash_frame_rows([k]), orc.hash_keys([k]))     # an Index is hashed directly
    k2 = np.array([7, 8, 9, 10, 11], dtype=np.int64)
    exp2 = pd.util.hash_pandas_object(pd.DataFrame({"a": k, "b": k2}), index=False, categorize=False).to_numpy()
    np.testing.assert_array_equal(orc.hash_frame_rows([k, k2]), exp2)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["groupby", "merge"])
@pytest.mark.parametrize("on", [["a"], ["a", "b"]])
def test_column_shuffles_on_the_gpu(cuda_device, kind, on):
    import mars_b200 as m

    _check(m, kind, on)


@pytest.mark.gpu
@pytest.mark.parametrize("n_keys", [1, 2])
def test_partition_ids_frame_match_pandas(cuda_device, n_keys):
    from mars_b200 import kernels as K

    rng = np.random.default_rng(n_keys)
    n = 300_000
    keys = [rng.integers(-10 ** 12, 10 ** 12, n, dtype=np.int64) for _ in range(n_keys)]
    got = K.partition_ids_frame([torch.from_numpy(k).cuda() for k in keys], 37).cpu().numpy()
    df = pd.DataFrame({f"k{j}": k for j, k in enumerate(keys)})
    exp = (pd.util.hash_pandas_object(df, index=False, categorize=False).to_numpy() % np.uint64(37)).astype(np.int32)
    np.testing.assert_array_equal(got, exp)
