"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the *unmodified reference*.

Run in the build container only (needs the shimmed reference build, `oracle/build_ref.sh`):

    python oracle/make_golden.py            # rewrites tests/golden/

For every case the reference's own operands are executed chunk by chunk (oracle/ref_runner.py); stored are
the inputs, the final frame, every result chunk (hash shuffle: one per reducer == the partition assignment
of every group key) and the stage=map partial tuples.  While generating, the CPU restatement
(oracle/mars_groupby_oracle.py) is checked against the reference on the same inputs, so a drift between the
two is caught here, in the container that has the reference.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLDEN = os.path.join(HERE, "..", "tests", "golden")

ALL6 = ["sum", "mean", "count", "min", "max", "var"]


def case_inputs(name):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)) + 20240229)
    if name == "single_key_all6":
        n = 1000
        return pd.DataFrame({"key": rng.integers(-25, 25, n), "v0": rng.random(n), "v1": rng.normal(10, 3, n),
                             "i0": rng.integers(-1000, 1000, n)})
    if name == "two_keys_minmaxvar":
        n = 1200
        return pd.DataFrame({"k1": rng.integers(0, 12, n), "k2": rng.integers(-5, 5, n),
                             "v0": rng.normal(10, 3, n), "v1": rng.normal(-4, 2, n), "v2": rng.random(n)})
    if name == "q1_dict":
        n = 1500
        price = rng.uniform(900, 105000, n)
        disc = rng.integers(0, 11, n) / 100.0
        tax = rng.integers(0, 9, n) / 100.0
        return pd.DataFrame({"returnflag": rng.integers(0, 3, n), "linestatus": rng.integers(0, 2, n),
                             "quantity": rng.integers(1, 51, n).astype("float64"), "extendedprice": price,
                             "discount": disc, "tax": tax, "disc_price": price * (1 - disc),
                             "charge": price * (1 - disc) * (1 + tax), "orderkey": rng.integers(1, 10 ** 7, n)})
    if name == "nan_and_extreme_keys":
        n = 800
        keys = rng.choice(np.array([0, -1, 2 ** 63 - 1, -2 ** 63, 7, 1 << 40, -(1 << 40)], dtype=np.int64), n)
        v0 = rng.random(n)
        v0[rng.random(n) < 0.3] = np.nan
        v1 = rng.normal(0, 1, n)
        v1[keys == 7] = np.nan  # an all-NaN group
        return pd.DataFrame({"key": keys, "v0": v0, "v1": v1})
    if name == "cfg1_small":
        n = 20000
        return pd.DataFrame({"key": rng.integers(0, 1000, n), "v0": rng.random(n), "v1": rng.random(n),
                             "v2": rng.random(n), "v3": rng.random(n)})
    if name == "high_card":
        n = 6000
        return pd.DataFrame({"key": rng.integers(0, 100000, n), "v": rng.random(n)})
    raise KeyError(name)


Q1_FUNC = {"quantity": ["sum", "mean"], "extendedprice": ["sum", "mean"], "disc_price": ["sum"],
           "charge": ["sum"], "discount": ["mean"], "orderkey": ["count"]}

CASES = [
    # name, inputs, by, func, chunk_size, method, sort
    ("single_key_all6_tree", "single_key_all6", ["key"], ALL6, 250, "tree", True),
    ("single_key_all6_shuffle", "single_key_all6", ["key"], ALL6, 250, "shuffle", False),
    ("two_keys_minmaxvar_shuffle", "two_keys_minmaxvar", ["k1", "k2"], ["min", "max", "var"], 200, "shuffle", False),
    ("two_keys_minmaxvar_tree", "two_keys_minmaxvar", ["k1", "k2"], ["min", "max", "var"], 200, "tree", True),
    ("q1_dict_tree", "q1_dict", ["returnflag", "linestatus"], Q1_FUNC, 300, "tree", True),
    ("nan_extreme_shuffle", "nan_and_extreme_keys", ["key"], ALL6, 100, "shuffle", False),
    ("nan_extreme_tree", "nan_and_extreme_keys", ["key"], ALL6, 100, "tree", True),
    ("cfg1_small_tree", "cfg1_small", ["key"], ["sum", "mean", "count"], 2000, "tree", True),
    ("cfg1_small_shuffle", "cfg1_small", ["key"], ["sum", "mean", "count"], 2000, "shuffle", False),
    ("high_card_shuffle", "high_card", ["key"], ["sum", "count"], 1000, "shuffle", False),
    # sort=True + shuffle: PSRS range partition (DataFrameGroupbySortShuffle)
    ("single_key_all6_sortshuffle", "single_key_all6", ["key"], ALL6, 250, "shuffle", True),
    ("two_keys_minmaxvar_sortshuffle", "two_keys_minmaxvar", ["k1", "k2"], ["min", "max", "var"], 200, "shuffle", True),
    ("high_card_sortshuffle", "high_card", ["key"], ["sum", "count"], 1000, "shuffle", True),
    ("nan_extreme_sortshuffle", "nan_and_extreme_keys", ["key"], ALL6, 100, "shuffle", True),
]


def frame_to_arrays(prefix, df, out):
    idx = df.index
    levels = [idx.get_level_values(i).to_numpy() for i in range(idx.nlevels)]
    out[f"{prefix}_nlevels"] = np.int64(len(levels))
    for i, lv in enumerate(levels):
        out[f"{prefix}_index{i}"] = lv
    out[f"{prefix}_ncols"] = np.int64(df.shape[1])
    for i in range(df.shape[1]):
        out[f"{prefix}_col{i}"] = df.iloc[:, i].to_numpy()
    labels = [list(c) if isinstance(c, tuple) else c for c in df.columns]
    out[f"{prefix}_meta"] = np.array(json.dumps(dict(columns=labels, index_names=list(idx.names))))


def main():
    from ref_runner import run_reference
    import mars_groupby_oracle as orc

    os.makedirs(GOLDEN, exist_ok=True)
    for name, inp, by, func, chunk_size, method, sort in CASES:
        raw = case_inputs(inp)
        ref = run_reference(raw, by if len(by) > 1 else by[0], func, chunk_size, method=method, sort=sort)
        out = {}
        for c in raw.columns:
            out[f"in_{c}"] = raw[c].to_numpy()
        out["spec"] = np.array(json.dumps(dict(by=by, func=func, chunk_size=chunk_size, method=method, sort=sort,
                                               in_columns=list(raw.columns))))
        frame_to_arrays("result", ref["result"], out)
        out["n_chunk_results"] = np.int64(len(ref["chunk_results"]))
        for i, cr in enumerate(ref["chunk_results"]):
            frame_to_arrays(f"chunk{i}", cr, out)
        # stage=map partial tuples
        maps = [r for r in ref["records"] if r["op"] == "DataFrameGroupByAgg" and r["stage"] == "map"]
        # ctx entries of consumed chunks are still in the runner's dict ctx (nothing is released there)
        out["n_maps"] = np.int64(len(maps))
        for i, r in enumerate(maps):
            tup = ref["ctx"][r["key"]]
            out[f"map{i}_steps"] = np.array(json.dumps(r["step_keys"]))
            for k, df in zip(r["step_keys"], tup):
                frame_to_arrays(f"map{i}_{k}", df, out)
        # check the restatement against the reference while we have it
        rec = {}
        got = orc.run_groupby_agg(raw, by, func, chunk_size, method=method, sort=sort, record=rec)
        exp = ref["result"]
        if not sort:
            got, exp = got.sort_index(), exp.sort_index()
        pd.testing.assert_frame_equal(got, exp, rtol=1e-12, atol=0, check_exact=False)
        assert len(rec["chunk_results"]) == len(ref["chunk_results"])
        for a, b in zip(rec["chunk_results"], ref["chunk_results"]):
            if not sort:
                a, b = a.sort_index(), b.sort_index()
            pd.testing.assert_frame_equal(a, b, rtol=1e-12, atol=0, check_exact=False)
        for i, r in enumerate(maps):
            tup = ref["ctx"][r["key"]]
            mine = rec["map"][i]
            assert len(tup) == len(mine), (name, len(tup), len(mine))
            mine_steps = [s.key for s in orc.compile_steps(func)[0]]
            assert sorted(mine_steps) == sorted(r["step_keys"]), (mine_steps, r["step_keys"])
            for k, b in zip(r["step_keys"], tup):
                a = mine[mine_steps.index(k)]
                # the reference's x**2 frame also carries the squared key columns (dropped later by the
                # post step's common-column join, aggregation.py:1219-1226); dead columns are not restated
                b = b[a.columns]
                pd.testing.assert_frame_equal(a.sort_index(), b.sort_index(), rtol=1e-13, check_names=False)
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: result {ref['result'].shape}, {len(ref['chunk_results'])} result chunks, "
              f"{len(maps)} map partials -> {os.path.getsize(path)} bytes; oracle == reference")


if __name__ == "__main__":
    main()
