"""The five BASELINE.json configs, concretised as SURVEY.md section 8(d) specifies them.

Shared by ``bench.py`` (sub-records per config), ``tests/test_baseline_configs.py`` (oracle parity at scaled
row counts) and ``tools/run_configs.py``.  Inputs are generated on the HOST with
``np.random.default_rng(SEED + global_chunk_index)`` per chunk, so the CPU oracle and the GPU path see
identical bytes; nothing here touches CUDA.

Scaling rule ("scaled rows, exact shape"): a scaled run keeps the columns, dtypes, key distributions (the
key SPACE is never scaled: cfg3 draws from 1e8 keys, cfg4 from 1000 x 1000, cfg5 from Zipf(1.2) capped at
1e8), functions, sort / method flags; only the number of rows and -- where the caller asks for it -- the
chunk size shrink.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np

SEED = 20240229

Q1_COLUMNS = ["returnflag", "linestatus", "quantity", "extendedprice", "discount", "disc_price", "charge", "orderkey"]
Q1_FUNC = {"quantity": ["sum", "mean"], "extendedprice": ["sum", "mean"], "disc_price": ["sum"],
           "charge": ["sum"], "discount": ["mean"], "orderkey": ["count"]}


def _gen_cfg1(chunk: int, rows: int) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(SEED + chunk)
    out = {"key": rng.integers(0, 1000, rows, dtype=np.int64)}
    for i in range(4):
        out[f"v{i}"] = rng.random(rows)
    return out


def _gen_cfg2(chunk: int, rows: int) -> Dict[str, np.ndarray]:
    """Q1-shaped synthetic lineitem chunk: keys drawn jointly with p(AF, NF, NO, RF), dictionary-encoded"""
    rng = np.random.default_rng(SEED + chunk)
    grp = rng.choice(4, size=rows, p=[0.2466, 0.0006, 0.5063, 0.2465])      # AF, NF, NO, RF
    returnflag = np.array([0, 1, 1, 2], dtype=np.int64)[grp]
    linestatus = np.array([0, 0, 1, 0], dtype=np.int64)[grp]
    quantity = rng.integers(1, 51, rows).astype(np.float64)
    price = rng.uniform(900.0, 105000.0, rows)
    discount = rng.integers(0, 11, rows) / 100.0
    tax = rng.integers(0, 9, rows) / 100.0
    disc_price = price * (1.0 - discount)
    charge = disc_price * (1.0 + tax)
    orderkey = rng.integers(1, 60_000_000, rows, dtype=np.int64)
    return dict(returnflag=returnflag, linestatus=linestatus, quantity=quantity, extendedprice=price,
                discount=discount, disc_price=disc_price, charge=charge, orderkey=orderkey)


def _gen_cfg3(chunk: int, rows: int) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(SEED + chunk)
    return {"key": rng.integers(0, 100_000_000, rows, dtype=np.int64), "v": rng.random(rows)}


def _gen_cfg4(chunk: int, rows: int) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(SEED + chunk)
    out = {"k1": rng.integers(0, 1000, rows, dtype=np.int64), "k2": rng.integers(0, 1000, rows, dtype=np.int64)}
    for i in range(8):
        out[f"v{i}"] = rng.normal(10.0, 3.0, rows)      # mean^2 / var ~ 11: the sum-of-squares formula is well conditioned
    return out


def _gen_cfg5(chunk: int, rows: int) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(SEED + chunk)
    key = np.minimum(rng.zipf(1.2, rows), 10 ** 8).astype(np.int64)     # hot key ~ 18 % of the rows
    return {"key": key, "v": rng.random(rows)}


class Config:
    def __init__(self, name, rows, chunk_rows, by, func, sort, method, gen, n_keys, n_vals_read, n_out_cols,
                 approx_groups, note=""):
        self.name, self.rows, self.chunk_rows = name, rows, chunk_rows
        self.by: List[str] = by
        self.func, self.sort, self.method = func, sort, method
        self.gen: Callable[[int, int], Dict[str, np.ndarray]] = gen
        self.n_keys, self.n_vals_read, self.n_out_cols = n_keys, n_vals_read, n_out_cols
        self.approx_groups = approx_groups
        self.note = note

    # ALGORITHMIC HBM bytes: every input column read once + the result written once
    @property
    def bytes_per_row(self) -> int:
        return 8 * (self.n_keys + self.n_vals_read)

    def algorithmic_bytes(self, rows: int, groups: int) -> int:
        return rows * self.bytes_per_row + groups * 8 * (self.n_keys + self.n_out_cols)

    def chunk_sizes(self, rows: Optional[int] = None, chunk_rows: Optional[int] = None) -> List[int]:
        rows = self.rows if rows is None else rows
        chunk_rows = self.chunk_rows if chunk_rows is None else chunk_rows
        return [min(chunk_rows, rows - s) for s in range(0, rows, chunk_rows)]

    def columns(self) -> List[str]:
        return list(self.gen(0, 1).keys())


CONFIGS: Dict[str, Config] = {
    # cfg1: default sort=True / method='auto' -> tree 10 -> 3 -> 1; "cfg1s" is the same data on the named hash path
    "cfg1": Config("cfg1", 10_000_000, 1_000_000, ["key"], ["sum", "mean", "count"], True, "auto", _gen_cfg1,
                   1, 4, 12, 1000),
    "cfg1s": Config("cfg1s", 10_000_000, 1_000_000, ["key"], ["sum", "mean", "count"], False, "shuffle", _gen_cfg1,
                    1, 4, 12, 1000, note="cfg1 data, sort=False + method='shuffle' (R = 10)"),
    "cfg2": Config("cfg2", 59_986_052, 1 << 22, ["returnflag", "linestatus"], Q1_FUNC, True, "auto", _gen_cfg2,
                   2, 6, 8, 4),
    "cfg3": Config("cfg3", 1_000_000_000, 15_625_000, ["key"], ["sum", "count"], False, "shuffle", _gen_cfg3,
                   1, 1, 2, 100_000_000),
    "cfg4": Config("cfg4", 500_000_000, 7_812_500, ["k1", "k2"], ["min", "max", "var"], False, "shuffle", _gen_cfg4,
                   2, 8, 24, 1_000_000),
    "cfg5": Config("cfg5", 1_000_000_000, 15_625_000, ["key"], ["mean", "count"], False, "shuffle", _gen_cfg5,
                   1, 1, 2, 21_000_000),
}


def host_chunks(cfg: Config, rows: int, chunk_rows: Optional[int] = None, first_chunk: int = 0):
    """list of dicts column -> numpy array, chunk c generated with default_rng(SEED + first_chunk + c)"""
    return [cfg.gen(first_chunk + c, n) for c, n in enumerate(cfg.chunk_sizes(rows, chunk_rows))]


def to_pandas(chunks):
    import pandas as pd

    return pd.concat([pd.DataFrame(c, copy=False) for c in chunks], axis=0, ignore_index=True)


def sampled_key_oracle(cfg: Config, chunks, got, n_sample: int = 2048, seed: int = 7):
    """Exact per-group expectations for a random sample of the result's groups, computed from the raw host
    chunks with pandas through the ORACLE's stage functions (one tree pass over the matching rows).  Used by
    bench.py where the full frame is too large to rebuild on the CPU in seconds.  Returns (expected frame,
    got frame restricted to the sampled keys), both key-sorted."""
    import os
    import sys

    import pandas as pd

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import mars_groupby_oracle as orc

    rng = np.random.default_rng(seed)
    take = np.sort(rng.choice(len(got), size=min(n_sample, len(got)), replace=False))
    sample = got.iloc[take]
    if len(cfg.by) == 1:
        keys = sample.index.to_numpy()
        parts = [pd.DataFrame(c, copy=False)[np.isin(c[cfg.by[0]], keys)] for c in chunks]
    else:
        code_s = sample.index.get_level_values(0).to_numpy() * 1_000_003 + sample.index.get_level_values(1).to_numpy()
        parts = [pd.DataFrame(c, copy=False)[np.isin(c[cfg.by[0]] * 1_000_003 + c[cfg.by[1]], code_s)] for c in chunks]
    raw = pd.concat(parts, axis=0, ignore_index=True)
    exp = orc.run_groupby_agg(raw, cfg.by, cfg.func, max(len(raw), 1), method="tree", sort=True)
    return exp, sample.sort_index()
