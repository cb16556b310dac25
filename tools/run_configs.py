#!/usr/bin/env python
"""The BASELINE.json configs through the operand API on ONE B200 (chunks resident in HBM).

    python tools/run_configs.py [--scale 0.125] [--configs cfg1,cfg1s,cfg3,cfg4,cfg5] [--no-parity]

Inputs are the host-seeded generators of tools/baseline_configs.py (numpy default_rng(20240229 + chunk), SURVEY.md
section 8d), uploaded before the timed region; every run ends with the oracle on a sample of the groups plus
size-independent invariants (tools/config_runner.check_parity).  One JSON line per config: rows/s over the
operand chain, fetch time of the result frame, algorithmic-bytes roofline fraction, plan, parity.
`--scale` multiplies the number of rows of cfg3 / cfg4 / cfg5 (whole chunks; the chunk size, key space,
columns and functions stay those of the config); cfg1 always runs at its full size.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import baseline_configs as bc  # noqa: E402
import config_runner  # noqa: E402


def run(cfg, scale, repeats, parity=True):
    c = bc.CONFIGS[cfg]
    if cfg.startswith("cfg1") or cfg == "cfg2":
        rows = c.rows
    else:
        rows = max(1, int(round(len(c.chunk_sizes()) * scale))) * c.chunk_rows
    rec = config_runner.run_record(cfg, rows, c.chunk_rows, repeats=repeats, parity=parity)
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.125)
    ap.add_argument("--configs", default="cfg1,cfg1s,cfg3,cfg4,cfg5")
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    for name in a.configs.split(","):
        run(name, a.scale, a.repeats, parity=not a.no_parity)
