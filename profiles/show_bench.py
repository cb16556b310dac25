"""Print the headline and the per-config sub-records of a bench.py JSON line.  usage: python profiles/show_bench.py <log>"""
import json
import sys

l = [x for x in open(sys.argv[1]) if x.startswith("{")][-1]
d = json.loads(l)
print("value %.4g rows/s  ms/step %.3f  e2e %.4g  launches %s  step_hbm_frac %.3f  kernel frac %.3f (traffic %.3f)" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["step_hbm_frac"],
    d["roofline"]["frac"], d["roofline"].get("frac_on_traffic") or float("nan")))
print("kernel_time_ms", d.get("kernel_time_ms"), "per_operand %.4g" % d["per_operand"]["value"])
for k, v in (d.get("configs") or {}).items():
    if "error" in v:
        print(k, v)
        continue
    print(k, "rows", v["rows"], "groups", v["groups"], "value %.3g" % v["value"], "frac %.3f" % v["roofline"]["frac"],
          "launches", v["gpu_launches"], "parity", v["parity"]["ok"], ["%.2f ms" % (1e3 * s) for s in v["seconds"]])
    print("    ", {a: round(b["ms_total"], 3) for a, b in (v["kernel_time_ms"] or {}).items()})
if d.get("shuffle"):
    s = d["shuffle"]
    print("shuffle", {k: s.get(k) for k in ("value", "ms_per_step", "rows", "groups", "sum_of_counts_eq_rows", "error")}, s.get("link"), s.get("parity"))
for k, v in (d.get("shuffle_configs") or {}).items():
    print("shuffle", k, {a: v.get(a) for a in ("value", "ms_per_step", "rows", "groups", "sum_of_counts_eq_rows", "error")},
          (v.get("link") or {}).get("GBps_per_gpu_per_direction"), (v.get("parity") or {}).get("ok"), v.get("kernel_time_ms_rank0"))
