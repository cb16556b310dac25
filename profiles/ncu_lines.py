"""Join an ncu source-page CSV (SASS view, possibly several kernels) with nvdisasm -g line info: stall samples and
warp instructions per source line.
usage: python profiles/ncu_lines.py <src.csv> <nvdisasm -g output> <mangled kernel name> <substring of the kernel's
display name in the csv> [top]"""
import csv
import re
import sys
from collections import Counter

src_csv, dis, target, display = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
fn = line = None
addr2line = {}
for l in open(dis):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and fn == target:
        addr2line[int(m.group(1), 16)] = line
samp, inst = Counter(), Counter()
cur, ci, k = False, None, 0
for r in csv.reader(open(src_csv)):
    if r and r[0] == "Kernel Name":
        cur, k = display in r[1], 0
        continue
    if r and r[0] == "Address":
        ci = {h: i for i, h in enumerate(r)}
        continue
    if not cur or ci is None or len(r) < 8:
        continue
    ln = addr2line.get(k * 16)
    k += 1
    samp[ln] += int(r[ci["# Samples"]] or 0)
    inst[ln] += int(r[ci["Instructions Executed"]] or 0)
tot, toti = sum(samp.values()), sum(inst.values())
print(f"samples {tot}, warp instructions {toti}")
for ln, s in samp.most_common(top):
    print(ln, s, f"{100 * s / max(tot, 1):.1f}%", inst[ln], f"{100 * inst[ln] / max(toti, 1):.1f}%")
