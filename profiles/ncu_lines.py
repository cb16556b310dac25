"""Join an ncu source-page CSV (SASS view) with nvdisasm -g line info: samples / instructions per source line.
usage: python profiles/ncu_lines.py <src.csv> <nvdisasm -g output> <mangled kernel name> [top]"""
import csv
import re
import sys
from collections import Counter

src_csv, dis, target = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
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
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
samp, inst = Counter(), Counter()
for i, r in enumerate(rows[2:]):
    ln = addr2line.get(i * 16)
    samp[ln] += int(r[ci["# Samples"]] or 0)
    inst[ln] += int(r[ci["Instructions Executed"]] or 0)
tot, toti = sum(samp.values()), sum(inst.values())
print(f"samples {tot}, warp instructions {toti}")
for ln, s in samp.most_common(top):
    print(ln, s, f"{100 * s / tot:.1f}%", inst[ln])
