#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with nvdisasm line info and print the hottest
source lines of a kernel.  usage: hotlines.py <src.csv> <cubin> <kernel-substring> [top]"""
import csv, re, subprocess, sys
from collections import defaultdict

src_csv, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# collect (line annotation) per instruction of the function
lines = []
infn = False
cur = ("?", 0)
for l in dis:
    if l.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", l):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
i_s, i_n, i_x = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
inst = [r for r in rows[2:] if len(r) == len(hdr)][: len(lines)]
print(f"# {len(inst)} SASS rows in csv, {len(lines)} instructions with line info")
agg = defaultdict(lambda: [0, 0])
tot = 0
for k, r in enumerate(inst):
    key = lines[k] if k < len(lines) else ("?", 0)
    s = int(r[i_s] or 0)
    agg[key][0] += s
    agg[key][1] += int(r[i_n] or 0)
    tot += s
tot_i = sum(v[1] for v in agg.values())
print(f"# total samples {tot}, warp instructions {tot_i}")
for key, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]}:{key[1]:<5d} samples {100.0*s/max(tot,1):5.1f}%  inst {100.0*n/max(tot_i,1):5.1f}%")
