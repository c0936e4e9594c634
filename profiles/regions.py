#!/usr/bin/env python
"""Per-source-line time (warp-state samples) and instruction shares of one kernel from an ncu report.
usage: regions.py <report.ncu-rep> <kernel-regex> <cubin> <mangled-substring> [top]"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, kre, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
import os
want = os.environ.get("KERNEL_SUBSTR", "")
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and want in r[1]]
first = starts[0]
nxt = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and i > first]
rows = rows[first:(nxt[0] if nxt else len(rows))]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, infn, cur = [], False, ("?", 0)
for l in dis:
    if l.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", l):
        infn = mangled in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
inst = [r for r in rows[hi + 1:] if len(r) == len(hdr)][: len(lines)]
i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")
agg = defaultdict(lambda: [0, 0])
for k, r in enumerate(inst):
    agg[lines[k]][0] += int(r[i_s] or 0)
    agg[lines[k]][1] += int(r[i_n] or 0)
ts = sum(v[0] for v in agg.values()) or 1
tn = sum(v[1] for v in agg.values()) or 1
print(f"# {len(inst)} instructions, {ts} samples, {tn} warp instructions")
cache = {}
def text(f, n):
    if f not in cache:
        try:
            cache[f] = open(f).read().splitlines()
        except OSError:
            cache[f] = []
    L = cache[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
order = 1 if len(sys.argv) > 6 and sys.argv[6] == "inst" else 0
for (f, n), (s, c) in sorted(agg.items(), key=lambda kv: -kv[1][order])[:top]:
    print(f"{100*s/ts:5.1f}% t {100*c/tn:5.1f}% i  {f.split('/')[-1]}:{n:<4d} {text(f, n)}")
