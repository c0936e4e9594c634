#!/usr/bin/env python
"""Coarse phase breakdown (time = warp-state samples, instructions) of a kernel from an ncu report.
usage: phases.py <report> <kernel-regex> <cubin> <mangled-substring> <file:lo-hi=name> ..."""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, kre, cubin, mangled = sys.argv[1:5]
rules = []
for spec in sys.argv[5:]:
    loc, name = spec.split("=")
    f, rng = loc.split(":")
    lo, hi = rng.split("-")
    rules.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
import os
want = os.environ.get("KERNEL_SUBSTR", "")
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and want in r[1]]
first = starts[0]
nxt = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and i > first]
rows = rows[first:(nxt[0] if nxt else len(rows))]
hi_ = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi_]
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
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
inst = [r for r in rows[hi_ + 1:] if len(r) == len(hdr)][: len(lines)]
i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")
agg = defaultdict(lambda: [0, 0])
for k, r in enumerate(inst):
    f, n = lines[k]
    name = f
    for rf, lo, hi, nm in rules:
        if f == rf and lo <= n <= hi:
            name = nm
            break
    agg[name][0] += int(r[i_s] or 0)
    agg[name][1] += int(r[i_n] or 0)
ts = sum(v[0] for v in agg.values()) or 1
tn = sum(v[1] for v in agg.values()) or 1
print(f"# {len(inst)} SASS instructions, {ts} samples, {tn} warp instructions")
for k, (s, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:34s} time {100*s/ts:5.1f}%   instructions {100*c/tn:5.1f}%")
