"""Opcode mix of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view):
warp-level executed instructions per opcode, and shared-memory wavefronts per LDS/STS."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
ops, wave, samples = Counter(), Counter(), Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ci["Source"]].strip()
    toks = src.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.rstrip(";")
    n = int(float(r[ci["Instructions Executed"]] or 0))
    ops[op] += n
    tot += n
    samples[op] += int(float(r[ci["# Samples"]] or 0))
    w = r[ci["L1 Wavefronts Shared"]]
    if w:
        wave[op] += int(float(w))
print("total warp instructions: %d" % tot)
ns = sum(samples.values())
for op, n in ops.most_common(40):
    print("%-28s %12d %6.2f%%  samples %5.2f%%  shared wavefronts %d" % (op, n, 100.0 * n / tot, 100.0 * samples[op] / max(ns, 1), wave[op]))
