"""Per-region instruction histogram from an `ncu --page source --csv --print-source sass` export.
Usage: sass_hist.py file.csv  -> prints every SASS line with executed count (thousands) and the
opcode histogram weighted by executed warp instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = 0
hist = collections.Counter()
verbose = len(sys.argv) > 2
for n, r in enumerate(rows[2:]):
    ex = int(r[iex]); tot += ex
    op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
    hist[op.split(".")[0]] += ex
    if verbose: print("%4d %9d %6s  %s" % (n, ex, r[ismp], r[isrc].strip()))
print("total warp instructions", tot)
for op, c in hist.most_common(40): print("  %-10s %12d  %5.1f%%" % (op, c, 100.0 * c / tot))
