#!/usr/bin/env python3
"""Compact per-instruction view of an ncu source-page CSV: for a range of SASS rows, prints address, stall samples and
the opcode; rows executed fewer than --min-exec times are dropped.  Usage: ncu_hot.py CSV [--from PATTERN] [--rows N] [--min-exec K]"""
import csv, sys, re
args = sys.argv[1:]
path = args[0]
def opt(name, default):
    return type(default)(args[args.index(name) + 1]) if name in args else default
start_pat, nrows, min_exec, occ = opt("--from", "BAR.SYNC"), opt("--rows", 400), opt("--min-exec", 1), opt("--occurrence", 1)
rows = list(csv.reader(open(path)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
recs = []
for r in rows[2:]:
    if len(r) < len(h): continue
    try: recs.append((int(r[ix["Address"]], 16), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r[ix["Source"]].strip()))
    except ValueError: pass
base = recs[0][0]; tot = sum(s for _, s, _, _ in recs)
hits = [i for i, r in enumerate(recs) if re.search(start_pat, r[3]) and r[2] >= min_exec]
i0 = hits[occ - 1] if len(hits) >= occ else 0
acc = 0
out = []
for a, s, n, t in recs[i0:i0 + nrows]:
    if n < min_exec: continue
    acc += s
    out.append(f"{a - base:05x} {s:5d} {n // 1000:6d}k {re.sub(r'\s+', ' ', t)[:44]}")
print(f"# total samples {tot}; shown {acc} ({100 * acc / tot:.1f}%)")
# three columns
k = (len(out) + 2) // 3
for i in range(k):
    print(" | ".join((out[j] if j < len(out) else "").ljust(66) for j in (i, i + k, i + 2 * k)))
