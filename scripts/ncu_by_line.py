#!/usr/bin/env python3
"""Attributes the stall samples / executed instructions of an ncu capture to CUDA source lines, by zipping
the report's SASS listing with `nvdisasm -g` of the same cubin (the CSV export of ncu's CUDA view carries
no metrics).  Usage: scripts/ncu_by_line.py REPORT.ncu-rep LIB.so KERNEL_MANGLED_SUBSTRING [top N]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, key = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
sass = ""
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if key in out:
            sass = out
lines = sass.split("\n")
# isolate the kernel's section; out-of-line device functions follow in the same section
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and key in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("\t.section")), len(lines))
seq, cur, inl = [], ("?", 0), ""
for l in lines[start:end]:
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(2)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
recs = []
for r in rows[2:]:
    if len(r) < len(h):
        continue
    try:
        recs.append((int(r[ix["Address"]], 16), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r[ix["Source"]].strip()))
    except ValueError:
        pass
base = recs[0][0]
by_off = {off: (cur, txt) for off, cur, txt in seq}
agg_s, agg_n = collections.Counter(), collections.Counter()
miss = 0
for addr, s, n, txt in recs:
    k = by_off.get(addr - base)
    if not k:
        miss += 1
        continue
    agg_s[k[0]] += s; agg_n[k[0]] += n
ts, tn = sum(agg_s.values()), sum(agg_n.values())
print(f"# {len(recs)} SASS rows, {miss} unmatched; {ts} samples, {tn} warp instructions")
srcs = {}
for (f, ln), s in agg_s.most_common(top):
    if f not in srcs:
        for root in ("polymutt_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                srcs[f] = open(p).read().split("\n"); break
        else:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
    print(f"{100*s/ts:5.1f}% samples {100*agg_n[(f,ln)]/tn:5.1f}% instr  {f}:{ln:<5d} {text}")
