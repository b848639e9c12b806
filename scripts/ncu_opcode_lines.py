#!/usr/bin/env python3
"""Executed warp instructions of an ncu capture per (source line, opcode), for one opcode prefix.
Usage: scripts/ncu_opcode_lines.py REPORT.ncu-rep LIB.so KERNEL_MANGLED_SUBSTRING OPCODE [top N]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, key, opc = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
sass = ""
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if key in out:
            sass = out
lines = sass.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and key in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("\t.section")), len(lines))
seq, cur = {}, ("?", 0)
for l in lines[start:end]:
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        seq[int(m.group(1), 16)] = (cur, m.group(2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
agg, ops = collections.Counter(), collections.Counter()
base, tot = None, 0
for r in rows[2:]:
    try:
        addr, n = int(r[ix["Address"]], 16), int(r[ix["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    if base is None: base = addr
    k = seq.get(addr - base)
    if not k: continue
    tot += n
    txt = re.sub(r"^@!?U?P\d+\s+", "", k[1])
    if txt.startswith(opc):
        agg[k[0]] += n; ops[txt.split()[0]] += n
print(f"# {opc}*: {sum(agg.values())} of {tot} warp instructions; variants: {dict(ops.most_common(8))}")
for (f, ln), n in agg.most_common(top):
    print(f"{100*n/tot:5.2f}%  {f}:{ln}")
