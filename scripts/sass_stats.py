#!/usr/bin/env python3
"""SASS statistics of one kernel of the built library (no GPU needed):
  scripts/sass_stats.py <regex on the demangled kernel name> [--dump FILE] [--lib PATH]
Prints instruction count, opcode histogram (top 25), local-memory (spill) instructions, calls and barriers."""
import collections
import re
import subprocess
import sys

lib = "polymutt_b200/lib/libpolymutt_b200.so"
args = sys.argv[1:]
dump = None
if "--lib" in args:
    i = args.index("--lib"); lib = args[i + 1]; del args[i:i + 2]
if "--dump" in args:
    i = args.index("--dump"); dump = args[i + 1]; del args[i:i + 2]
pat = re.compile(args[0])
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
names = list(funcs)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
for name, d in zip(names, dem):
    short = re.sub(r"\(.*", "", d)
    if not pat.search(short):
        continue
    ins = funcs[name]
    ops = collections.Counter()
    for _, t in ins:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    print(f"== {short}: {len(ins)} instructions")
    print("   " + "  ".join(f"{k}={v}" for k, v in ops.most_common(25)))
    print(f"   local: STL={ops['STL']} LDL={ops['LDL']}  CALL={ops['CALL']}  BAR={ops['BAR']}  fp64={ops['DFMA']+ops['DMUL']+ops['DADD']+ops['DSETP']+ops['DMNMX']}")
    if dump:
        with open(dump, "w") as f:
            for a, t in ins:
                f.write(f"{a:06x} {t}\n")
