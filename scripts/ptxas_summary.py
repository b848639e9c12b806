#!/usr/bin/env python3
"""Registers / stack / spills per site kernel from the ptxas logs of the last build (polymutt_b200/lib/obj/*.ptxas.log)."""
import glob, re, subprocess, sys
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else "k_sites|k_post|k_compact")
for f in sorted(glob.glob("polymutt_b200/lib/obj/*.ptxas.log")):
    s = open(f).read()
    for e in re.split(r"ptxas info\s+: Compiling entry function '", s)[1:]:
        name = e.split("'")[0]
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        m = re.search(r"(\w+<[^>]*>|\w+)\(", dem)
        short = m.group(1) if m else dem[:60]
        if not pat.search(short):
            continue
        used = re.search(r"Used (\d+) registers", e)
        sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", e)
        sm = re.search(r"(\d+) bytes smem", e)
        print(f"{short:48s} regs {used.group(1):>3s}  stack {sp.group(1):>5s}  spill st/ld {sp.group(2):>5s}/{sp.group(3):<5s} static smem {sm.group(1) if sm else 0}")
