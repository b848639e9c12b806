#!/usr/bin/env python3
"""Per-phase cycle breakdown of the wide kernel (library built with PM_DEFS=-DPM_PHASE_TIMING)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from polymutt_b200 import Engine, Params, capi, synth
ped = synth.trios(1000); n = 32768
dev = torch.device("cuda", 0)
eng = Engine(ped, Params(denovo=True))
if len(sys.argv) > 2: eng.force_wide_plan(int(sys.argv[1]), int(sys.argv[2]))
hdr = torch.empty((n, 8), dtype=torch.uint8, device=dev); recs = torch.empty((n, ped.n_person, 16), dtype=torch.uint8, device=dev)
synth.generate_sites(ped, n, seed=5, device=dev, out_hdr=hdr, out_recs=recs, chunk=4096)
status = torch.empty(n, dtype=torch.uint16, device=dev); res = torch.empty((n, 256), dtype=torch.uint8, device=dev)
per = torch.empty((2048, ped.n_person, 96), dtype=torch.uint8, device=dev); nres = torch.zeros(1, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
step = lambda: eng.call_glf_sites_device(hdr.data_ptr(), recs.data_ptr(), n, capi.PM_OUT_EMITTED, status.data_ptr(), res.data_ptr(), per.data_ptr(), 2048, nres.data_ptr())
step(); eng.sync(); eng.reset_counters(); step(); eng.sync()
c = eng.counters(); out = np.zeros(8, dtype=np.uint64)
eng.lib.pm_debug_phase_cycles.argtypes = [C.c_void_p, C.c_void_p]
eng.lib.pm_debug_phase_cycles(eng.ctx, out.ctypes.data)
names = ["tma_wait", "stats", "setup", "spec_resolve", "generic_brent_step", "decide_write", "spec_eval_reduce", "generic_eval_reduce"]
sites = c["sites_evaluated"]; tot = float(out[:8].sum())
print(eng.describe_plan()); print("main kernel ms", eng.last_timing()[0], "evals/site", c["evaluations"] / sites)
for k, nm in enumerate(names):
    print(f"{nm:16s} {out[k] / sites:10.0f} cycles/site  {100 * out[k] / tot:5.1f}%")
print("total cycles/site (thread 0 of a block)", tot / sites, " per evaluation:", {names[k]: round(out[k] / c["evaluations"]) for k in (6, 3, 7, 4)})
