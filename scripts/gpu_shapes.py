#!/usr/bin/env python3
"""Device-resident throughput of the site kernels for several pedigree shapes (exploration; not the bench)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from polymutt_b200 import Engine, Params, capi, synth

def run(name, ped, params, n_sites, reps=3, plan=None):
    dev = torch.device("cuda", 0)
    eng = Engine(ped, params)
    if plan:
        eng.force_wide_plan(*plan)
    hdr = torch.empty((n_sites, 8), dtype=torch.uint8, device=dev)
    recs = torch.empty((n_sites, ped.n_person, 16), dtype=torch.uint8, device=dev)
    synth.generate_sites(ped, n_sites, seed=5, device=dev, out_hdr=hdr, out_recs=recs, chunk=max(256, (1 << 22) // ped.n_person))
    cap = n_sites
    status = torch.empty(n_sites, dtype=torch.uint16, device=dev)
    res = torch.empty((cap, 256), dtype=torch.uint8, device=dev)
    per = torch.empty((max(1, cap // 4), ped.n_person, 96), dtype=torch.uint8, device=dev)
    nres = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    def step():
        eng.call_glf_sites_device(hdr.data_ptr(), recs.data_ptr(), n_sites, capi.PM_OUT_EMITTED, status.data_ptr(), res.data_ptr(), per.data_ptr(), cap // 4, nres.data_ptr())
    step(); eng.sync()
    eng.timer_start()
    for _ in range(reps): step()
    ms = eng.timer_stop() / reps
    eng.sync()
    main_ms, total_ms, _ = eng.last_timing()
    out = dict(name=name, persons=ped.n_person, sites=n_sites, sites_per_s=round(n_sites / (ms * 1e-3)), ms=round(ms, 3), main_ms=round(main_ms, 3), post_ms=round(total_ms - main_ms, 3), emitted=int(nres.item()), plan=eng.describe_plan()[:110])
    print(json.dumps(out), flush=True)
    eng.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["all"]
    S = synth
    shapes = {
        "trios1000_dn": (S.trios(1000), Params(denovo=True), 32768),
        "trios1000_ba": (S.trios(1000), Params(), 32768),
        "mixed100_dn": (S.concat(S.trios(50), S.families([4] * 50)), Params(denovo=True), 131072),
        "mixed100_ba": (S.concat(S.trios(50), S.families([4] * 50)), Params(), 131072),
        "fam200x5_ba": (S.families([5] * 200), Params(), 65536),
        "trios20_dn": (S.trios(20), Params(denovo=True), 262144),
        "quartets3_ba": (S.families([4, 4, 4]), Params(), 1 << 20),
        "quartets3_dn": (S.families([4, 4, 4]), Params(denovo=True), 1 << 20),
        "ceph20_ba": (S.ceph(), Params(), 1 << 19),
        "ceph20_dn": (S.ceph(), Params(denovo=True), 1 << 16),
        "single_trio_ba": (S.trios(1), Params(), 1 << 21),
    }
    # "name" or "name@variant,threads" (pm_force_wide_plan: a given instantiation of the block-per-site kernel)
    todo = [(k, None) for k in shapes] if which == ["all"] else [(w.split("@")[0], tuple(int(x) for x in w.split("@")[1].split(",")) if "@" in w else None) for w in which]
    for k, plan in todo:
        ped, par, n = shapes[k]
        try:
            run(k if not plan else f"{k}@{plan[0]},{plan[1]}", ped, par, n, plan=plan)
        except Exception as e:
            print(json.dumps(dict(name=k, plan=plan, error=repr(e)[:300])), flush=True)
