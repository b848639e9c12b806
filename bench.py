#!/usr/bin/env python3
"""bench.py — sites/sec of the per-site family-likelihood hot path on B200, with roofline and CPU baseline.

Default workload (BASELINE.json configs[1]): 1,000 independent trios (3,000 people), --denovo, synthetic GLF sites
(polymutt_b200/synth.py).  `--workload` selects the other BASELINE configs (same JSON contract, their own
algorithmic flop / byte formulas from SURVEY.md 8d):

  trios1000_dn   configs[1]  1,000 trios, --denovo                       block-per-site kernel  (default)
  ceph20_ba      configs[2]  20-member 3-generation pedigree             thread-per-site kernel, bi-allelic peel
  ceph20_dn      configs[2]  the same under --denovo                     thread-per-site kernel, ten-state peel
  vcf200x5       configs[3]  --in_vcf, 200 nuclear families x 5, PL      block-per-site kernel through pm_call_vcf_records
  mixed100       configs[4]  50 trios + 50 quartets                      block-per-site kernel (one warp per site)

A step = one pass of the hot path (k_sites_* -> k_compact -> k_post) over one batch of `--sites-per-step` packed sites
that already sits in HBM; consecutive steps cycle through `--resident-batches` distinct batches, each far larger than
the 126 MB L2, so no L2 flush is needed.  Sites are independent, so N GPUs each take their own contiguous site range
(weak scaling: per-GPU batch fixed) with no collective on the data path.

  python bench.py [--workload W] [--gpus N] [--steps K] [--warmup W]     our arm (one JSON line on rank 0)
  python bench.py --impl reference [--workload W] --steps K --warmup W   the reference's own CPU path
                                                                          (oracle/_ref/polymutt) on a bounded sample

`value` is device-resident throughput timed with CUDA events on the library's own stream (pm_timer_start/stop), max
over ranks.  `e2e` is the same metric through the host-buffer C-ABI call (pm_call_glf_sites / pm_call_vcf_records) from
pinned host memory, H2D and D2H copies inside the timed region.  The `cpu_baseline` leg also runs the drop-in
executable on the very shards the reference just processed and compares the two VCFs (`parity_checked_sites`).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import resource
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
_JSON_OUT = None   # the original stdout when library chatter has been redirected (multi-rank runs)
sys.path.insert(0, ROOT)

METRIC = "sites/sec for family variant+de novo calling"
UNIT = "sites/s"
SEED = 20261018


# ------------------------------------------------------------------------------------------------
# workloads: pedigree, parameters, step size, algorithmic work per site
# ------------------------------------------------------------------------------------------------
def _workloads():
    from polymutt_b200 import Params, synth
    S = synth
    return {
        "trios1000_dn": dict(title="configs[1]: 1,000 independent trios (3,000 people), --denovo, synthetic GLF sites",
                             ped=lambda: S.trios(1000), params=Params(denovo=True), ref_args=["--denovo"], sites=1 << 17, vcf=False,
                             kernel="k_sites_wide", nuclear=(1000, 1.0), peel=None),
        "ceph20_ba": dict(title="configs[2]: 3-generation 20-member CEPH-like pedigree (4 grandparents, 2 parents, 14 kids), synthetic GLF sites",
                          ped=lambda: S.ceph(), params=Params(), ref_args=[], sites=1 << 21, vcf=False,
                          kernel="k_sites_narrow", nuclear=None, peel=(3, 14, 2, 1, 20)),
        "ceph20_dn": dict(title="configs[2]: 3-generation 20-member CEPH-like pedigree, --denovo (ten-state peel), synthetic GLF sites",
                          ped=lambda: S.ceph(), params=Params(denovo=True), ref_args=["--denovo"], sites=1 << 19, vcf=False,
                          kernel="k_sites_narrow", nuclear=None, peel=(10, 14, 2, 1, 20)),
        "vcf200x5": dict(title="configs[3]: --in_vcf, 200 nuclear families x 5 members (1,000 samples), PL records",
                         ped=lambda: S.families([5] * 200), params=Params(vcf_input=True), ref_args=[], sites=1 << 18, vcf=True,
                         kernel="k_sites_wide", nuclear=(200, 3.0), peel=None),
        "mixed100": dict(title="configs[4]: 50 trios + 50 quartets (350 people), synthetic GLF sites",
                         ped=lambda: S.concat(S.trios(50), S.families([4] * 50)), params=Params(), ref_args=[], sites=1 << 19, vcf=False,
                         kernel="k_sites_wide", nuclear=(100, 1.5), peel=None),
    }


def algorithmic_flops(w, denovo, hypotheses, evaluations):
    """FP64 flops per site of the REFERENCE's formulation (FMA = 2), SURVEY.md 8d.
    Nuclear families: coefficient set-up S = 87k+18 (--denovo) or 36k+9 per (family, hypothesis), k kids; one objective
    evaluation = 18 flops per family + 20 for the shared priors (one log10 per family counted apart, costed at 0).
    Elston-Stewart: per peel  leaf 2A^3, roof ~4A^3, spouse 2A^2, init 2A*famSize  with A states."""
    if w["nuclear"]:
        n_fam, kids = w["nuclear"]
        setup = (87 * kids + 18) if denovo else (36 * kids + 9)
        return hypotheses * n_fam * setup + evaluations * (18 * n_fam + 20)
    A, leaves, roofs, spouses, fam_size = w["peel"]
    per_peel = leaves * 2 * A ** 3 + roofs * 4 * A ** 3 + spouses * 2 * A ** 2 + 2 * A * fam_size
    return evaluations * per_peel


def algorithmic_bytes_per_site(n_person, vcf):
    # GLF: 10 likelihood bytes + 3 depth + 1 mapQ per person, 4 pos + 1 ref per site; VCF: 3 PL bytes per person + 8
    return 3 * n_person + 8 if vcf else 14 * n_person + 5


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            t = [x.strip() for x in l.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0])); mx.append(float(t[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference (oracle/_ref/polymutt) on a bounded sample
# ------------------------------------------------------------------------------------------------
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "polymutt")
    if os.path.exists(p):
        return p, "reference"
    p = os.path.join(ROOT, "oracle", "_build", "polymutt_oracle_cli")  # the restatement, if the reference did not travel
    if not os.path.exists(p):
        subprocess.run(["make", "-s", "oracle"], cwd=ROOT, check=True)
    return p, "port"


def prepare_reference_shards(w, n_sites, n_shards, tmp):
    """Writes the sample as n_shards disjoint site shards (each: one GLF file per person + ped/dat/gif, or ped/dat/VCF).
    Returns the argument lists (without --out_vcf) for a polymutt-compatible executable."""
    from polymutt_b200 import capi, glfio, synth
    ped = w["ped"]()
    h, r = synth.generate_sites(ped, n_sites, seed=SEED + 7919, device="cpu", cfg=synth.SynthConfig(poly_boost=50.0) if w["vcf"] else None)
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n_sites, ped.n_person)
    shards = []
    for k in range(n_shards):
        lo, hi = k * n_sites // n_shards, (k + 1) * n_sites // n_shards
        d = os.path.join(tmp, f"shard{k}")
        if w["vcf"]:
            p = glfio.write_vcf_run_dir(d, ped, hdr[lo:hi], recs[lo:hi])
            shards.append(["-p", p[0], "-d", p[1], "--in_vcf", p[2]] + w["ref_args"])
        else:
            p = glfio.write_run_dir(d, ped, hdr[lo:hi], recs[lo:hi])
            shards.append(["-p", p[0], "-d", p[1], "-g", p[2]] + w["ref_args"])
    return shards, ped.n_person


def run_reference_once(exe, shards, tmp, n_person, tag="o"):
    """The reference has no site-level parallelism (its OpenMP sections scale 1.6x on 8 threads), so
    "all the host cores" = one single-threaded process per disjoint site shard, run concurrently; wall time of all."""
    soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
    want = n_person + 256
    if soft < want:
        resource.setrlimit(resource.RLIMIT_NOFILE, (min(max(want, soft), hard), hard))
    t0 = time.perf_counter()
    procs = [subprocess.Popen([exe] + a + ["--nthreads", "1", "--out_vcf", os.path.join(tmp, f"{tag}{k}.vcf")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for k, a in enumerate(shards)]
    rcs = [p.wait() for p in procs]
    dt = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError(f"reference exited with {rcs}")
    return dt


def host_cores():
    try:
        return max(1, min(len(os.sched_getaffinity(0)), 32))
    except Exception:
        return max(1, min(os.cpu_count() or 1, 32))


def _vcf_body(path):
    with open(path, "rb") as f:
        return [l for l in f if not l.startswith(b"##")]


def parity_check(shards, tmp, n_shards, sites_per_shard):
    """Runs the drop-in executable (CUDA engine) on the first shards the reference has just processed and compares the
    VCF bodies line by line.  Returns (sites checked, rows compared, differing rows)."""
    cli = os.path.join(ROOT, "polymutt_b200", "bin", "polymutt-b200")
    if not os.path.exists(cli):
        return 0, 0, None
    rows = bad = 0
    for k in range(n_shards):
        out = os.path.join(tmp, f"gpu{k}.vcf")
        p = subprocess.run([cli] + shards[k] + ["--out_vcf", out], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        if p.returncode != 0:
            raise RuntimeError("polymutt-b200 failed on a reference shard: " + p.stderr.decode(errors="replace")[-300:])
        a, b = _vcf_body(out), _vcf_body(os.path.join(tmp, f"o{k}.vcf"))
        rows += max(len(a), len(b))
        bad += sum(1 for x, y in zip(a, b) if x != y) + abs(len(a) - len(b))
    return n_shards * sites_per_shard, rows, bad


def cli_e2e(w, n_sites):
    """The whole drop-in executable on one input of n_sites sites of the workload: input files in -> VCF out, wall clock
    (process start, CUDA context, opening one GLF per person, ingest, engine, VCF text).  The executable's own phase
    timing (PM_TIMING) tells set-up from the per-site loop."""
    cli = os.path.join(ROOT, "polymutt_b200", "bin", "polymutt-b200")
    if not os.path.exists(cli) or n_sites <= 0:
        return None
    tmp = tempfile.mkdtemp(prefix="pm_cli_e2e_")
    try:
        shards, n_person = prepare_reference_shards(w, n_sites, 1, tmp)
        soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
        if soft < n_person + 256:
            resource.setrlimit(resource.RLIMIT_NOFILE, (min(max(n_person + 256, soft), hard), hard))
        best = None
        for _ in range(2):   # the second run has the input files in the page cache and the CUDA driver warm
            t0 = time.perf_counter()
            p = subprocess.run([cli] + shards[0] + ["--out_vcf", os.path.join(tmp, "cli.vcf")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                               env=dict(os.environ, PM_TIMING="1"))
            dt = time.perf_counter() - t0
            if p.returncode != 0:
                raise RuntimeError("polymutt-b200 failed: " + p.stdout.decode(errors="replace")[-300:])
            timing = [l for l in p.stdout.decode(errors="replace").splitlines() if l.startswith("[pm timing]") and ("loop" in l or "vcf mode" in l)]
            if best is None or dt < best[0]:
                best = (dt, timing[0] if timing else None)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    loop_s = None   # the executable's own clock around its per-site loop (everything after set-up), from its phase line
    if best[1]:
        m = re.search(r"loop ([0-9.]+) s", best[1])
        if m:
            loop_s = float(m.group(1))
        elif "vcf mode" in best[1]:
            loop_s = sum(float(x) for x in re.findall(r"(?:read|parse|engine|format|write) ([0-9.]+) s", best[1]))
    return {"value": n_sites / best[0], "unit": UNIT, "sites": n_sites, "wall_s": best[0], "phases": best[1],
            "loop_s": loop_s, "loop_value": (n_sites / loop_s) if loop_s else None,
            "note": "polymutt-b200 executable, input files -> VCF, wall clock of the whole process (best of 2 runs) on a SHORT input: "
                    "process start and CUDA context creation (1-2 s) dominate it; `loop_value` = the same sites / the executable's own clock around its per-site loop (`phases`)"}


def cpu_baseline(w, sites_per_core):
    """~10-30 s of the reference on this box's host cores; returns the cpu_baseline object."""
    exe, kind = reference_binary()
    cores = host_cores()
    n_sites = sites_per_core * cores
    tmp = tempfile.mkdtemp(prefix="pm_cpu_base_")
    try:
        shards, n_person = prepare_reference_shards(w, n_sites, cores, tmp)
        dt = run_reference_once(exe, shards, tmp, n_person)
        checked, rows, bad = parity_check(shards, tmp, min(2, cores), sites_per_core)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    out = {"value": n_sites / dt, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"{n_sites} synthetic sites of the same workload as {cores} disjoint shards of {sites_per_core} sites (one input file per "
                     f"person each), one single-threaded reference process per shard run concurrently; wall time {dt:.1f} s incl. opening "
                     f"the files and VCF writing",
           "parity_checked_sites": checked, "parity_rows_compared": rows, "parity_rows_differing": bad,
           "parity_note": "the drop-in executable (CUDA engine) re-ran the first shards of this sample; its VCF is compared line by line "
                          "(non-## lines) with the reference's"}
    if bad:
        raise RuntimeError(f"parity check failed: {bad} of {rows} VCF rows differ from the reference's on the bench workload")
    return out


def reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    exe, kind = reference_binary()
    cores = host_cores()
    n_sites = args.ref_sites_per_core * cores
    tmp = tempfile.mkdtemp(prefix="pm_ref_arm_")
    try:
        shards, n_person = prepare_reference_shards(w, n_sites, cores, tmp)
        for _ in range(args.warmup):
            run_reference_once(exe, shards, tmp, n_person)
        t = [run_reference_once(exe, shards, tmp, n_person) for _ in range(args.steps)]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    total = sum(t)
    value = n_sites * args.steps / total
    sample = (f"{n_sites} synthetic sites per step as {cores} disjoint shards of {args.ref_sites_per_core} sites, one single-threaded reference "
              f"process per shard run concurrently (the reference cannot shard sites itself)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["title"], "sites_per_step": n_sites, "persons": n_person,
                   "note": "each step = the unmodified reference binary end to end on a bounded sample (input files in, VCF out) on all host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="trios1000_dn", choices=["trios1000_dn", "ceph20_ba", "ceph20_dn", "vcf200x5", "mixed100"])
    ap.add_argument("--sites-per-step", type=int, default=0, help="0 = the workload's default")
    ap.add_argument("--resident-batches", type=int, default=2)
    ap.add_argument("--e2e-sites", type=int, default=0, help="0 = about 400 MB of wire-format input")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-sites-per-core", type=int, default=300)
    ap.add_argument("--cpu-baseline-sites-per-core", type=int, default=300)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-wide-plan", default="", help="tuning only: 'variant,threads' for pm_force_wide_plan (an instantiation of the block-per-site kernel)")
    ap.add_argument("--cli-e2e-sites", type=int, default=8192, help="sites of the whole-executable run reported as cli_e2e (0 = skip)")
    args = ap.parse_args()
    W = _workloads()
    w = W[args.workload]
    if args.impl == "reference":
        return reference_arm(args, w)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist

    from polymutt_b200 import Engine, capi, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0's stdout carries exactly one JSON line: whatever the libraries print on file descriptor 1 (NCCL's
        # version banner comes with the first collective) is sent to stderr; the JSON line goes to the saved descriptor
        global _JSON_OUT
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    ped = w["ped"]()
    npers = ped.n_person
    params = w["params"]
    vcf = w["vcf"]
    S = args.sites_per_step or w["sites"]
    R, K, Wm = args.resident_batches, args.steps, args.warmup
    lut = np.array([pow(10, -float(i) / 10.0) for i in range(256)]) if vcf else None   # FamilyLikelihoodSeq_VCF.cpp:21-22
    eng = Engine(ped, params, device=local_rank, lut=lut)
    if args.force_wide_plan:
        eng.force_wide_plan(*[int(x) for x in args.force_wide_plan.split(",")])

    # ---- synthetic batches, generated on the device; rank r owns a contiguous range of the job's world*R*S sites
    # (polymutt_b200/shard.py, the logic the world-size-2 gloo test covers) ----
    job_lo, job_hi = shard.site_range(world * R * S, rank, world)
    assert job_hi - job_lo == R * S
    gi = lambda a, b: torch.where(a < b, (a - 1) * (10 - a) // 2 + (b - a), (b - 1) * (10 - b) // 2 + (a - b))
    batches = []
    for b in range(R):
        hdr = torch.empty((S, 8), dtype=torch.uint8, device=dev)
        recs = torch.empty((S, npers, 16), dtype=torch.uint8, device=dev)
        synth.generate_sites(ped, S, seed=SEED + 1000 * rank + b, device=dev, out_hdr=hdr, out_recs=recs, chunk=max(256, (1 << 24) // npers),
                             pos0=job_lo + b * S, cfg=synth.SynthConfig(poly_boost=50.0) if vcf else None)
        mono = None
        if vcf:
            # a VCF record of the site: (REF, ALT) = (ref, its transition), the three PLs of that pair kept, the rest cleared;
            # mono = sum over samples of -PL[ref/ref]/10 (what the host front end computes while tokenising)
            ref = hdr[:, 4].long()
            alt = ((ref - 1) ^ 2) + 1
            g = torch.stack([gi(ref, ref), gi(ref, alt), gi(alt, alt)], dim=1)                     # [S, 3]
            keep = torch.zeros((S, 16), dtype=torch.bool, device=dev)
            keep.scatter_(1, g, True)
            keep[:, 10:14] = True
            recs *= keep[:, None, :].to(torch.uint8)
            hdr[:, 6] = alt.to(torch.uint8)
            hdr[:, 7] = 0
            mono = -(recs.gather(2, g[:, None, :1].expand(S, npers, 1))[:, :, 0].to(torch.float64)).sum(dim=1) / 10.0
        batches.append((hdr, recs, mono))
    cap = S if vcf else max(4096, S // 16)
    status = torch.empty(S, dtype=torch.uint16, device=dev)
    res_out = torch.empty((cap, capi.SITE_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    # per-person results: 96-byte rows for the emitted GLF sites; VCF input: 2 bytes per sample (best | gq << 8: what its writer prints from)
    per_out = (torch.empty((cap, npers), dtype=torch.uint16, device=dev) if vcf else
               torch.empty((cap, npers, capi.PERSON_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev))
    n_res = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step(i):
        hdr, recs, mono = batches[i % R]
        if vcf:
            eng.call_vcf_records_calls_device(hdr.data_ptr(), recs.data_ptr(), mono.data_ptr(), S, False, status.data_ptr(), res_out.data_ptr(), per_out.data_ptr())
        else:
            eng.call_glf_sites_device(hdr.data_ptr(), recs.data_ptr(), S, capi.PM_OUT_EMITTED, status.data_ptr(), res_out.data_ptr(),
                                      per_out.data_ptr(), cap, n_res.data_ptr())

    for i in range(Wm):
        step(i)
    eng.sync()
    eng.reset_counters()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.timer_start()
    for i in range(K):
        step(i)
    ms = eng.timer_stop()
    eng.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    counters = eng.counters()
    emitted_last = S if vcf else int(n_res.item())
    # main-kernel time per launch (CUDA events around the site kernel inside the library), from one more step
    step(0)
    eng.sync()
    main_ms, total_ms, launches_per_step = eng.last_timing()

    # max over ranks of the device time, sum over ranks of the sites -> whole-job sites/s
    ms_max, _, value = shard.reduce_timing(ms, K * S, dist if world > 1 else None, dev)

    # ---- e2e through the host-buffer C-ABI call, pinned host memory: every rank at the same time (they share the
    # host's memory and PCIe root complexes), whole-job value = all ranks' sites / the slowest rank's time ----
    Se = min(args.e2e_sites or max(2048, (400 << 20) // (npers * (3 if vcf else 14))), S)
    h_hdr = torch.empty((Se, 8), dtype=torch.uint8, pin_memory=True)
    # the wire forms of the records: GLF 14 bytes per (site, person) (pm_person_site_wire), VCF 3 PL bytes per (record, sample)
    h_hdr.copy_(batches[0][0][:Se])
    if vcf:
        ref = batches[0][0][:Se, 4].long(); alt = batches[0][0][:Se, 6].long()
        g = torch.stack([gi(ref, ref), gi(ref, alt), gi(alt, alt)], dim=1)
        rec_bytes = 3
        h_recs = torch.empty((Se, npers, 3), dtype=torch.uint8, pin_memory=True)
        h_recs.copy_(batches[0][1][:Se].gather(2, g[:, None, :].expand(Se, npers, 3)))
    else:
        rec_bytes = 14
        h_recs = torch.empty((Se, npers, 14), dtype=torch.uint8, pin_memory=True)
        h_recs.copy_(batches[0][1][:Se, :, :14])
    cap_e = Se if vcf else max(1024, Se // 8)
    h_status = torch.empty(Se, dtype=torch.uint16, pin_memory=True)
    h_res = torch.empty((cap_e, capi.SITE_RESULT_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
    h_per = None if vcf else torch.empty((cap_e, npers, capi.PERSON_RESULT_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
    h_mono = h_calls = None
    if vcf:
        h_mono = torch.empty(Se, dtype=torch.float64, pin_memory=True)
        h_mono.copy_(batches[0][2][:Se])
        h_calls = torch.empty((Se, npers), dtype=torch.uint16, pin_memory=True)   # best | gq << 8: what the --in_vcf writer prints from
    torch.cuda.synchronize()
    nres = C.c_size_t(0)

    def e2e_step():
        if vcf:
            rc = eng.lib.pm_call_vcf_records_pl(eng.ctx, h_hdr.data_ptr(), h_recs.data_ptr(), h_mono.data_ptr(), Se, h_res.data_ptr(), h_calls.data_ptr())
            nres.value = Se
        else:
            rc = eng.lib.pm_call_glf_sites_wire(eng.ctx, h_hdr.data_ptr(), h_recs.data_ptr(), Se, capi.PM_OUT_EMITTED, h_status.data_ptr(),
                                                h_res.data_ptr(), h_per.data_ptr(), cap_e, C.byref(nres))
        if rc != 0:
            raise RuntimeError(eng.lib.pm_last_error().decode())

    e2e_step()  # warm-up (allocates the staging buffers)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    e2e_dt = (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
        te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_dt = float(te.item())
    rows = nres.value
    e2e = {"value": world * Se / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": world * Se * (npers * rec_bytes + 8 + (8 if vcf else 0)),
           "d2h_bytes_per_step": world * (Se * (capi.SITE_RESULT_DTYPE.itemsize + npers * 2) if vcf else
                                          Se * 2 + rows * (capi.SITE_RESULT_DTYPE.itemsize + npers * capi.PERSON_RESULT_DTYPE.itemsize) + 4),
           "sites_per_step": world * Se, "ms_per_step": e2e_dt * 1e3, "n_gpus": world,
           "note": ("pm_call_vcf_records_pl (3 PL bytes per sample in, 2 bytes per sample out)" if vcf else "pm_call_glf_sites_wire (14-byte records)") +
                   " from pinned host buffers; H2D of the sites and D2H of status + "
                   "result rows inside the timed region; all ranks run it at the same time, time = max over ranks"}

    if rank == 0:
        # ---- roofline of the dominant kernel ----
        fp64_peak = eng.measure_fp64_peak()
        copy_bw = eng.measure_copy_bw()
        peaks = measured_peaks()
        sites_ev = max(1, counters["sites_evaluated"])
        hyp_per_site = counters["hypotheses"] / sites_ev
        ev_per_site = counters["evaluations"] / sites_ev
        flops_per_site = algorithmic_flops(w, bool(params.denovo), hyp_per_site, ev_per_site)
        flops_per_launch = flops_per_site * S
        achieved_tflops = flops_per_launch / (main_ms * 1e-3) / 1e12
        bytes_per_launch = algorithmic_bytes_per_site(npers, vcf) * S
        hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
        # DRAM traffic of the kernel: not measurable from inside this process; taken per site from the committed
        # `ncu --set full` capture of the same kernel on the same workload and scaled to this launch's sites
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tj = json.load(f).get(args.workload)
            per_site = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["sites_in_capture"]
            traffic = per_site * S
            traffic_src = "%s: %.0f DRAM bytes/site (algorithmic %d)" % (tj["source"], per_site, algorithmic_bytes_per_site(npers, vcf))
        except (OSError, KeyError, ValueError, TypeError):
            pass
        hbm_frac = bytes_per_launch / (main_ms * 1e-3) / 1e9 / hbm_peak
        fp_frac = achieved_tflops / (fp64_peak / 1e12)
        bound = "fp64" if fp_frac >= hbm_frac else "hbm"
        roofline = {
            "bound": bound, "kernel": w["kernel"], "achieved": achieved_tflops if bound == "fp64" else bytes_per_launch / (main_ms * 1e-3) / 1e9,
            "peak": fp64_peak / 1e12 if bound == "fp64" else hbm_peak, "unit": "TFLOP/s" if bound == "fp64" else "GB/s",
            "frac": max(fp_frac, hbm_frac), "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
            "peak_source": "DFMA microbenchmark measured live on this GPU (pm_measure_fp64_peak; method and history: profiles/fp64_peak.json); "
                           "MEASURED_PEAKS.json has no FP64 figure",
            "note": "achieved = the REFERENCE formulation's flops per site (SURVEY.md 8d: H x set-up + E x evaluation, H and E counted by the kernel) / "
                    "kernel time. The kernels execute fewer: quartic units about a quarter of them, the factorised ten-state Elston-Stewart peel about a "
                    "tenth -- so the fraction of an extended-pedigree workload can pass 1; the FP64 pipe's own utilisation is in the ncu summaries under profiles/",
            "flops_per_site": flops_per_site, "hypotheses_per_site": hyp_per_site, "evaluations_per_site": ev_per_site, "kernel_ms_per_launch": main_ms,
            "fp64": {"achieved": achieved_tflops, "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": fp_frac},
            "hbm": {"achieved": bytes_per_launch / (main_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_frac,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "copy_kernel_gbs_live": copy_bw / 1e9},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["title"], "workload_key": args.workload, "sites_per_step_per_gpu": S, "persons": npers,
                       "resident_batches": R, "input_bytes_per_step": S * (npers * 16 + 8),
                       "l2": ("each step reads a different resident batch of %.2f GB, larger than the 126 MB L2 (no flush needed)" if S * (npers * 16 + 8) > 126e6
                              else "WARNING: a resident batch is only %.2f GB, not larger than the 126 MB L2 -- raise --sites-per-step") % (S * (npers * 16 + 8) / 1e9),
                       "kernel_plan": eng.describe_plan(),
                       "emitted_rows_last_step": emitted_last, "result_capacity_rows": cap},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step) * K, "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(w, args.cpu_baseline_sites_per_core)
            except Exception as ex:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)[:300]}
            try:
                line["cli_e2e"] = cli_e2e(w, args.cli_e2e_sites)
            except Exception as ex:
                line["cli_e2e"] = {"value": None, "unit": UNIT, "note": repr(ex)[:300]}
        print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
