"""Whole-executable throughput on the GPU box: GLF files in -> VCF out, and VCF in -> VCF out.

Writes synthetic inputs under /tmp, runs polymutt-b200 on them with different --ingest_threads and prints one JSON
line per run (sites or records per second of wall time, rows written).  Not a bench.py metric: it shows where the
executable's time goes once the likelihood engine is on the GPU.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from polymutt_b200 import capi, glfio, synth  # noqa: E402

CLI = os.path.join(ROOT, "polymutt_b200", "bin", "polymutt-b200")


def run(cmd, label, units, unit_name):
    t = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=dict(os.environ, PM_TIMING="1"))
    dt = time.time() - t
    timing = [ln for ln in p.stdout.decode(errors="replace").splitlines() if ln.startswith("[pm timing]")]
    out = cmd[cmd.index("--out_vcf") + 1]
    rows = sum(1 for ln in open(out, "rb") if not ln.startswith(b"#")) if os.path.exists(out) else -1
    print(json.dumps({"run": label, "exit": p.returncode, "seconds": round(dt, 3), unit_name + "_per_s": round(units / dt),
                      "rows": rows, "out_MB": round(os.path.getsize(out) / 1e6, 1) if rows >= 0 else None,
                      "timing": timing[0] if timing else None}), flush=True)
    if p.returncode != 0:
        print(p.stdout.decode(errors="replace")[-1500:], flush=True)


def glf_mode(n_sites, outdir="/tmp/e2e_glf", all_sites_only=False):
    ped = synth.trios(1000)
    h, r = synth.generate_sites(ped, n_sites, seed=20261018)
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n_sites, ped.n_person)
    t = time.time()
    pedf, datf, giff = glfio.write_run_dir(outdir, ped, hdr, recs)
    print(json.dumps({"wrote": "3000 GLF files", "sites": n_sites, "seconds": round(time.time() - t, 1)}), flush=True)
    base = [CLI, "-p", pedf, "-d", datf, "-g", giff]
    if all_sites_only:
        run(base + ["--all_sites", "--out_vcf", outdir + "/all.vcf"], "glf_all_sites", n_sites, "sites")
        os.remove(outdir + "/all.vcf")
        return
    for thr in (1, 0):
        run(base + ["--denovo", "--out_vcf", "/tmp/e2e_glf/dn.vcf", "--ingest_threads", str(thr)], f"glf_denovo_threads{thr}", n_sites, "sites")
    run(base + ["--out_vcf", "/tmp/e2e_glf/ba.vcf"], "glf_variants", n_sites, "sites")


def vcf_mode(n_rec):
    os.makedirs("/tmp/e2e_vcf", exist_ok=True)
    rng = np.random.default_rng(1)
    nf = 200
    names, ped = [], []
    for f in range(nf):
        ids = [f"F{f}_{j}" for j in range(5)]
        ped.append(f"fam{f}\t{ids[0]}\t0\t0\t1\t0\n")
        ped.append(f"fam{f}\t{ids[1]}\t0\t0\t2\t0\n")
        for j in range(2, 5):
            ped.append(f"fam{f}\t{ids[j]}\t{ids[0]}\t{ids[1]}\t{1 + j % 2}\t0\n")
        names += ids
    open("/tmp/e2e_vcf/v.ped", "w").writelines(ped)
    open("/tmp/e2e_vcf/v.dat", "w").write("T\tGLF_Index\n")
    bases = "ACGT"
    n = len(names)
    with open("/tmp/e2e_vcf/v.vcf", "w") as fh:
        fh.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        for s in range(n_rec):
            r = int(rng.integers(0, 4))
            a = (r + int(rng.integers(1, 4))) % 4
            g = rng.binomial(2, rng.random() * 0.5, size=n)
            dp = rng.poisson(15, size=n)
            cols = [f"{('0/0', '0/1', '1/1')[gg]}:{d}:{min(255, abs(gg) * d * 3)},{min(255, abs(gg - 1) * d * 3)},{min(255, abs(gg - 2) * d * 3)}" for gg, d in zip(g.tolist(), dp.tolist())]
            fh.write(f"1\t{1000 + s}\t.\t{bases[r]}\t{bases[a]}\t50\tPASS\tNS=1000\tGT:DP:PL\t" + "\t".join(cols) + "\n")
    print(json.dumps({"wrote": "VCF 200 families x 5", "records": n_rec, "MB": round(os.path.getsize("/tmp/e2e_vcf/v.vcf") / 1e6)}), flush=True)
    base = [CLI, "-p", "/tmp/e2e_vcf/v.ped", "-d", "/tmp/e2e_vcf/v.dat", "--in_vcf", "/tmp/e2e_vcf/v.vcf"]
    for thr in (1, 0):
        run(base + ["--out_vcf", "/tmp/e2e_vcf/out.vcf", "--ingest_threads", str(thr)], f"vcf_in_threads{thr}", n_rec, "records")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "vcf":   # only the --in_vcf runs: gpu_cli_e2e.py vcf <records>
        vcf_mode(int(sys.argv[2]) if len(sys.argv) > 2 else 10000)
        sys.exit(0)
    glf_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 12000)
    glf_mode(int(sys.argv[3]) if len(sys.argv) > 3 else 8000, outdir="/tmp/e2e_glf_all", all_sites_only=True)
    vcf_mode(int(sys.argv[2]) if len(sys.argv) > 2 else 10000)
