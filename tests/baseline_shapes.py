"""Synthetic inputs of the shapes BASELINE.json names (configs 2, 4 and 5), small enough for the unmodified reference.

The same seeded generator that fills bench.py's HBM batches (polymutt_b200/synth.py, CPU tensors here) writes
  * 3,000 / 350 GLF files + ped/dat/gif for the GLF-input shapes, and
  * a PL-typed VCF + ped/dat for the --in_vcf shape (REF/ALT = reference base / its transition, the three PLs of that
    allele pair, DP = depth).
tests/golden/make_golden.py runs the UNMODIFIED reference binary on them and commits sha256 + line count of its output
(`ref_<name>.sha`, third field: sha256 of the generated input, so a drift of the generator is told apart from a
likelihood bug).  The CPU suite runs the oracle executable against those files, the GPU suite the product executable.
"""
import hashlib
import os

import numpy as np

from polymutt_b200 import capi, glfio, synth

# name -> (pedigree factory, reference arguments, sites, polymorphism boost, input kind, injected de novo fraction)
SHAPES = {
    # configs[1]: 1,000 independent trios, --denovo
    "cfg2_trios1000_dn": (lambda: synth.trios(1000), ["--denovo"], 240, 6.0, "glf", 0.08),
    # configs[4]: mixed trios + quartets (50 + 50), bi-allelic and --denovo
    "cfg5_mixed100_ba": (lambda: synth.concat(synth.trios(50), synth.families([4] * 50)), [], 400, 30.0, "glf", 0.02),
    "cfg5_mixed100_dn": (lambda: synth.concat(synth.trios(50), synth.families([4] * 50)), ["--denovo"], 300, 30.0, "glf", 0.06),
    # configs[3]: --in_vcf, 200 nuclear families x 5, PL
    "cfg4_vcf200x5": (lambda: synth.families([5] * 200), [], 400, 200.0, "vcf", 0.02),
}
SEED = 20261018


def generate(name):
    mk, args, n_sites, boost, kind, inj = SHAPES[name]
    ped = mk()
    h, r = synth.generate_sites(ped, n_sites, seed=SEED + sorted(SHAPES).index(name), cfg=synth.SynthConfig(poly_boost=boost, injected_denovo=inj))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n_sites, ped.n_person)
    digest = hashlib.sha256(hdr.tobytes() + recs.tobytes()).hexdigest()
    return ped, hdr, recs, digest


def _gi(a, b):
    a, b = min(a, b), max(a, b)
    return (a - 1) * (10 - a) // 2 + (b - a)


def write_vcf_inputs(outdir, ped, hdr, recs):
    """ped/dat + a VCF with one PL triple per sample; returns (ped path, dat path, vcf path)."""
    os.makedirs(outdir, exist_ok=True)
    names, lines = [], []
    col = 0
    for f in range(ped.n_fam):
        ids = [f"F{f + 1:04d}_{j + 1}" for j in range(int(ped.fam_size[f]))]
        for j, pid in enumerate(ids):
            fa, mo = int(ped.father[col]), int(ped.mother[col])
            lines.append(f"fam{f + 1:04d}\t{pid}\t{ids[fa] if fa >= 0 else 0}\t{ids[mo] if mo >= 0 else 0}\t{int(ped.sex[col])}\t0\n")
            col += 1
        names += ids
    paths = [os.path.join(outdir, n) for n in ("v.ped", "v.dat", "v.vcf")]
    open(paths[0], "w").writelines(lines)
    open(paths[1], "w").write("T\tGLF_Index\n")
    bases = "ACGT"
    ts = {1: 3, 2: 4, 3: 1, 4: 2}
    depth = recs["depth"][:, :, 0].astype(np.int64) | (recs["depth"][:, :, 1].astype(np.int64) << 8)
    with open(paths[2], "w") as fh:
        fh.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        for s in range(len(hdr)):
            r = int(hdr["ref_base"][s])
            a = ts[r] if s % 3 else (r % 4) + 1     # mostly transitions, every third record a transversion
            g = (_gi(r, r), _gi(r, a), _gi(a, a))
            lk = recs["lk"][s]
            cols = []
            for c in range(len(names)):
                pl = [int(lk[c, g[0]]), int(lk[c, g[1]]), int(lk[c, g[2]])]
                m = min(pl)
                cols.append(f"0/0:{int(depth[s, c])}:{pl[0] - m},{pl[1] - m},{pl[2] - m}")
            fh.write(f"1\t{int(hdr['pos'][s]) + 1}\t.\t{bases[r - 1]}\t{bases[a - 1]}\t50\tPASS\tNS={len(names)}\tGT:DP:PL\t" + "\t".join(cols) + "\n")
    return paths


def materialise(name, outdir):
    """Writes the inputs of a shape; returns (argv tail for a polymutt-compatible executable, input sha256)."""
    ped, hdr, recs, digest = generate(name)
    args, kind = SHAPES[name][1], SHAPES[name][4]
    if kind == "glf":
        pedf, datf, giff = glfio.write_run_dir(os.path.join(outdir, name), ped, hdr, recs)
        return ["-p", pedf, "-d", datf, "-g", giff] + args, digest
    pedf, datf, vcff = write_vcf_inputs(os.path.join(outdir, name), ped, hdr, recs)
    return ["-p", pedf, "-d", datf, "--in_vcf", vcff] + args, digest
