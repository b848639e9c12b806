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


def materialise(name, outdir):
    """Writes the inputs of a shape; returns (argv tail for a polymutt-compatible executable, input sha256)."""
    ped, hdr, recs, digest = generate(name)
    args, kind = SHAPES[name][1], SHAPES[name][4]
    if kind == "glf":
        pedf, datf, giff = glfio.write_run_dir(os.path.join(outdir, name), ped, hdr, recs)
        return ["-p", pedf, "-d", datf, "-g", giff] + args, digest
    pedf, datf, vcff = glfio.write_vcf_run_dir(os.path.join(outdir, name), ped, hdr, recs)
    return ["-p", pedf, "-d", datf, "--in_vcf", vcff] + args, digest
