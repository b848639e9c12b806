"""Builds (pedigree arrays, packed sites) for a test pedigree from the example fixture."""
import os
import subprocess

import numpy as np

import cli_util as U
from polymutt_b200 import load_pmpk


def pedigree_from_file(ped_file, tmpdir):
    """Pedigree arrays + GLF_Index per VCF column, via the product's own loader (pm-tools pack without -g)."""
    out = os.path.join(tmpdir, os.path.basename(ped_file) + ".pmpk")
    subprocess.run([U.PM_TOOLS, "pack", "-p", ped_file, "-d", os.path.join(U.GOLDEN, "peds", "test.dat"), "-o", out],
                   check=True, stderr=subprocess.DEVNULL)
    p = load_pmpk(out)
    return p.ped, p.glf_index


def sites_for(example12, glf_index, n_sites=None, start=0):
    """Gathers the fixture's streams (column i = GLF i+1) into the column order of a pedigree."""
    stop = len(example12.hdr) if n_sites is None else min(len(example12.hdr), start + n_sites)
    hdr = example12.hdr[start:stop].copy()
    src = example12.recs[start:stop]
    recs = np.zeros((stop - start, len(glf_index)), dtype=src.dtype)
    for c, gi in enumerate(glf_index):
        if gi > 0:
            recs[:, c] = src[:, gi - 1]
    return hdr, recs
