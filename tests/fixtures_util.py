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


# three "chromosomes" cut out of the example's single section (site index ranges), the last one named X
MULTI_SECTIONS = [("1", 0, 30000), ("2", 30000, 55000), ("X", 55000, None)]


def write_multisection_glfs(example12, outdir):
    """One GLF per example stream (GLF_Index 1..12) holding MULTI_SECTIONS back to back, plus the gif file.
    Returns the gif path.  Positions are kept, every section's maxPosition is its last position + 1."""
    from polymutt_b200 import glfio
    os.makedirs(outdir, exist_ok=True)
    hdr, recs = example12.hdr, example12.recs
    gif = []
    for g in range(recs.shape[1]):
        blob = b""
        for k, (label, a, b) in enumerate(MULTI_SECTIONS):
            b = len(hdr) if b is None else b
            tmp = os.path.join(outdir, "sec.tmp")
            pos = hdr["pos"][a:b].astype(np.int64)
            glfio.write_glf(tmp, label, int(pos.max()) + 1, pos, hdr["ref_base"][a:b], recs[a:b, g])
            raw = open(tmp, "rb").read()
            blob += raw if k == 0 else raw[8:]   # later sections go without the 8-byte file header
        path = os.path.join(outdir, f"ms{g + 1}.glf")
        open(path, "wb").write(blob)
        gif.append(f"{g + 1} {path}\n")
    os.remove(os.path.join(outdir, "sec.tmp"))
    gif_path = os.path.join(outdir, "gif")   # names that cli_util.run_cli expects in a GLF directory
    open(gif_path, "w").write("".join(gif))
    open(os.path.join(outdir, "dat"), "w").write(open(os.path.join(U.GOLDEN, "peds", "test.dat")).read())
    return gif_path


def repeat_sites(n_sites, n_streams):
    """Which sites get a repeated record in which example stream (write_repeat_glfs): 30 seeded sites in streams 1, 4, 5, 8
    and 11; streams 4 and 5 also share ten sites (two streams repeat the same position), stream 8 repeats five of its
    sites twice (three records at one position), and stream 11 repeats its last site."""
    rng = np.random.default_rng(20261019)
    rep = {g: sorted(int(x) for x in rng.integers(0, n_sites - 1, 30)) for g in (0, 3, 4, 7, 10)}
    shared = [int(x) for x in rng.integers(0, n_sites - 1, 10)]
    rep[3] = sorted(rep[3] + shared)
    rep[4] = sorted(rep[4] + shared)
    rep[7] = sorted(rep[7] + rep[7][:5])
    rep[10] = sorted(rep[10] + [n_sites - 1])
    return rep


def write_repeat_glfs(example12, outdir):
    """The example's 12 GLFs (one section, GLF_Index 1..12) with base records of offset 0 -- repeated positions -- put
    into five of them (repeat_sites), plus the gif and dat files.  Returns the gif path."""
    from polymutt_b200 import glfio
    os.makedirs(outdir, exist_ok=True)
    hdr, recs = example12.hdr, example12.recs
    rep = repeat_sites(len(hdr), recs.shape[1])
    pos = hdr["pos"].astype(np.int64)
    gif = []
    for g in range(recs.shape[1]):
        path = os.path.join(outdir, f"rep{g + 1}.glf")
        glfio.write_glf(path, example12.label, int(example12.max_position), pos, hdr["ref_base"], recs[:, g], repeats=rep.get(g, ()))
        gif.append(f"{g + 1} {path}\n")
    gif_path = os.path.join(outdir, "gif")
    open(gif_path, "w").write("".join(gif))
    open(os.path.join(outdir, "dat"), "w").write(open(os.path.join(U.GOLDEN, "peds", "test.dat")).read())
    return gif_path
