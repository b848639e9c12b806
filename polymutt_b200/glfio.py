"""Minimal GLF v3 writer (core/glfHandler.h:21-42 record layout) used to hand synthetic sites to the
reference binary / the drop-in executable in bench.py and the tests."""
from __future__ import annotations

import os
import struct

import numpy as np

_REC = np.dtype([("tag", "u1"), ("offset", "<u4"), ("depth_minllk", "<u4"), ("mapq", "u1"), ("lk", "u1", (10,))])
assert _REC.itemsize == 20
_BACK = np.array([15, 1, 2, 4, 8], dtype=np.uint8)  # glfHandler::backTranslateBase


def write_glf(path: str, label: str, max_position: int, pos: np.ndarray, ref_base: np.ndarray, recs: np.ndarray,
              indel_every: int = 0, end_marker: bool = True, repeats=()):
    """recs: [n_sites] PERSON_SITE_DTYPE for one person; rows that are all zero are left out.
    indel_every > 0 interleaves an indel record (type 2, offset 0) after every indel_every-th base record;
    end_marker=False leaves the section without its terminating byte (a truncated file);
    repeats: site indices whose record is followed by one more base record with offset 0 -- a repeated position --
    carrying the data of the person's next record (an index listed k times gets k repeats; sites without a record are
    skipped)."""
    depth = recs["depth"][:, 0].astype(np.uint32) | (recs["depth"][:, 1].astype(np.uint32) << 8) | (recs["depth"][:, 2].astype(np.uint32) << 16)
    keep = (depth > 0) | (recs["map_quality"] > 0) | (recs["lk"].max(axis=1) > 0)
    p = pos[keep].astype(np.int64)
    out = np.zeros(len(p), dtype=_REC)
    out["tag"] = (1 << 4) | _BACK[ref_base[keep]]
    out["offset"] = np.diff(np.concatenate([[0], p])).astype(np.uint32)
    out["depth_minllk"] = depth[keep]
    out["mapq"] = recs["map_quality"][keep]
    out["lk"] = recs["lk"][keep]
    if len(repeats) and len(out):
        site_to_rec = np.cumsum(keep) - 1
        at, extra = [], []
        for sidx in repeats:
            if not keep[sidx]:
                continue
            k = int(site_to_rec[sidx])
            rec = out[min(k + 1, len(out) - 1)].copy()
            rec["tag"] = out[k]["tag"]
            rec["offset"] = 0
            at.append(k + 1)
            extra.append(rec)
        if at:
            out = np.insert(out, at, np.array(extra, dtype=_REC))
        p = np.cumsum(out["offset"].astype(np.int64))
    lab = label.encode() + b"\0"
    with open(path, "wb") as f:
        f.write(b"GLF\x03" + struct.pack("<I", 0))
        f.write(struct.pack("<i", len(lab)) + lab + struct.pack("<i", max_position))
        if indel_every > 0:
            indel = bytes([(2 << 4) | 15]) + struct.pack("<I", 0) + struct.pack("<I", 7) + bytes([60]) + bytes([10, 20, 30]) + struct.pack("<hh", 2, -3) + b"AC" + b"GTT"
            raw = out.tobytes()
            for k in range(len(p)):
                f.write(raw[20 * k:20 * (k + 1)])
                if (k + 1) % indel_every == 0:
                    f.write(indel)
        else:
            f.write(out.tobytes())
        if end_marker:
            f.write(b"\0")


def write_run_dir(outdir: str, ped, hdr: np.ndarray, recs: np.ndarray, label: str = "1"):
    """Writes ped/dat/gif + one GLF per person for a packed batch; returns (ped, dat, gif) paths."""
    os.makedirs(outdir, exist_ok=True)
    firsts = ped.family_first()
    lines, gif = [], []
    pos = hdr["pos"].astype(np.int64)
    max_position = int(pos.max()) + 1 if len(pos) else 1
    col = 0
    for f in range(ped.n_fam):
        for j in range(int(ped.fam_size[f])):
            fa, mo = int(ped.father[col]), int(ped.mother[col])
            pid = lambda k: f"p{k + 1:02d}"
            # pids sort naturally in column order; founders come first inside each family by construction
            lines.append(f"fam{f + 1:05d}\t{pid(j)}\t{pid(fa) if fa >= 0 else 0}\t{pid(mo) if mo >= 0 else 0}\t{int(ped.sex[col])}\t{col + 1}\n")
            path = os.path.join(outdir, f"g{col + 1}.glf")
            write_glf(path, label, max_position, pos, hdr["ref_base"], recs[:, col])
            gif.append(f"{col + 1} {path}\n")
            col += 1
    paths = [os.path.join(outdir, n) for n in ("run.ped", "run.dat", "run.gif")]
    open(paths[0], "w").write("".join(lines))
    open(paths[1], "w").write("T\tGLF_Index\n")
    open(paths[2], "w").write("".join(gif))
    return paths


def _gi(a, b):
    a, b = min(a, b), max(a, b)
    return (a - 1) * (10 - a) // 2 + (b - a)


def write_vcf_run_dir(outdir, ped, hdr, recs):
    """--in_vcf inputs for a packed batch: ped/dat + a VCF whose records carry REF / ALT (the transition of REF, every
    third record a transversion), DP and the three PLs of that allele pair per sample; returns (ped, dat, vcf) paths."""
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
