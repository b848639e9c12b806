"""Reader for .pmpk fixtures written by `pm-tools pack` (polymutt_b200/csrc/tools/pm_tools.cpp)."""
from __future__ import annotations

import gzip
import struct
from dataclasses import dataclass

import numpy as np

from .capi import PEEL_STEP_DTYPE, PERSON_SITE_DTYPE, SITE_HDR_DTYPE, PedigreeArrays


@dataclass
class Pmpk:
    ped: PedigreeArrays
    glf_index: np.ndarray      # [n_person] GLF_Index per VCF column
    people: list               # [(famid, pid, fatid, motid, sex, glf_index)] per column
    label: str
    max_position: int
    hdr: np.ndarray            # [n_sites] SITE_HDR_DTYPE
    recs: np.ndarray           # [n_sites, n_person] PERSON_SITE_DTYPE


def load_pmpk(path: str, max_sites: int | None = None) -> Pmpk:
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        buf = f.read()
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, buf, off)
        off += struct.calcsize(fmt)
        return v

    def arr(dtype, n):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=n, offset=off).copy()
        off += a.nbytes
        return a

    magic, = take("<4s")
    if magic != b"PMPK":
        raise ValueError(f"{path}: not a .pmpk file")
    _version, n_fam, n_person, n_steps = take("<Iiii")
    fam_size, fam_founders, fam_gen = arr("<i4", n_fam), arr("<i4", n_fam), arr("<i4", n_fam)
    sex = arr("u1", (n_person + 3) // 4 * 4)[:n_person]
    father, mother, glf_index = arr("<i4", n_person), arr("<i4", n_person), arr("<i4", n_person)
    peel_first = arr("<i4", n_fam + 1)
    peel = arr(PEEL_STEP_DTYPE, n_steps)
    tl, = take("<I")
    text = buf[off:off + tl].decode()
    off += tl
    max_position, n_sites = take("<iQ")
    if max_sites is not None:
        n_keep = min(n_sites, max_sites)
    else:
        n_keep = n_sites
    hdr = np.frombuffer(buf, dtype=SITE_HDR_DTYPE, count=n_keep, offset=off).copy()
    off += n_sites * SITE_HDR_DTYPE.itemsize
    recs = np.frombuffer(buf, dtype=PERSON_SITE_DTYPE, count=n_keep * n_person, offset=off).copy().reshape(n_keep, n_person)
    people, label = [], ""
    for line in text.splitlines():
        if line.startswith("#label "):
            label = line[7:]
        elif line.strip():
            t = line.split()
            people.append((t[0], t[1], t[2], t[3], int(t[4]), int(t[5])))
    ped = PedigreeArrays(fam_size, fam_founders, fam_gen, sex, father, mother, peel_first, peel)
    return Pmpk(ped, glf_index, people, label, max_position, hdr, recs)
