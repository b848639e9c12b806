"""Seeded synthetic pedigrees and packed GLF sites (the shapes BASELINE.json names).

Data model: reference base uniform over A,C,G,T; a site is polymorphic with the
Watterson prior theta * sum_{i<=2F} 1/i; the alternative allele is the transition with probability
2/3; the founder allele frequency is drawn proportional to 1/x on [1/2F, 1-1/2F]; genotypes are
gene-dropped through the pedigree with a per-transmission de novo rate (plus an injected stratum so
that de novo rows exist); depth ~ Poisson(mean_depth) with a fraction of empty person-sites; reads are
binomial with a base error rate; the ten phred likelihoods come from the read counts, min-normalised
and capped at 255; mapping quality 100.

Everything is torch so that the same code fills a small CPU tensor for the parity tests and a
multi-GB HBM buffer for bench.py (no host round trip).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from .capi import PedigreeArrays

_GENO = [(0, 0), (0, 1), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 3)]  # AA AC AG AT CC CG CT GG GT TT


def trios(n: int) -> PedigreeArrays:
    """n independent father/mother/child trios."""
    return families([3] * n)


def families(sizes) -> PedigreeArrays:
    """Nuclear families (size >= 3: two founders + kids) and unrelated singletons (size 1)."""
    fam_size, fam_founders, fam_gen, sex, father, mother = [], [], [], [], [], []
    for s in sizes:
        fam_size.append(s)
        if s == 1:
            fam_founders.append(1); fam_gen.append(1)
            sex.append(1 + len(fam_size) % 2); father.append(-1); mother.append(-1)
        else:
            assert s >= 3
            fam_founders.append(2); fam_gen.append(2)
            sex += [1, 2] + [1 + ((k + len(fam_size)) % 2) for k in range(s - 2)]
            father += [-1, -1] + [0] * (s - 2)
            mother += [-1, -1] + [1] * (s - 2)
    return PedigreeArrays(np.array(fam_size), np.array(fam_founders), np.array(fam_gen), np.array(sex, dtype=np.uint8),
                          np.array(father), np.array(mother))


def ceph(n_kids: int = 14) -> PedigreeArrays:
    """Three-generation CEPH-like pedigree: 4 grandparents, 2 parents, n_kids children (one family).
    Column order = founders first (g1 g2 g3 g4), then p1 (child of g1,g2), p2 (child of g3,g4), then kids."""
    n = 6 + n_kids
    sex = [1, 2, 1, 2, 1, 2] + [1 + (k % 2) for k in range(n_kids)]
    father = [-1, -1, -1, -1, 0, 2] + [4] * n_kids
    mother = [-1, -1, -1, -1, 1, 3] + [5] * n_kids
    return PedigreeArrays(np.array([n]), np.array([4]), np.array([3]), np.array(sex, dtype=np.uint8), np.array(father), np.array(mother))


def clan(n_children: int = 10, kids_per_couple: int = 2) -> PedigreeArrays:
    """Three generations, many marriage nodes: a founder couple with n_children children, every child married to a founder
    and every such couple with kids_per_couple kids.  One family of 2 + 2 n_children + n_children kids_per_couple members and
    1 + n_children couples.  Column order: founders first (the couple, then the n_children spouses), then the children,
    then the grandchildren."""
    nf = 2 + n_children
    sex = [1, 2] + [2 - (c % 2) for c in range(n_children)]                  # spouse of child c has the opposite sex of the child
    father, mother = [-1] * nf, [-1] * nf
    for c in range(n_children):                                              # child c: sex alternates
        sex.append(1 + (c % 2)); father.append(0); mother.append(1)
    for c in range(n_children):
        child, spouse = nf + c, 2 + c
        fa, mo = (child, spouse) if sex[child] == 1 else (spouse, child)
        for k in range(kids_per_couple):
            sex.append(1 + ((c + k) % 2)); father.append(fa); mother.append(mo)
    n = len(sex)
    return PedigreeArrays(np.array([n]), np.array([nf]), np.array([3]), np.array(sex, dtype=np.uint8), np.array(father), np.array(mother))


def concat(*peds: PedigreeArrays) -> PedigreeArrays:
    return PedigreeArrays(*(np.concatenate([getattr(p, k) for p in peds]) for k in
                            ("fam_size", "fam_founders", "fam_generations", "sex", "father", "mother")))


@dataclass
class SynthConfig:
    theta: float = 1e-3
    denovo_rate: float = 1.5e-8
    injected_denovo: float = 1e-5     # fraction of sites with one forced de novo allele in a random non-founder
    mean_depth: float = 15.0
    empty_fraction: float = 0.02
    base_error: float = 0.01
    poly_boost: float = 1.0           # multiply the polymorphism prior (tests use > 1 to see more variants)


def generate_sites(ped: PedigreeArrays, n_sites: int, seed: int, device="cpu", cfg: SynthConfig | None = None,
                   out_hdr: torch.Tensor | None = None, out_recs: torch.Tensor | None = None, chunk: int = 1 << 14,
                   pos0: int = 0):
    """Returns (hdr uint8 [n_sites, 8], recs uint8 [n_sites, n_person, 16]) on `device`."""
    cfg = cfg or SynthConfig()
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    npers = ped.n_person
    firsts = ped.family_first()
    fam_of = np.repeat(np.arange(ped.n_fam), ped.fam_size)
    fa_abs = np.where(ped.father >= 0, firsts[fam_of] + ped.father, -1)
    mo_abs = np.where(ped.mother >= 0, firsts[fam_of] + ped.mother, -1)
    founders = np.flatnonzero(fa_abs < 0)
    nonfounders = np.flatnonzero(fa_abs >= 0)
    F = len(founders)
    prior = min(0.5, cfg.poly_boost * cfg.theta * sum(1.0 / i for i in range(1, 2 * F + 1)))
    if out_hdr is None:
        out_hdr = torch.empty((n_sites, 8), dtype=torch.uint8, device=dev)
    if out_recs is None:
        out_recs = torch.empty((n_sites, npers, 16), dtype=torch.uint8, device=dev)
    # per-genotype base probabilities: q[g, b] = 0.5 (P(b|x) + P(b|y))
    e = cfg.base_error
    pb = torch.full((4, 4), e / 3, dtype=torch.float64)
    pb.fill_diagonal_(1 - e)
    q = torch.stack([0.5 * (pb[x] + pb[y]) for x, y in _GENO])          # [10, 4]
    logq = torch.log10(q).t().contiguous().to(dev, torch.float32)       # [4, 10]
    ts_of = torch.tensor([2, 3, 0, 1], device=dev)
    tv1_of = torch.tensor([1, 0, 1, 0], device=dev)
    tv2_of = torch.tensor([3, 2, 3, 2], device=dev)
    fa_t = torch.as_tensor(fa_abs, device=dev)
    mo_t = torch.as_tensor(mo_abs, device=dev)
    lo, hi = 1.0 / (2 * F), 1.0 - 1.0 / (2 * F)

    for s0 in range(0, n_sites, chunk):
        S = min(chunk, n_sites - s0)
        u = lambda *shape: torch.rand(shape, generator=g, device=dev)
        ref = torch.randint(0, 4, (S,), generator=g, device=dev)
        is_poly = u(S) < prior
        r = u(S)
        alt = torch.where(r < 2.0 / 3, ts_of[ref], torch.where(r < 5.0 / 6, tv1_of[ref], tv2_of[ref]))
        # allele frequency of the ALT allele, density ~ 1/x on [lo, hi]
        x = lo * (hi / lo) ** u(S) if hi > lo else torch.full((S,), 0.5, device=dev)
        x = torch.where(is_poly, x, torch.zeros_like(x))
        # alleles[s, person, 2] as 0 = ref / 1 = alt flags
        al = torch.zeros((S, npers, 2), dtype=torch.bool, device=dev)
        fidx = torch.as_tensor(founders, device=dev)
        al[:, fidx, :] = u(S, F, 2) < x[:, None, None]
        # gene dropping in column order (parents precede children inside a family)
        for i in nonfounders:
            pick_f = (u(S) < 0.5).long()
            pick_m = (u(S) < 0.5).long()
            a_f = al[:, int(fa_abs[i]), :].gather(1, pick_f[:, None])[:, 0]
            a_m = al[:, int(mo_abs[i]), :].gather(1, pick_m[:, None])[:, 0]
            mut = u(S, 2) < cfg.denovo_rate
            al[:, i, 0] = a_f ^ mut[:, 0]
            al[:, i, 1] = a_m ^ mut[:, 1]
        if len(nonfounders) and cfg.injected_denovo > 0:
            inj = u(S) < cfg.injected_denovo
            who = torch.as_tensor(nonfounders, device=dev)[torch.randint(0, len(nonfounders), (S,), generator=g, device=dev)]
            rows = torch.nonzero(inj)[:, 0]
            al[rows, who[rows], 0] = ~al[rows, who[rows], 0]
        # mono sites carry no alt allele (x == 0), injected mutations use `alt`
        base = torch.where(al, alt[:, None, None], ref[:, None, None])   # [S, N, 2] base codes 0..3
        depth = torch.poisson(torch.full((S, npers), cfg.mean_depth, device=dev), generator=g)
        depth = torch.where(u(S, npers) < cfg.empty_fraction, torch.zeros_like(depth), depth)
        n1 = torch.binomial(depth, torch.full_like(depth, 0.5), generator=g)   # reads from allele 0
        counts = torch.zeros((S, npers, 4), dtype=torch.float32, device=dev)
        for k, nk in ((0, n1), (1, depth - n1)):
            nerr = torch.binomial(nk, torch.full_like(nk, e), generator=g)
            counts.scatter_add_(2, base[:, :, k:k + 1], (nk - nerr)[:, :, None].float())
            # errors go to one of the three other bases
            e1 = torch.binomial(nerr, torch.full_like(nerr, 1.0 / 3), generator=g)
            e2 = torch.binomial(nerr - e1, torch.full_like(nerr, 0.5), generator=g)
            e3 = nerr - e1 - e2
            for j, ej in enumerate((e1, e2, e3)):
                counts.scatter_add_(2, (base[:, :, k:k + 1] + j + 1) % 4, ej[:, :, None].float())
        ll = counts @ logq                                                # [S, N, 10] log10 likelihoods
        pl = torch.round(-10.0 * (ll - ll.max(dim=2, keepdim=True).values)).clamp_(0, 255).to(torch.uint8)
        rec = out_recs[s0:s0 + S]
        rec.zero_()
        has = depth > 0
        rec[:, :, 0:10] = pl * has[:, :, None]
        d = depth.to(torch.int32)
        rec[:, :, 10] = (d & 0xff).to(torch.uint8)
        rec[:, :, 11] = ((d >> 8) & 0xff).to(torch.uint8)
        rec[:, :, 12] = ((d >> 16) & 0xff).to(torch.uint8)
        rec[:, :, 13] = torch.where(has, torch.full_like(d, 100), torch.zeros_like(d)).to(torch.uint8)
        hdr = out_hdr[s0:s0 + S]
        hdr.zero_()
        pos = torch.arange(pos0 + s0, pos0 + s0 + S, device=dev, dtype=torch.int64)
        for b in range(4):
            hdr[:, b] = ((pos >> (8 * b)) & 0xff).to(torch.uint8)
        hdr[:, 4] = (ref + 1).to(torch.uint8)
    return out_hdr, out_recs
