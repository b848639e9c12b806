"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bit-exact: site status / filtering, argmax hypothesis, alleles, flags, best genotypes, read statistics.
Within 1e-6 relative (the tolerance BASELINE.json's north_star states): hypothesis log-likelihoods,
variant posterior, QUAL, allele frequency, de novo LR, AB, genotype posteriors, dosage.
"""
import os

import numpy as np
import pytest

import cli_util as U
import fixtures_util as F
import parity
from oracle_lib import OracleEngine
from polymutt_b200 import Engine, Params, capi, synth

pytestmark = pytest.mark.gpu

PED = lambda name: os.path.join(U.GOLDEN, "peds", name)

# (id, pedigree, params, number of sites (None = all 81,016), first site)
EXAMPLE_CASES = [
    ("quartets_cmd1", "test.ped", dict(posterior_cutoff=0.9, min_total_depth=150, max_total_depth=200), None, 0),
    ("mix_default", "test.mix.ped", dict(), None, 0),
    ("quartets_denovo", "test.ped", dict(denovo=True, denovo_mut_rate=1.5e-7), None, 0),
    ("mix_all_sites", "test.mix.ped", dict(out_all_sites=True), 30000, 20000),
    ("mix_strict", "test.mix.ped", dict(posterior_cutoff=0.99, min_map_quality=50, min_ps=90.0, theta=0.01, poly_tstv=3.0), None, 0),
    ("mix_denovo_loose", "test.mix.ped", dict(denovo=True, denovo_mut_rate=1e-4, denovo_min_llr=1e-3, denovo_tstv=1.0), None, 0),
    ("denovo_all_sites", "test.ped", dict(denovo=True, out_all_sites=True, denovo_min_llr=1e-9), 20000, 0),
    ("ext_ba", "ext.ped", dict(), None, 0),
    ("ext_denovo", "ext.ped", dict(denovo=True), 12000, 0),
    ("ceph_ba", "ceph.ped", dict(), 40000, 0),
    ("ceph_denovo", "ceph.ped", dict(denovo=True), 2500, 0),
    ("ceph_all_sites", "ceph.ped", dict(out_all_sites=True), 3000, 5000),
    # --quick_call: the everybody-unrelated pre-pass (main.cpp:354-437) decides which sites the real model sees
    ("quartets_quick", "test.ped", dict(quick_call=True), None, 0),
    ("mix_quick_denovo", "test.mix.ped", dict(quick_call=True, denovo=True, denovo_mut_rate=1e-4, denovo_min_llr=1e-3, denovo_tstv=1.0), None, 0),
    ("single_quick", "single.ped", dict(quick_call=True), 40000, 0),
    ("ext_quick", "ext.ped", dict(quick_call=True, posterior_cutoff=0.9), 30000, 0),
]


def _run_both(ped, params, hdr, recs):
    eng = Engine(ped, params)
    st_g, res_g, per_g = eng.call_glf_sites(hdr, recs, capi.PM_OUT_ALL)
    # the same sites as 14-byte wire records (pm_call_glf_sites_wire, widened on the device): identical bytes out
    st_w, res_w, per_w = eng.call_glf_sites_wire(hdr, recs, capi.PM_OUT_ALL)
    eng.close()
    assert np.array_equal(st_w, st_g) and res_w.tobytes() == res_g.tobytes() and per_w.tobytes() == per_g.tobytes()
    ora = OracleEngine(ped, params)
    st_o, res_o, per_o = ora.call_glf_sites(hdr, recs)
    ora.close()
    return (st_g, res_g, per_g), (st_o, res_o, per_o)


@pytest.mark.parametrize("case", EXAMPLE_CASES, ids=lambda c: c[0])
def test_example_data_parity(case, example12, oracle_built, tools_built, tmp_path):
    name, pedfile, kw, n_sites, start = case
    ped, glf_index = F.pedigree_from_file(PED(pedfile), str(tmp_path))
    hdr, recs = F.sites_for(example12, glf_index, n_sites, start)
    params = Params(**kw)
    g, o = _run_both(ped, params, hdr, recs)
    rep = parity.compare(*g, *o, denovo=params.denovo, label=name)
    print(rep)
    assert rep["emitted"] > 0
    parity.assert_parity(rep, len(hdr))


def _single_family(example12, cols, father, mother, sex, gen):
    ped = capi.PedigreeArrays(np.array([len(cols)]), np.array([sum(1 for f in father if f < 0)]), np.array([gen]),
                              np.array(sex, dtype=np.uint8), np.array(father), np.array(mother))
    hdr, recs = F.sites_for(example12, np.array(cols))
    return ped, hdr, recs


@pytest.mark.parametrize("kw", [dict(), dict(denovo=True, denovo_mut_rate=1.5e-7), dict(out_all_sites=True), dict(denovo=True, out_all_sites=True, denovo_min_llr=1e-9)],
                         ids=["ba", "denovo", "all_sites", "denovo_all_sites"])
def test_single_nuclear_family_fixed_prior(kw, example12, oracle_built):
    """One trio/quartet only: no Brent, the fixed parent-pair table (NucFam.cpp:383-420)."""
    ped, hdr, recs = _single_family(example12, [1, 2, 3, 4], [-1, -1, 0, 0], [-1, -1, 1, 1], [1, 2, 2, 1], 2)
    n = 25000
    params = Params(**kw)
    g, o = _run_both(ped, params, hdr[:n], recs[:n])
    rep = parity.compare(*g, *o, denovo=params.denovo, label="single_quartet")
    print(rep)
    parity.assert_parity(rep, n)


def test_unrelated_only(example12, oracle_built):
    ped = synth.families([1] * 12)
    hdr, recs = F.sites_for(example12, np.arange(1, 13))
    g, o = _run_both(ped, Params(), hdr, recs)
    rep = parity.compare(*g, *o, denovo=False, label="unrelated12")
    print(rep)
    parity.assert_parity(rep, len(hdr))


def test_edge_cases_missing_data_bad_ref_and_filters(example12, oracle_built, tools_built, tmp_path):
    ped, glf_index = F.pedigree_from_file(PED("test.mix.ped"), str(tmp_path))
    hdr, recs = F.sites_for(example12, glf_index, 6000, 1000)
    rng = np.random.default_rng(5)
    hdr["ref_base"][rng.integers(0, len(hdr), 300)] = 0          # N reference: skipped
    drop = rng.random(recs.shape) < 0.15                         # persons without a record at the site
    recs[drop] = np.zeros((), dtype=recs.dtype)
    recs[100:140] = np.zeros((), dtype=recs.dtype)               # sites where nobody has data
    recs["lk"][200:260, :, :] = 255                              # saturated likelihoods everywhere
    recs["depth"][300:330, :, 2] = 1                             # 24-bit depths (> 65535)
    for kw in (dict(), dict(min_total_depth=100, max_total_depth=170, min_ps=80.0, min_map_quality=90),
               dict(denovo=True, denovo_mut_rate=1e-6), dict(out_all_sites=True)):
        params = Params(**kw)
        g, o = _run_both(ped, params, hdr, recs)
        rep = parity.compare(*g, *o, denovo=params.denovo, label=f"edge{kw}")
        print(rep)
        parity.assert_parity(rep, len(hdr))
    # empty batch
    eng = Engine(ped, Params())
    st, res, per = eng.call_glf_sites(hdr[:0], recs[:0], capi.PM_OUT_ALL)
    assert len(st) == 0 and len(res) == 0
    eng.close()


def test_emitted_mode_is_the_ordered_subset_of_all_mode(example12, tools_built, tmp_path):
    ped, glf_index = F.pedigree_from_file(PED("test.ped"), str(tmp_path))
    hdr, recs = F.sites_for(example12, glf_index, 30000)
    eng = Engine(ped, Params())
    st_a, res_a, per_a = eng.call_glf_sites(hdr, recs, capi.PM_OUT_ALL)
    st_e, res_e, per_e = eng.call_glf_sites(hdr, recs, capi.PM_OUT_EMITTED)
    assert np.array_equal(st_a, st_e)
    idx = np.flatnonzero((st_a & 0xF) == capi.PM_SITE_EMITTED)
    assert len(res_e) == len(idx) > 100
    assert np.array_equal(res_e["site"], idx)
    for name in res_e.dtype.names:      # field by field: numpy fancy indexing does not preserve struct padding
        assert np.array_equal(res_e[name], res_a[idx][name], equal_nan=True), name
    for name in per_e.dtype.names:
        assert np.array_equal(per_e[name], per_a[idx][name], equal_nan=True), name
    # too small a result buffer is an error that reports the needed size
    with pytest.raises(RuntimeError, match="res_cap"):
        eng.call_glf_sites(hdr, recs, capi.PM_OUT_EMITTED, res_cap=10)
    eng.close()


def test_unknown_chromosome_class_fails_loudly(example12, tools_built, tmp_path):
    ped, glf_index = F.pedigree_from_file(PED("test.ped"), str(tmp_path))
    hdr, recs = F.sites_for(example12, glf_index, 100)
    hdr["chr_class"][50] = 7
    eng = Engine(ped, Params())
    with pytest.raises(RuntimeError, match="chr_class"):
        eng.call_glf_sites(hdr, recs, capi.PM_OUT_ALL)
    eng.close()


# ---- chrX / chrY / MT (SetPolyPrior_chrX.., SetParentPrior, likelihoodONEKid, GetTransmissionProb_BA) ----
# chr_class: 1, 2, 3 = the whole batch on that chromosome; -1 = every site draws its own class (0..3) and one site
# in eight carries PM_HDR_FIRST_POSTPROB, so one call exercises all rule sets side by side
NONAUTO_CASES = [(f"{pedfile}_{'dn' if kw.get('denovo') else 'ba'}_cls{cls}", pedfile, kw, n, cls)
                 for pedfile, n in (("test.ped", 30000), ("test.mix.ped", 30000), ("single.ped", 30000), ("ext.ped", 10000), ("ceph.ped", 3000))
                 for kw in (dict(), dict(denovo=True, denovo_mut_rate=1.5e-7))
                 for cls in (1, 2, 3, -1)
                 if not (kw.get("denovo") and pedfile in ("ext.ped", "ceph.ped") and cls in (2, 3))]


def _set_classes(hdr, cls, seed=11):
    if cls >= 0:
        hdr["chr_class"][:] = cls
        return
    rng = np.random.default_rng(seed)
    hdr["chr_class"][:] = rng.integers(0, 4, size=len(hdr))
    hdr["reserved"][:] = (rng.integers(0, 8, size=len(hdr)) == 0).astype(np.uint16)  # PM_HDR_FIRST_POSTPROB


@pytest.mark.parametrize("case", NONAUTO_CASES, ids=lambda c: c[0])
def test_sex_chromosome_and_mt_parity(case, example12, oracle_built, tools_built, tmp_path):
    name, pedfile, kw, n_sites, cls = case
    ped, glf_index = F.pedigree_from_file(PED(pedfile), str(tmp_path))
    if pedfile == "ceph.ped" and kw.get("denovo"):
        n_sites = 1200   # minutes of oracle time otherwise
    hdr, recs = F.sites_for(example12, glf_index, n_sites, 0)
    _set_classes(hdr, cls)
    params = Params(**kw)
    g, o = _run_both(ped, params, hdr, recs)
    rep = parity.compare(*g, *o, denovo=params.denovo, label=name)
    print(rep)
    assert rep["emitted"] > 0
    parity.assert_parity(rep, len(hdr))


# ---- wide kernel (one block per site) on synthetic pedigrees -------------------------------------
WIDE_CASES = [
    ("mixed70", lambda: synth.concat(synth.trios(40), synth.families([4] * 10 + [1] * 20)), dict(), 2500, 40.0),
    ("mixed70_denovo", lambda: synth.concat(synth.trios(40), synth.families([4] * 10 + [1] * 20)), dict(denovo=True), 1500, 40.0),
    ("trios300", lambda: synth.trios(300), dict(denovo=True), 600, 10.0),
    ("quartets_and_sibships", lambda: synth.families([4] * 100 + [5] * 40 + [6] * 10), dict(), 500, 10.0),
    ("trios1000_denovo", lambda: synth.trios(1000), dict(denovo=True), 200, 4.0),
    ("trios1000_ba", lambda: synth.trios(1000), dict(), 200, 4.0),
    # extended families among many units: the ES instances of the wide kernel
    ("trios20_ceph", lambda: synth.concat(synth.trios(20), synth.ceph()), dict(), 1500, 20.0),
    ("ceph_trios12_ceph_denovo", lambda: synth.concat(synth.ceph(5), synth.trios(12), synth.ceph(3)), dict(denovo=True), 250, 20.0),
    ("mixed70_quick", lambda: synth.concat(synth.trios(40), synth.families([4] * 10 + [1] * 20)), dict(quick_call=True), 2500, 40.0),
    ("trios300_quick_denovo", lambda: synth.trios(300), dict(quick_call=True, denovo=True), 400, 10.0),
]
# the same through the chrX / chrY / MT instance of the wide kernel (last field: chr_class, -1 = mixed per site)
WIDE_NONAUTO_CASES = [
    ("mixed70_x", lambda: synth.concat(synth.trios(40), synth.families([4] * 10 + [1] * 20)), dict(), 2500, 40.0, 1),
    ("mixed70_y", lambda: synth.concat(synth.trios(40), synth.families([4] * 10 + [1] * 20)), dict(), 2500, 40.0, 2),
    ("mixed70_mt", lambda: synth.concat(synth.trios(40), synth.families([4] * 10 + [1] * 20)), dict(), 2500, 40.0, 3),
    ("mixed70_denovo_any", lambda: synth.concat(synth.trios(40), synth.families([4] * 10 + [1] * 20)), dict(denovo=True), 1500, 40.0, -1),
    ("quartets_and_sibships_any", lambda: synth.families([4] * 100 + [5] * 40 + [6] * 10), dict(), 600, 10.0, -1),
    ("trios1000_any", lambda: synth.trios(1000), dict(), 240, 4.0, -1),
    ("trios1000_denovo_x", lambda: synth.trios(1000), dict(denovo=True), 160, 4.0, 1),
    ("trios20_ceph_any", lambda: synth.concat(synth.trios(20), synth.ceph()), dict(), 1200, 20.0, -1),
]


@pytest.mark.parametrize("case", WIDE_NONAUTO_CASES, ids=lambda c: c[0])
def test_wide_kernel_sex_chromosome_and_mt_parity(case, oracle_built):
    name, mk, kw, n_sites, boost, cls = case
    ped = mk()
    h, r = synth.generate_sites(ped, n_sites, seed=20261019, cfg=synth.SynthConfig(poly_boost=boost, injected_denovo=0.02))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1).copy()
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n_sites, ped.n_person)
    _set_classes(hdr, cls)
    params = Params(**kw)
    g, o = _run_both(ped, params, hdr, recs)
    rep = parity.compare(*g, *o, denovo=params.denovo, label=name)
    print(rep)
    assert rep["emitted"] > 0
    parity.assert_parity(rep, n_sites)


@pytest.mark.parametrize("case", WIDE_CASES, ids=lambda c: c[0])
def test_wide_kernel_parity_on_synthetic_pedigrees(case, oracle_built):
    name, mk, kw, n_sites, boost = case
    ped = mk()
    h, r = synth.generate_sites(ped, n_sites, seed=20261018, cfg=synth.SynthConfig(poly_boost=boost, injected_denovo=0.02))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n_sites, ped.n_person)
    params = Params(**kw)
    g, o = _run_both(ped, params, hdr, recs)
    rep = parity.compare(*g, *o, denovo=params.denovo, label=name)
    print(rep)
    assert rep["emitted"] > 0
    parity.assert_parity(rep, n_sites)


# ---- extended families beyond round 1's workspace (32 members / 8 marriage nodes) -------------------
BIG_ES_CASES = [
    ("ceph40_ba", lambda: synth.ceph(34), dict(), 4000, 30.0),                 # 40 members, 3 couples
    ("ceph40_denovo", lambda: synth.ceph(34), dict(denovo=True), 700, 30.0),
    ("clan42_ba", lambda: synth.clan(10, 2), dict(), 3000, 30.0),              # 42 members, 11 couples
    ("clan42_denovo", lambda: synth.clan(10, 2), dict(denovo=True), 400, 30.0),
    ("clan58_quartets_ba", lambda: synth.concat(synth.clan(14, 2), synth.families([4] * 12)), dict(), 1500, 30.0),  # 58 members, 15 couples, + 12 units: wide kernel
]


@pytest.mark.parametrize("case", BIG_ES_CASES, ids=lambda c: c[0])
def test_large_extended_families(case, oracle_built):
    name, mk, kw, n_sites, boost = case
    ped = mk()
    h, r = synth.generate_sites(ped, n_sites, seed=20261020, cfg=synth.SynthConfig(poly_boost=boost, injected_denovo=0.02))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n_sites, ped.n_person)
    params = Params(**kw)
    g, o = _run_both(ped, params, hdr, recs)
    rep = parity.compare(*g, *o, denovo=params.denovo, label=name)
    print(rep)
    assert rep["emitted"] > 0
    parity.assert_parity(rep, n_sites)


# ---- the block-per-site kernel on the edge cases, every instantiation ------------------------------
# (variant, threads) of pm_wide.cu's PM_WIDE_VARIANTS; test.mix.ped has 10 units, so every plan holds it
WIDE_PLANS = [(0, 32), (1, 32), (2, 32), (3, 32), (4, 64), (4, 128), (5, 256), (6, 128)]  # (variant 6: the site read from global memory)


def _run_wide(ped, params, hdr, recs, plan):
    eng = Engine(ped, params)
    eng.force_wide_plan(*plan)
    assert "k_sites_wide" in eng.describe_plan()
    g = eng.call_glf_sites(hdr, recs, capi.PM_OUT_ALL)
    eng.close()
    return g


@pytest.mark.parametrize("plan", WIDE_PLANS, ids=lambda p: f"v{p[0]}_t{p[1]}")
def test_wide_kernel_edge_cases_every_plan(plan, example12, oracle_built, tools_built, tmp_path):
    """The edge block of test_edge_cases_missing_data_bad_ref_and_filters (N reference, people and whole sites without
    data, saturated likelihoods, 24-bit depths, filters) through k_sites_wide: pm_force_wide_plan sends the 10-unit
    mixture pedigree to every instantiation instead of the thread-per-site kernel."""
    ped, glf_index = F.pedigree_from_file(PED("test.mix.ped"), str(tmp_path))
    hdr, recs = F.sites_for(example12, glf_index, 5000, 1000)
    rng = np.random.default_rng(5)
    hdr["ref_base"][rng.integers(0, len(hdr), 250)] = 0
    drop = rng.random(recs.shape) < 0.15
    recs[drop] = np.zeros((), dtype=recs.dtype)
    recs[100:140] = np.zeros((), dtype=recs.dtype)
    recs["lk"][200:260, :, :] = 255
    recs["depth"][300:330, :, 2] = 1
    for kw in (dict(), dict(min_total_depth=100, max_total_depth=170, min_ps=80.0, min_map_quality=90),
               dict(denovo=True, denovo_mut_rate=1e-6), dict(out_all_sites=True)):
        params = Params(**kw)
        g = _run_wide(ped, params, hdr, recs, plan)
        ora = OracleEngine(ped, params)
        o = ora.call_glf_sites(hdr, recs)
        ora.close()
        rep = parity.compare(*g, *o, denovo=params.denovo, label=f"wide-edge{plan}{kw}")
        print(rep)
        parity.assert_parity(rep, len(hdr))


@pytest.mark.parametrize("kw", [dict(), dict(denovo=True), dict(out_all_sites=True)], ids=["ba", "denovo", "all_sites"])
def test_wide_kernel_family_likelihood_underflow_gives_minus_infinity(kw, oracle_built):
    """Sibships of 14 and 16 with saturated likelihoods: 16+ factors of 10^-25.5 underflow, the family likelihood is
    exactly 0 in the reference and log10(0) = -inf runs through its objective (FamilyLikelihoodSeq.cpp:222-240,
    NucFamGenotypeLikelihood.cpp:941-985).  The kernel evaluates such families the reference's way ("fragile units")."""
    ped = synth.families([3] * 30 + [16, 18] + [1] * 6)
    n = 400
    h, r = synth.generate_sites(ped, n, seed=31, cfg=synth.SynthConfig(poly_boost=100.0, injected_denovo=0.05))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1).copy()
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n, ped.n_person).copy()
    first = int(ped.family_first()[30])
    rng = np.random.default_rng(3)
    sat = rng.random(n) < 0.5
    recs["lk"][sat, first:first + 34, :] = 255            # both big sibships saturated: every conditional underflows
    half = rng.random(n) < 0.3
    recs["lk"][half, first + 16:first + 34, :] = 254      # ... or only the second one, one phred off
    params = Params(**kw)
    g, o = _run_both(ped, params, hdr, recs)
    assert np.isneginf(o[1]["varllk"][:, 1]).sum() > 50, "the fixture is meant to make the reference's objective -inf"
    rep = parity.compare(*g, *o, denovo=params.denovo, label=f"underflow{kw}")
    print(rep)
    parity.assert_parity(rep, n)
    both_inf = np.isneginf(g[1]["varllk"][:, 1]) == np.isneginf(o[1]["varllk"][:, 1])
    assert both_inf.all()


@pytest.mark.parametrize("n_fam,kw", [(4400, dict(denovo=True)), (4500, dict()), (5000, dict(denovo=True))],
                         ids=["4400_trios_denovo", "4500_trios_ba", "5000_trios_denovo_site_in_global_memory"])
def test_wide_kernel_more_units_than_registers_hold(n_fam, kw, oracle_built):
    """More than 4,096 units: 512 threads x 8 units in registers, the rest in the L2 scratch.  A site of more than
    ~13,800 people (16 bytes each) does not fit in an SM's shared memory: its records are read from global memory."""
    ped = synth.trios(n_fam)
    n = 24
    h, r = synth.generate_sites(ped, n, seed=78, cfg=synth.SynthConfig(poly_boost=5.0, injected_denovo=0.1))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n, ped.n_person)
    params = Params(**kw)
    eng = Engine(ped, params)
    assert "in the L2 scratch" in eng.describe_plan() and f"{n_fam - 4096} in the L2" in eng.describe_plan()
    assert ("read from global memory" in eng.describe_plan()) == (n_fam == 5000), eng.describe_plan()
    g = eng.call_glf_sites(hdr, recs, capi.PM_OUT_ALL)
    eng.close()
    ora = OracleEngine(ped, params)
    o = ora.call_glf_sites(hdr, recs)
    ora.close()
    rep = parity.compare(*g, *o, denovo=params.denovo, label=f"{n_fam} trios")
    print(rep)
    parity.assert_parity(rep, n)


# ---- VCF-input records through the C ABI ---------------------------------------------------------
def _vcf_records_from_sites(hdr, recs, rng):
    """Turns packed GLF sites into VCF-mode records: (REF, ALT) = (ref, a random other base), the three PLs of that
    allele pair kept, every other byte cleared; mono = sum of -PL[ref/ref]/10 in column order."""
    n, npers = recs.shape
    gi = lambda a, b: np.where(a < b, (a - 1) * (10 - a) // 2 + (b - a), (b - 1) * (10 - b) // 2 + (a - b))
    ref = hdr["ref_base"].astype(int)
    alt = (ref - 1 + rng.integers(1, 4, n)) % 4 + 1
    out = np.zeros_like(recs)
    g0, g1, g2 = gi(ref, ref), gi(ref, alt), gi(alt, alt)
    rows = np.arange(n)[:, None]
    for g in (g0, g1, g2):
        out["lk"][rows, np.arange(npers)[None, :], g[:, None]] = recs["lk"][rows, np.arange(npers)[None, :], g[:, None]]
    out["depth"] = recs["depth"]
    h = hdr.copy()
    h["reserved"] = alt.astype(np.uint16)
    h["reserved"][::17] |= 0x100                      # some records flagged as indels (prior term only)
    mono = np.zeros(n)
    for c in range(npers):                            # sequential sum in column order, as the host front end does
        mono += -out["lk"][np.arange(n), c, g0].astype(np.float64) / 10.0
    return h, out, mono


@pytest.mark.parametrize("shape,n", [("fam200x5", 1500), ("trios334", 800), ("mixed_350", 2500)])
def test_vcf_records_parity_wide_plans(shape, n, oracle_built):
    """BASELINE config 4's own shape (200 nuclear families x 5 = 1,000 samples: one warp x 8 units per record) and two
    more many-unit pedigrees through pm_call_vcf_records, against the oracle."""
    ped = {"fam200x5": lambda: synth.families([5] * 200), "trios334": lambda: synth.trios(334),
           "mixed_350": lambda: synth.concat(synth.trios(50), synth.families([4] * 50))}[shape]()
    hh, rr = synth.generate_sites(ped, n, seed=41, cfg=synth.SynthConfig(poly_boost=150.0, injected_denovo=0.02))
    hdr = hh.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1).copy()
    recs = rr.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n, ped.n_person).copy()
    h, r, mono = _vcf_records_from_sites(hdr, recs, np.random.default_rng(4))
    lut = np.array([pow(10, -float(i) / 10.0) for i in range(256)])
    params = Params(vcf_input=True)
    eng = Engine(ped, params, lut=lut)
    assert "k_sites_wide" in eng.describe_plan()
    res_g, per_g = eng.call_vcf_records(h, r, mono)
    res_c, calls = eng.call_vcf_records_calls(h, r, mono)   # the compact form: best | gq << 8, written by k_post directly
    res_p, calls_p = eng.call_vcf_records_pl(h, r, mono)    # the same from three PL bytes per sample (what the executable sends)
    eng.close()
    assert np.array_equal(calls, (per_g["best"].astype(np.uint16) & 0xff) | (per_g["gq"].astype(np.uint16) << 8))
    assert np.array_equal(calls_p, calls)
    for f in ("poly_qual", "freq", "varllk", "varllk_noprior", "allele1", "allele2", "site"):
        assert np.array_equal(res_c[f], res_g[f]) and np.array_equal(res_p[f], res_g[f]), f
    ora = OracleEngine(eng.ped, params, lut=lut)
    res_o, per_o = ora.call_vcf_records(h, r, mono)
    ora.close()
    close = lambda a, b, rt=1e-6, at=0.0: parity._close(a, b, rt, at)
    freq_bad = ~close(res_g["freq"], res_o["freq"], 1e-6, 1e-9)
    same_optimum = close(res_g["varllk_noprior"][:, 1], res_o["varllk_noprior"][:, 1], 1e-12, 1e-11)
    knife = freq_bad & same_optimum
    ok = ~knife
    bad = {
        "llk_ref": int(np.sum(~close(res_g["varllk"][:, 0], res_o["varllk"][:, 0]))),
        "llk_alt": int(np.sum(~close(res_g["varllk"][:, 1], res_o["varllk"][:, 1]))),
        "qual": int(np.sum(~close(res_g["poly_qual"], res_o["poly_qual"], 1e-6, 1e-5))),
        "freq_not_knife_edge": int(np.sum(freq_bad & ~same_optimum)),
        "best": int(np.sum(per_g["best"][ok] != per_o["best"][ok])),
        "post": int(np.sum(~close(per_g["post"][ok], per_o["post"][ok], 1e-6, 1e-15))),
        "gq": int(np.sum(np.abs(per_g["gq"][ok].astype(int) - per_o["gq"][ok].astype(int)) > 1)),
    }
    print(shape, bad, "knife-edge records:", int(knife.sum()), "of", n)
    assert not any(bad.values()), bad
    assert knife.sum() <= max(2, n // 1000), int(knife.sum())


@pytest.mark.parametrize("pedfile,n,mixed_classes", [("test.ped", 20000, False), ("test.mix.ped", 20000, False), ("single.ped", 20000, False),
                                                    ("ext.ped", 8000, False), ("ceph.ped", 4000, False),
                                                    # every record draws its chromosome class: chrX / chrY / MT records go through the
                                                    # all-families-peeled description (FLSeq_VCF.cpp:101, 148), autosomal ones do not
                                                    ("test.ped", 12000, True), ("test.mix.ped", 12000, True), ("single.ped", 12000, True),
                                                    ("ext.ped", 6000, True), ("mixext.ped", 6000, True)])
def test_vcf_records_parity(pedfile, n, mixed_classes, example12, oracle_built, tools_built, tmp_path):
    ped, glf_index = F.pedigree_from_file(PED(pedfile), str(tmp_path))
    hdr, recs = F.sites_for(example12, glf_index, n)
    h, r, mono = _vcf_records_from_sites(hdr, recs, np.random.default_rng(3))
    if mixed_classes:
        h["chr_class"][:] = np.random.default_rng(5).integers(0, 4, size=len(h))
    lut = np.array([pow(10, -float(i) / 10.0) for i in range(256)])
    params = Params(vcf_input=True)
    eng = Engine(ped, params, lut=lut)
    res_g, per_g = eng.call_vcf_records(h, r, mono)
    res_p, calls_p = eng.call_vcf_records_pl(h, r, mono)    # three PL bytes per sample in, two bytes per sample out
    eng.close()
    assert np.array_equal(calls_p, (per_g["best"].astype(np.uint16) & 0xff) | (per_g["gq"].astype(np.uint16) << 8))
    for f in ("poly_qual", "freq", "varllk", "allele1", "allele2", "site"):
        assert np.array_equal(res_p[f], res_g[f]), f
    ora = OracleEngine(eng.ped, params, lut=lut)
    res_o, per_o = ora.call_vcf_records(h, r, mono)
    ora.close()
    close = lambda a, b, rt=1e-6, at=0.0: parity._close(a, b, rt, at)
    # Records whose allele-frequency optimum is not unique to ~1e-13 in the objective (uninformative or exactly
    # symmetric data) are knife-edge for Brent: a one-ulp difference in f(p) changes the path, on the CPU between
    # compilers as much as on the GPU.  They are identified by the two optima being equally good, reported, bounded,
    # and excluded from the per-person comparison; everything else must agree to 1e-6.
    freq_bad = ~close(res_g["freq"], res_o["freq"], 1e-6, 1e-9)
    same_optimum = close(res_g["varllk_noprior"][:, 1], res_o["varllk_noprior"][:, 1], 1e-12, 1e-11)
    knife = freq_bad & same_optimum
    ok = ~knife
    for i in np.flatnonzero(freq_bad)[:6]:
        print("freq mismatch at record", i, "gpu", res_g["freq"][i], "oracle", res_o["freq"][i], "maxlogL gpu-oracle",
              res_g["varllk_noprior"][i, 1] - res_o["varllk_noprior"][i, 1], "alleles", res_g["allele1"][i], res_g["allele2"][i],
              "PL rows", r["lk"][i][:, :].tolist()[:4])
    bad = {
        "llk_ref": int(np.sum(~close(res_g["varllk"][:, 0], res_o["varllk"][:, 0]))),
        "llk_alt": int(np.sum(~close(res_g["varllk"][:, 1], res_o["varllk"][:, 1]))),
        "qual": int(np.sum(~close(res_g["poly_qual"], res_o["poly_qual"], 1e-6, 1e-5))),
        "freq_not_knife_edge": int(np.sum(freq_bad & ~same_optimum)),
        "best": int(np.sum(per_g["best"][ok] != per_o["best"][ok])),
        "post": int(np.sum(~close(per_g["post"][ok], per_o["post"][ok], 1e-6, 1e-15))),
        "gq": int(np.sum(np.abs(per_g["gq"][ok].astype(int) - per_o["gq"][ok].astype(int)) > 1)),
    }
    print(pedfile, bad, "knife-edge records:", int(knife.sum()), "of", n)
    assert not any(bad.values()), bad
    # (mixext.ped with a random chromosome class per record: on chrY the females' data drops out and families without an
    # informative male leave the objective exactly flat: more such records than anywhere else)
    assert knife.sum() <= max(2, n // (300 if (pedfile, mixed_classes) == ("mixext.ped", True) else 1000)), int(knife.sum())
