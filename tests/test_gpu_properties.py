"""Size-independent properties of the CUDA path at the bench workload's shape (1,000 trios, --denovo), where the
CPU oracle is too slow to check every site: sites are independent, so results must not depend on how the
job is cut into calls, on the order of the sites, or on which entry point (host buffers / device buffers) is used;
and the call must be deterministic.  A sample of the sites is still checked against the oracle."""
import numpy as np
import pytest
import torch

import parity
from oracle_lib import OracleEngine
from polymutt_b200 import Engine, Params, capi, synth

pytestmark = pytest.mark.gpu

N_SITES = 24000


@pytest.fixture(scope="module")
def workload():
    ped = synth.trios(1000)
    dev = torch.device("cuda", 0)
    h, r = synth.generate_sites(ped, N_SITES, seed=424242, device=dev, chunk=4096,
                                cfg=synth.SynthConfig(poly_boost=3.0, injected_denovo=0.01))
    hdr = h.cpu().numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.cpu().numpy().view(capi.PERSON_SITE_DTYPE).reshape(N_SITES, ped.n_person)
    return ped, hdr, recs, h, r


def _fields_equal(a, b):
    return all(np.array_equal(a[k], b[k], equal_nan=True) for k in a.dtype.names)


def test_split_permutation_determinism_and_entry_points(workload):
    ped, hdr, recs, d_hdr, d_recs = workload
    params = Params(denovo=True)
    eng = Engine(ped, params)
    st, res, per = eng.call_glf_sites(hdr, recs, capi.PM_OUT_ALL)
    emitted = np.flatnonzero((st & 0xF) == capi.PM_SITE_EMITTED)
    assert len(emitted) >= 5 and (st & 0xF).max() <= capi.PM_SITE_DENOVO_DROPPED
    # determinism
    st2, res2, per2 = eng.call_glf_sites(hdr, recs, capi.PM_OUT_ALL)
    assert np.array_equal(st, st2) and _fields_equal(res, res2) and _fields_equal(per[emitted], per2[emitted])
    # cutting the job into three calls of awkward sizes
    cuts = [0, 7001, 7002, N_SITES]
    for a, b in zip(cuts[:-1], cuts[1:]):
        s3, r3, p3 = eng.call_glf_sites(hdr[a:b], recs[a:b], capi.PM_OUT_ALL)
        r3 = r3.copy(); r3["site"] += a
        assert np.array_equal(s3, st[a:b]) and _fields_equal(r3, res[a:b])
        em = np.flatnonzero((s3 & 0xF) == 0)
        assert _fields_equal(p3[em], per[a:b][em])
    # site order does not matter
    perm = np.random.default_rng(1).permutation(N_SITES)
    s4, r4, p4 = eng.call_glf_sites(hdr[perm], recs[perm], capi.PM_OUT_ALL)
    r4 = r4.copy(); r4["site"] = res["site"][perm]
    assert np.array_equal(s4, st[perm]) and _fields_equal(r4, res[perm])
    # device-buffer entry point == host-buffer entry point (PM_OUT_EMITTED, compacted in site order)
    dev = d_hdr.device
    cap = len(emitted) + 8
    d_status = torch.empty(N_SITES, dtype=torch.uint16, device=dev)
    d_res = torch.zeros((cap, 256), dtype=torch.uint8, device=dev)
    d_per = torch.zeros((cap, ped.n_person, 96), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    eng.call_glf_sites_device(d_hdr.data_ptr(), d_recs.data_ptr(), N_SITES, capi.PM_OUT_EMITTED, d_status.data_ptr(), d_res.data_ptr(),
                              d_per.data_ptr(), cap, d_n.data_ptr())
    eng.sync()
    assert int(d_n.item()) == len(emitted)
    assert np.array_equal(d_status.cpu().numpy(), st)
    r5 = d_res.cpu().numpy().view(capi.SITE_RESULT_DTYPE).reshape(-1)[:len(emitted)]
    p5 = d_per.cpu().numpy().view(capi.PERSON_RESULT_DTYPE).reshape(cap, ped.n_person)[:len(emitted)]
    assert np.array_equal(r5["site"], emitted) and _fields_equal(r5, res[emitted]) and _fields_equal(p5, per[emitted])
    # host buffers, PM_OUT_EMITTED: 23 chunks in flight on three streams, the rows of all of them compacted in site order
    # -- as 16-byte records and as 14-byte wire records
    for call in (eng.call_glf_sites, eng.call_glf_sites_wire):
        s6, r6, p6 = call(hdr, recs, capi.PM_OUT_EMITTED, res_cap=cap)
        assert np.array_equal(s6, st) and len(r6) == len(emitted)
        assert np.array_equal(r6["site"], emitted) and _fields_equal(r6, res[emitted]) and _fields_equal(p6, per[emitted])
    with pytest.raises(RuntimeError, match="too small"):
        eng.call_glf_sites(hdr, recs, capi.PM_OUT_EMITTED, res_cap=3)
    s7, r7, p7 = eng.call_glf_sites(hdr[:5000], recs[:5000], capi.PM_OUT_EMITTED, res_cap=cap)   # the ctx is still good after the error
    k7 = int(np.sum(emitted < 5000))
    assert np.array_equal(s7, st[:5000]) and _fields_equal(r7, res[emitted[:k7]])
    # a result buffer that is too small is reported, not overrun
    d_n.zero_()
    eng.call_glf_sites_device(d_hdr.data_ptr(), d_recs.data_ptr(), N_SITES, capi.PM_OUT_EMITTED, d_status.data_ptr(), d_res.data_ptr(),
                              d_per.data_ptr(), 3, d_n.data_ptr())
    eng.sync()
    assert int(d_n.item()) == len(emitted)   # the count still says how many rows were due
    c = eng.counters()
    assert c["sites_evaluated"] > 0 and c["evaluations"] > 20 * c["sites_evaluated"] / 2
    eng.close()


def test_sampled_sites_match_oracle(workload, oracle_built):
    ped, hdr, recs, _, _ = workload
    idx = np.sort(np.random.default_rng(9).choice(N_SITES, 150, replace=False))
    params = Params(denovo=True)
    eng = Engine(ped, params)
    g = eng.call_glf_sites(hdr[idx], recs[idx], capi.PM_OUT_ALL)
    eng.close()
    ora = OracleEngine(ped, params)
    o = ora.call_glf_sites(hdr[idx], recs[idx])
    ora.close()
    rep = parity.compare(*g, *o, denovo=True, label="bench-shape sample")
    print(rep)
    parity.assert_parity(rep, len(idx))


@pytest.mark.parametrize("n_fam,kw", [(2048, dict()), (1500, dict(denovo=True))], ids=["2048_trios_ba", "1500_quartets_denovo"])
def test_large_pedigrees(n_fam, kw, oracle_built):
    """Upper end of the wide kernel's plans (more than 1,024 units -> 256 threads x 8 units)."""
    ped = synth.trios(n_fam) if not kw else synth.families([4] * n_fam)
    n = 40
    h, r = synth.generate_sites(ped, n, seed=77, cfg=synth.SynthConfig(poly_boost=5.0, injected_denovo=0.05))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n, ped.n_person)
    params = Params(**kw)
    eng = Engine(ped, params)
    assert "k_sites_wide" in eng.describe_plan()
    g = eng.call_glf_sites(hdr, recs, capi.PM_OUT_ALL)
    eng.close()
    ora = OracleEngine(ped, params)
    o = ora.call_glf_sites(hdr, recs)
    ora.close()
    rep = parity.compare(*g, *o, denovo=params.denovo, label=f"{n_fam} families")
    print(rep)
    parity.assert_parity(rep, n)


def test_unsupported_shapes_fail_loudly():
    # more than 512 threads x (8 + 32) units: the per-thread mask of fragile units in the L2 scratch is 32 bits wide
    # (a site of more than ~13,800 people no longer fits in shared memory either, but that is served from global memory:
    # test_wide_kernel_more_units_than_registers_hold)
    with pytest.raises(RuntimeError, match="not supported"):
        Engine(synth.trios(21000), Params())
    with pytest.raises(RuntimeError, match="quick_call"):
        Engine(synth.trios(3), Params(quick_call=True, vcf_input=True))
