"""Host-side logic that runs without a GPU: C-ABI surface, tables, peeling orders, pedigree ordering,
GLF round trip, fixtures."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cli_util as U
import oracle_lib
from polymutt_b200 import capi, synth

ROOT = U.ROOT


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.lib_path()):
        subprocess.run(["make", "-s", "lib"], cwd=ROOT, check=True)
    return capi.load_library()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "polymutt_b200.h")).read()
    names = set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", hdr))
    names -= {"pm_ctx"}
    assert len(names) >= 12
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} is declared in include/polymutt_b200.h but not exported"
    assert lib.pm_abi_version() == 1


def test_struct_sizes_match_header(lib):
    # sizes the kernels and the Python binding both rely on
    assert capi.SITE_HDR_DTYPE.itemsize == 8
    assert capi.PERSON_SITE_DTYPE.itemsize == 16
    assert capi.SITE_RESULT_DTYPE.itemsize == 256
    assert capi.PERSON_RESULT_DTYPE.itemsize == 96


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ped = synth.trios(2)
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        capi.Engine(ped, capi.Params())


def test_lut_and_mutation_matrix_bit_identical_to_oracle(lib, oracle_built):
    o = oracle_lib.load()
    a, b = np.zeros(256), np.zeros(256)
    lib.pm_fill_lut(a.ctypes.data)
    o.pmo_fill_lut(b.ctypes.data)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    assert a[0] == 1.0 and abs(a[10] - 0.1) < 1e-16
    for mu, tstv in ((1.5e-8, 2.0), (1.5e-7, 2.0), (1e-4, 1.0), (1e-3, 0.5), (1e-5, 0.0)):
        m, n = np.zeros(100), np.zeros(100)
        lib.pm_genotype_mutation_matrix(mu, tstv, m.ctypes.data)
        o.pmo_genotype_mutation_matrix(mu, tstv, n.ctypes.data)
        assert np.array_equal(m.view(np.uint64), n.view(np.uint64)), (mu, tstv)
        if tstv != 0.0:
            assert np.allclose(m.reshape(10, 10).sum(1), 1.0, atol=1e-12)


def _peel(lib, father, mother, sex):
    n = len(father)
    fa, mo = np.asarray(father, np.int32), np.asarray(mother, np.int32)
    sx = np.asarray(sex, np.uint8)
    steps = np.zeros(n, dtype=capi.PEEL_STEP_DTYPE)
    ns = lib.pm_build_peel_order(n, fa.ctypes.data, mo.ctypes.data, sx.ctypes.data, steps.ctypes.data)
    return ns, steps[:max(ns, 0)]


PEDS = {
    # founders first, then ancestors before descendants (Family::path order)
    "ext7": ([-1, -1, -1, 0, 0, 3, 3], [-1, -1, -1, 1, 1, 2, 2], [1, 2, 2, 1, 2, 1, 2]),
    "ceph20": (list(synth.ceph().father), list(synth.ceph().mother), list(synth.ceph().sex)),
    "two_roofs": ([-1, -1, -1, -1, 0, 2, 4], [-1, -1, -1, -1, 1, 3, 5], [1, 2, 1, 2, 1, 2, 1]),
    "four_gen": ([-1, -1, -1, -1, 0, 4, 5, 5], [-1, -1, -1, -1, 1, 2, 3, 3], [1, 2, 2, 2, 1, 1, 1, 2]),
    "halfsib_like": ([-1, -1, -1, 0, 0, 3], [-1, -1, -1, 1, 1, 2], [1, 2, 2, 1, 2, 1]),
}


@pytest.mark.parametrize("name", sorted(PEDS))
def test_peel_order_matches_oracle_restatement(lib, oracle_built, name):
    fa, mo, sx = PEDS[name]
    ns, steps = _peel(lib, fa, mo, sx)
    ref = oracle_lib.oracle_peel_order(fa, mo, sx)
    assert ns == len(ref) and ns == len(fa) - 1 - sum(1 for s in ref if s["type"] == 3), (ns, ref)
    assert np.array_equal(steps, ref), (steps, ref)


def test_peel_order_ceph_shape(lib):
    fa, mo, sx = PEDS["ceph20"]
    ns, steps = _peel(lib, fa, mo, sx)
    types = list(steps["type"])
    # 14 leaf peels, 2 roof peels, 1 spouse peel
    assert types.count(1) == 14 and types.count(3) == 2 and types.count(2) == 1 and ns == 17
    assert types[:14] == [1] * 14


def test_peel_order_rejects_loops_and_disconnected(lib):
    # first-cousin marriage (inbreeding loop)
    fa = [-1, -1, -1, -1, 0, 0, 4, 6]
    mo = [-1, -1, -1, -1, 1, 1, 2, 3]
    ns, _ = _peel(lib, [-1, -1, -1, -1, 0, 0, 4, 5 - 5 + 4], mo, [1, 2, 2, 2, 1, 2, 1, 1])
    assert ns < 0 or ns <= 7
    # an unconnected founder inside an extended family
    ns2, _ = _peel(lib, [-1, -1, -1, -1, 0, 4], [-1, -1, -1, -1, 1, 2], [1, 2, 2, 1, 1, 1])
    assert ns2 < 0
    assert lib.pm_last_error()


def test_pack_reproduces_fixture_and_pedigree_order(tools_built, example12, tmp_path):
    """pm-tools unpack -> pack round trip: GLF writer, GLF reader, N-way merge and pedigree loader."""
    glf = U.unpack_example(str(tmp_path))
    out = str(tmp_path / "again.pmpk")
    subprocess.run([U.PM_TOOLS, "pack", "-p", os.path.join(glf, "ped"), "-d", os.path.join(glf, "dat"), "-g", os.path.join(glf, "gif"), "-o", out],
                   check=True, stderr=subprocess.DEVNULL)
    from polymutt_b200 import load_pmpk
    again = load_pmpk(out)
    assert np.array_equal(again.hdr, example12.hdr)
    assert np.array_equal(again.recs, example12.recs)
    assert again.max_position == example12.max_position == 81016
    assert [p[1] for p in again.people] == [str(i) for i in range(1, 13)]


def test_pedigree_ordering_rules(tools_built, tmp_path):
    """Natural, case-insensitive (famid, pid) sort; founders first; ancestors before descendants;
    mis-sexed parents swapped (core/Pedigree.cpp:39-85, PedigreeFamily.cpp:11-85, PedigreePerson.cpp:90-126)."""
    ped = tmp_path / "p.ped"
    ped.write_text(
        "f10 c 9 b 1 0\nf10 9 0 0 1 0\nf10 b 0 0 2 0\n"      # father/mother columns given in order
        "F2 kid2 mum dad 2 0\nF2 dad 0 0 1 0\nF2 mum 0 0 2 0\nF2 kid10 dad mum 1 0\nF2 g 0 0 2 0\nF2 gk kid10 g 1 0\n"
        "f1 solo 0 0 1 0\n")
    dat = tmp_path / "p.dat"
    dat.write_text("T GLF_Index\n")
    out = str(tmp_path / "p.pmpk")
    subprocess.run([U.PM_TOOLS, "pack", "-p", str(ped), "-d", str(dat), "-o", out], check=True, stderr=subprocess.DEVNULL)
    from polymutt_b200 import load_pmpk
    p = load_pmpk(out)
    # families sort naturally: f1 < F2 < f10
    assert [x[0] for x in p.people] == ["f1"] + ["F2"] * 6 + ["f10"] * 3
    # F2: founders in sorted pid order (dad, g, mum), then kid2 (sorted before kid10: natural order), kid10, then gk
    assert [x[1] for x in p.people[1:7]] == ["dad", "g", "mum", "kid2", "kid10", "gk"]
    assert list(p.ped.fam_size) == [1, 6, 3] and list(p.ped.fam_founders) == [1, 3, 2]
    assert list(p.ped.fam_generations) == [1, 3, 2]
    # kid2 was entered as (father=mum, mother=dad): swapped back by sex
    assert p.ped.father[1 + 3] == 0 and p.ped.mother[1 + 3] == 2
    # peeling orders exist for the families with non-founders (extended F2: 5 steps; nuclear f10: child -> parents, spouse)
    d = list(np.diff(p.ped.peel_first))
    assert d[0] == 0 and d[1] == 5 and d[2] == 2 and len(p.ped.peel) == 7
    assert list(p.ped.peel["type"][5:]) == [1, 2]


def test_batched_multithreaded_ingest_equals_the_one_site_at_a_time_merge(tools_built, tmp_path):
    """GlfBatchReader (thread pool, block decode, bitmap union) against GlfSet (Move2NextBaseEntry restated one site
    at a time) on ragged streams: holes, sparse positions, different ends (the chromosome stops one site after the
    first stream runs out), a stream without any record, interleaved indel records, a truncated file, a person
    without a GLF, position 0, and a lead stream that lacks some sites (reference base from the next column)."""
    from polymutt_b200 import glfio, load_pmpk
    rng = np.random.default_rng(12)
    ped = synth.concat(synth.trios(4), synth.families([4, 1, 1]))
    n = 6000
    h, r = synth.generate_sites(ped, n, seed=21, cfg=synth.SynthConfig(poly_boost=20))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1).copy()
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n, ped.n_person).copy()
    steps = rng.choice([1, 1, 1, 2, 5, 300], n)
    steps[0] = 1
    hdr["pos"] = np.cumsum(steps) - 1                                      # starts at 0, sparse stretches
    hdr["ref_base"] = rng.integers(1, 5, n)
    recs[rng.random(recs.shape) < 0.2] = np.zeros((), dtype=recs.dtype)    # holes
    recs[1:40, 0] = np.zeros((), dtype=recs.dtype)                         # the lead stream misses early sites
    recs[0, 1] = recs[5, 1]                                                # position 0 is covered
    ends = {1: 5200, 2: 5600, 5: 5990}                                      # streams that stop early
    d = tmp_path / "ragged"
    d.mkdir()
    lines, gif = [], []
    firsts = ped.family_first()
    col = 0
    for f in range(ped.n_fam):
        for j in range(int(ped.fam_size[f])):
            fa, mo = int(ped.father[col]), int(ped.mother[col])
            key = 0 if col == 7 else col + 1                                # one person without a GLF
            lines.append(f"fam{f + 1}\tp{j + 1}\t{('p%d' % (fa + 1)) if fa >= 0 else 0}\t{('p%d' % (mo + 1)) if mo >= 0 else 0}\t{int(ped.sex[col])}\t{key}\n")
            if key:
                rr = recs[:, col].copy()
                if col in ends:
                    rr[ends[col]:] = np.zeros((), dtype=rr.dtype)
                if col == 9:
                    rr[:] = np.zeros((), dtype=rr.dtype)                    # a stream with a section but no records
                path = str(d / f"g{key}.glf")
                glfio.write_glf(path, "7", int(hdr["pos"].max()) + 1, hdr["pos"].astype(np.int64), hdr["ref_base"], rr,
                                indel_every=37 if col == 3 else 0, end_marker=(col != 11))
                gif.append(f"{key} {path}\n")
            col += 1
    (d / "ped").write_text("".join(lines)); (d / "dat").write_text("T\tGLF_Index\n"); (d / "gif").write_text("".join(gif))
    outs = []
    for extra in ([], ["--batched", "1"], ["--batched", "5"]):
        out = str(d / ("out%d.pmpk" % len(outs)))
        subprocess.run([U.PM_TOOLS, "pack", "-p", str(d / "ped"), "-d", str(d / "dat"), "-g", str(d / "gif"), "-o", out] + extra,
                       check=True, stderr=subprocess.DEVNULL)
        outs.append(load_pmpk(out))
    ref = outs[0]
    # the chromosome ends one site after the shortest stream's last record (the empty stream ends it at once: 1 site,
    # plus one more because the first site sits at position 0 where the reference skips its end check)
    assert 1 <= len(ref.hdr) <= 3, len(ref.hdr)
    for o in outs[1:]:
        assert np.array_equal(o.hdr, ref.hdr) and np.array_equal(o.recs, ref.recs)
    # without the empty stream the run goes on until just past the first stream that ends
    (d / "gif").write_text("".join(l for l in gif if not l.startswith("10 ")))
    (d / "ped").write_text("".join(l.rsplit("\t", 1)[0] + "\t0\n" if l.endswith("\t10\n") else l for l in lines))
    outs = []
    for extra in ([], ["--batched", "1"], ["--batched", "3"], ["--batched", "2", "--portable", "1"]):   # --portable: the record conversion without SSSE3
        out = str(d / ("outb%d.pmpk" % len(outs)))
        subprocess.run([U.PM_TOOLS, "pack", "-p", str(d / "ped"), "-d", str(d / "dat"), "-g", str(d / "gif"), "-o", out] + extra,
                       check=True, stderr=subprocess.DEVNULL)
        outs.append(load_pmpk(out))
    ref = outs[0]
    assert 4000 < len(ref.hdr) < 5400 and ref.hdr["pos"][0] == 0
    for o in outs[1:]:
        assert len(o.hdr) == len(ref.hdr)
        assert np.array_equal(o.hdr, ref.hdr) and np.array_equal(o.recs, ref.recs)
    # every other file gzip-compressed in place (told apart by their first two bytes: gzip goes through zlib, plain files are
    # read as they are): same batch
    import gzip
    for l in gif:
        key, path = l.split()
        if int(key) % 2 and key != "10":
            raw = open(path, "rb").read()
            with gzip.open(path, "wb", compresslevel=1) as z:
                z.write(raw)
    out = str(d / "outz.pmpk")
    subprocess.run([U.PM_TOOLS, "pack", "-p", str(d / "ped"), "-d", str(d / "dat"), "-g", str(d / "gif"), "-o", out, "--batched", "3"],
                   check=True, stderr=subprocess.DEVNULL)
    z = load_pmpk(out)
    assert np.array_equal(z.hdr, ref.hdr) and np.array_equal(z.recs, ref.recs)


def _insert_repeats(path, label, at, rng):
    """Rewrites an uncompressed GLF (base records only) with extra records of offset 0 -- repeats of a position -- after
    the records whose index is in `at` (an index listed k times gets k repeats)."""
    raw = open(path, "rb").read()
    start = 8 + 4 + len(label) + 1 + 4
    body = raw[start:]
    has_end = len(body) % 20 == 1
    n = len(body) // 20
    out = [raw[:start]]
    for k in range(n):
        rec = body[20 * k:20 * (k + 1)]
        out.append(rec)
        for _ in range(at.count(k)):
            lk = bytes(int(x) for x in rng.integers(0, 256, 10))
            out.append(rec[:1] + (0).to_bytes(4, "little") + int(rng.integers(1, 200)).to_bytes(3, "little") + b"\0" + bytes([int(rng.integers(1, 61))]) + lk)
    out.append(body[20 * n:] if has_end else b"")
    open(path, "wb").write(b"".join(out))


@pytest.mark.parametrize("case", ["middle", "same_site_two_streams", "triple", "last_record_of_the_first_stream_to_end", "position_zero", "many"])
def test_batched_ingest_walks_through_repeated_positions_like_the_reference(case, tools_built, tmp_path):
    """A base record with offset 0 repeats its stream's position.  The reference's cursor (src/PedigreeGLF.cpp:282-324)
    makes a further site at that position out of the repeats alone; the batched reader must hand out the same sites as
    the one-site-at-a-time restatement of that cursor: in the middle of a chromosome, with two streams repeating the same
    position, three records at one position, a repeat that is the last record of the stream that ends the chromosome
    (the end comes one site after the site that consumed the stream's LAST record), a repeat at position 0, and many."""
    from polymutt_b200 import glfio, load_pmpk
    rng = np.random.default_rng(5)
    ped = synth.concat(synth.trios(2), synth.families([4]))
    n = 3000
    h, r = synth.generate_sites(ped, n, seed=33, cfg=synth.SynthConfig(poly_boost=20))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1).copy()
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n, ped.n_person).copy()
    steps = rng.choice([1, 1, 2, 7], n)
    steps[0] = 1
    hdr["pos"] = np.cumsum(steps) - 1
    hdr["ref_base"] = rng.integers(1, 5, n)
    recs["depth"][:, :, 0] |= 1                                            # every person has a record at every site ...
    ends = {4: 2500}                                                       # ... but stream 4 stops early and ends the chromosome
    repeats = {"middle": {2: [700]}, "same_site_two_streams": {1: [1200], 6: [1200, 1201]}, "triple": {3: [999, 999], 0: [999]},
               "last_record_of_the_first_stream_to_end": {4: [2499]}, "position_zero": {5: [0], 0: [0, 0]},
               "many": {c: sorted(int(x) for x in rng.integers(0, 2400, 40)) for c in range(ped.n_person)}}[case]
    d = tmp_path / case
    d.mkdir()
    lines, gif = [], []
    col = 0
    for f in range(ped.n_fam):
        for j in range(int(ped.fam_size[f])):
            fa, mo = int(ped.father[col]), int(ped.mother[col])
            lines.append(f"fam{f + 1}\tp{j + 1}\t{('p%d' % (fa + 1)) if fa >= 0 else 0}\t{('p%d' % (mo + 1)) if mo >= 0 else 0}\t{int(ped.sex[col])}\t{col + 1}\n")
            rr = recs[:, col].copy()
            if col in ends:
                rr[ends[col]:] = np.zeros((), dtype=rr.dtype)
            path = str(d / f"g{col + 1}.glf")
            glfio.write_glf(path, "3", int(hdr["pos"].max()) + 1, hdr["pos"].astype(np.int64), hdr["ref_base"], rr)
            if col in repeats:
                _insert_repeats(path, "3", repeats[col], rng)
            gif.append(f"{col + 1} {path}\n")
            col += 1
    (d / "ped").write_text("".join(lines)); (d / "dat").write_text("T\tGLF_Index\n"); (d / "gif").write_text("".join(gif))
    outs = []
    for extra in ([], ["--batched", "1"], ["--batched", "4"], ["--batched", "3", "--portable", "1"]):
        out = str(d / ("out%d.pmpk" % len(outs)))
        subprocess.run([U.PM_TOOLS, "pack", "-p", str(d / "ped"), "-d", str(d / "dat"), "-g", str(d / "gif"), "-o", out] + extra, check=True)
        outs.append(load_pmpk(out))
    ref = outs[0]
    assert len(ref.hdr) > 2400 and int(np.sum(np.diff(ref.hdr["pos"].astype(np.int64)) == 0)) >= 1, "the one-site reader saw no repeated position"
    for o in outs[1:]:
        assert len(o.hdr) == len(ref.hdr), (len(o.hdr), len(ref.hdr))
        assert np.array_equal(o.hdr, ref.hdr) and np.array_equal(o.recs, ref.recs)


def test_synthetic_generator_is_seeded_and_well_formed():
    ped = synth.concat(synth.trios(3), synth.families([4, 1]))
    h1, r1 = synth.generate_sites(ped, 500, 11, cfg=synth.SynthConfig(poly_boost=30))
    h2, r2 = synth.generate_sites(ped, 500, 11, cfg=synth.SynthConfig(poly_boost=30))
    assert (h1 == h2).all() and (r1 == r2).all()
    recs = r1.numpy().view(capi.PERSON_SITE_DTYPE).reshape(500, ped.n_person)
    depth = recs["depth"][..., 0].astype(int) + 256 * recs["depth"][..., 1].astype(int)
    has = depth > 0
    assert (recs["lk"].min(axis=2)[has] == 0).all()          # min-normalised
    assert (recs["lk"][~has] == 0).all() and (recs["map_quality"][~has] == 0).all()
    assert 0.005 < (~has).mean() < 0.06


def test_row_formatter_prints_what_printf_prints():
    """The VCF row formatter's %.Nf / %d restatements (vcf_writer.cpp) agree with snprintf on random values,
    exact decimal ties and their neighbours, negative zeros and non-finite values."""
    import json
    import subprocess
    from cli_util import PM_TOOLS
    out = subprocess.run([PM_TOOLS, "fmt-selftest", "300000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rep = json.loads(out.stdout)
    assert rep["checked"] > 900000 and rep["fixed_mismatches"] == 0 and rep["int_mismatches"] == 0


def test_wire_forms_of_the_records():
    """capi.to_wire / capi.to_pl3: the byte layouts pm_call_glf_sites_wire and pm_call_vcf_records_pl take
    (include/polymutt_b200.h: pm_person_site_wire; three PL bytes per sample in the order a1a1, a1a2, a2a2)."""
    from polymutt_b200 import capi
    rng = np.random.default_rng(7)
    n, npers = 5, 4
    raw = rng.integers(0, 256, size=(n, npers, 16), dtype=np.uint8)
    recs = raw.view(capi.PERSON_SITE_DTYPE).reshape(n, npers)
    wire = capi.to_wire(recs)
    assert wire.shape == (n * npers, 14) and wire.dtype == np.uint8
    assert np.array_equal(wire.reshape(n, npers, 14), raw[:, :, :14])           # lk[10], depth[3], mapQ; the two pad bytes dropped
    hdr = np.zeros(n, dtype=capi.SITE_HDR_DTYPE)
    hdr["ref_base"] = [1, 2, 3, 4, 1]
    hdr["reserved"] = [2, 4, 1, 3, 4 | 0x100]                                    # ALT allele (| indel flag)
    pl3 = capi.to_pl3(hdr, recs, npers)
    assert pl3.shape == (n, npers, 3)
    for r in range(n):
        a1, a2 = int(hdr["ref_base"][r]), int(hdr["reserved"][r]) & 0xFF
        for k, (x, y) in enumerate(((a1, a1), (a1, a2), (a2, a2))):
            assert np.array_equal(pl3[r, :, k], raw[r, :, capi.geno_index(x, y)])
    # glfHandler.h:102-106: AA AC AG AT CC CG CT GG GT TT
    assert [capi.geno_index(a, b) for a in range(1, 5) for b in range(a, 5)] == list(range(10))
    assert capi.geno_index(3, 1) == capi.geno_index(1, 3)
