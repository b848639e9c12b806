"""Helpers that drive the front-end executables on fixture data."""
import gzip
import hashlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PM_TOOLS = os.path.join(ROOT, "polymutt_b200", "bin", "pm-tools")
ORACLE_CLI = os.path.join(ROOT, "oracle", "_build", "polymutt_oracle_cli")
PRODUCT_CLI = os.path.join(ROOT, "polymutt_b200", "bin", "polymutt-b200")


def unpack_example(tmpdir):
    """Regenerates the 12 example GLF streams (one per column of example/test.ped) from the packed
    fixture and writes a GLF index keyed like example/test.gif (1..12)."""
    raw = os.path.join(tmpdir, "example12.pmpk")
    with gzip.open(os.path.join(GOLDEN, "example12.pmpk.gz"), "rb") as src, open(raw, "wb") as dst:
        dst.write(src.read())
    out = os.path.join(tmpdir, "glf")
    subprocess.run([PM_TOOLS, "unpack", raw, out], check=True, stderr=subprocess.DEVNULL)
    os.remove(raw)
    return out  # holds ped, dat, gif (keys 1..12 = test.ped's GLF_Index values)


def body(text: bytes) -> bytes:
    return b"".join(l for l in text.splitlines(keepends=True) if not l.startswith(b"##"))


def run_cli(exe, glfdir, ped, extra, out_path, timeout=1800):
    cmd = [exe, "-p", ped, "-d", os.path.join(glfdir, "dat"), "-g", os.path.join(glfdir, "gif"), "--out_vcf", out_path] + list(extra)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    log = p.stdout.decode(errors="replace")
    assert p.returncode == 0, log[-2000:]
    with open(out_path, "rb") as f:
        return body(f.read()), log


def golden_text(name):
    with gzip.open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def golden_sha(name):
    sha, lines = open(os.path.join(GOLDEN, name)).read().split()
    return sha, int(lines)


def sha_of(text: bytes):
    return hashlib.sha256(text).hexdigest(), text.count(b"\n")


def first_diff(a: bytes, b: bytes):
    la, lb = a.splitlines(), b.splitlines()
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return f"line {i}: got {x[:300]!r}\n      expected {y[:300]!r}"
    return f"line counts differ: got {len(la)} expected {len(lb)}"


# (case name, pedigree file, extra args, golden file) — the reference's shipped goldens (example/run.sh
# commands 1, 3, 4) and outputs of the unmodified reference on cases the shipped goldens do not cover.
CASES = [
    ("cmd1", "test.ped", ["-c", "0.9", "--minDepth", "150", "--maxDepth", "200", "--nthreads", "4"], "golden_cmd1.vcf.gz"),
    ("cmd3", "test.mix.ped", [], "golden_cmd3.vcf.gz"),
    ("cmd4", "test.ped", ["--nthreads", "4", "--denovo", "--rate_denovo", "1.5e-07"], "golden_cmd4.vcf.gz"),
    ("ext_ba", "ext.ped", [], "ref_ext_ba.vcf.gz"),
    ("ceph_ba", "ceph.ped", [], "ref_ceph_ba.vcf.gz"),
    ("mix_strict", "test.mix.ped", ["-c", "0.99", "--minMapQuality", "50", "--minPercSampleWithData", "90", "--theta", "0.01", "--poly_tstv", "3.0"], "ref_mix_strict.vcf.gz"),
    ("mix_denovo_loose", "test.mix.ped", ["--denovo", "--rate_denovo", "1e-4", "--minLLR_denovo", "1e-3", "--tstv_denovo", "1.0"], "ref_mix_denovo_loose.vcf.gz"),
    ("mix_all_sites", "test.mix.ped", ["--all_sites"], "ref_mix_all_sites.sha"),
    ("ext_denovo", "ext.ped", ["--denovo"], "ref_ext_denovo.vcf.gz"),
    ("ceph_denovo", "ceph.ped", ["--denovo"], "ref_ceph_denovo.sha"),
]
# chrX / chrY / MT (the example's section "1" declared to be that chromosome): nuclear families, singletons, a lone
# nuclear family (fixed parent table) and an extended pedigree, bi-allelic and --denovo.
NONAUTO_CASES = []
for _c, _flag in (("x", "--chrX"), ("y", "--chrY"), ("mt", "--MT")):
    for _p, _ped in (("quartets", "test.ped"), ("mix", "test.mix.ped"), ("single", "single.ped"), ("ext", "ext.ped")):
        _full = (_c, _p) in (("x", "quartets"), ("y", "mix"), ("mt", "ext"))
        NONAUTO_CASES.append((f"{_c}_{_p}_ba", _ped, [_flag, "1"], f"ref_{_c}_{_p}_ba" + (".vcf.gz" if _full else ".sha")))
        NONAUTO_CASES.append((f"{_c}_{_p}_dn", _ped, [_flag, "1", "--denovo", "--rate_denovo", "1.5e-07"],
                              f"ref_{_c}_{_p}_dn" + (".vcf.gz" if (_c, _p) == ("y", "ext") else ".sha")))
# an extended family next to more than eight nuclear / single units: one block per site AND a peel (the ES instances of
# the wide kernel)
MIXEXT_CASES = [
    ("mixext_ba", "mixext.ped", [], "ref_mixext_ba.sha"),
    ("mixext_dn", "mixext.ped", ["--denovo", "--rate_denovo", "1.5e-07"], "ref_mixext_dn.vcf.gz"),
    ("mixext_x", "mixext.ped", ["--chrX", "1"], "ref_mixext_x.vcf.gz"),
]
# --pos FILE (force-call the listed positions, stop after the last one; duplicate and foreign-chromosome lines in the
# list count towards the stop rule, src/main.cpp:332-337, 593)
POS_CASES = [
    ("pos_mix", "test.mix.ped", ["--pos", os.path.join(GOLDEN, "pos_300.txt")], "ref_pos_mix.sha"),
    ("pos_quartets_dn", "test.ped", ["--pos", os.path.join(GOLDEN, "pos_300.txt"), "--denovo", "--rate_denovo", "1.5e-07"], "ref_pos_quartets_dn.sha"),
    ("pos_ext_dup", "ext.ped", ["--pos", os.path.join(GOLDEN, "pos_dup.txt")], "ref_pos_ext_dup.sha"),
    ("pos_single_c099", "single.ped", ["--pos", os.path.join(GOLDEN, "pos_300.txt"), "-c", "0.99"], "ref_pos_single_c099.sha"),
]
# GLF files with three sections ("1", "2", "X": the example cut in three, tests/fixtures_util.py): --chr2process, the
# per-chromosome class switch, --gl_off, one summary block per processed chromosome
MS_CASES = [
    ("ms_all", "test.mix.ped", ["--chrX", "X"], "ref_ms_all.sha"),
    ("ms_chr2", "test.mix.ped", ["--chr2process", "2,X", "--chrX", "X"], "ref_ms_chr2.sha"),
    ("ms_gl_off", "test.ped", ["--chr2process", "1", "--gl_off"], "ref_ms_gl_off.sha"),
    ("ms_denovo", "test.ped", ["--denovo", "--rate_denovo", "1.5e-07", "--chrX", "X"], "ref_ms_denovo.sha"),
    ("ms_y_mt", "ext.ped", ["--chrY", "2", "--MT", "X"], "ref_ms_y_mt.sha"),
]
# GLF files in which some streams repeat a position (a base record with offset 0 after a base record,
# tests/fixtures_util.py:write_repeat_glfs): the reference's cursor makes one more site at that position out of the repeats
# alone (src/PedigreeGLF.cpp:282-324)
REP_CASES = [
    ("rep_mix", "test.mix.ped", [], "ref_rep_mix.vcf.gz"),
    ("rep_quartets_dn", "test.ped", ["--denovo", "--rate_denovo", "1.5e-07"], "ref_rep_quartets_dn.sha"),
    ("rep_ext_all", "ext.ped", ["--pos", os.path.join(GOLDEN, "pos_rep.txt")], "ref_rep_ext_all.vcf.gz"),
]
# --quick_call: outputs of the unmodified reference with the everybody-unrelated pre-pass switched on
QUICK_CASES = [
    ("q_quartets", "test.ped", ["--quick_call"], "ref_q_quartets.sha"),
    ("q_mix", "test.mix.ped", ["--quick_call"], "ref_q_mix.sha"),
    ("q_single", "single.ped", ["--quick_call"], "ref_q_single.sha"),
    ("q_ext", "ext.ped", ["--quick_call"], "ref_q_ext.sha"),
    ("q_quartets_dn", "test.ped", ["--quick_call", "--denovo", "--rate_denovo", "1.5e-07"], "ref_q_quartets_dn.vcf.gz"),
    ("q_mix_c099", "test.mix.ped", ["--quick_call", "-c", "0.99"], "ref_q_mix_c099.sha"),
    ("q_x_quartets", "test.ped", ["--quick_call", "--chrX", "1"], "ref_q_x_quartets.sha"),
]
# Rows whose allele frequency is rounding noise in the reference itself: the objective Brent minimises there does not
# depend on p at all (every likelihood that multiplies p equals the one that multiplies q, so L(p) = l (p + q)), its
# value wobbles in the last bit and the reference's optimiser ends wherever that noise sends it.  No other
# implementation of the same arithmetic can land on the same point; such rows are compared without AF and what is
# derived from it.  (case, POS) -> reason
FLAT_OBJECTIVE_ROWS = {("y_ext_dn", 71912): "chrY, every male's two homozygous likelihoods are equal: the mutation-free refit is flat in p"}
# Rows with a genotype posterior that ties exactly in exact arithmetic (a son on chrX with PL x,0,x under a heterozygous
# mother: P(0) = P(1) = 1/2, GQ 3): which side wins is the rounding noise of the allele frequency's last digits, and the
# wide kernel's Brent objective is not evaluated in the reference's order.  The sample columns of these rows are not
# compared; CHROM..INFO (incl. QUAL and AF to four decimals) still are.
TIED_GENOTYPE_ROWS = {("mixext_x", 35856): "famA member 5, PL 187,0,187", ("mixext_x", 75468): "famA member 5, PL 169,0,169"}
# VCF-input mode (--in_vcf): (case, pedigree, input VCF fixture, golden) — the shipped golden of run.sh command 2
# and outputs of the unmodified reference on edge-case inputs (tests/golden/make_golden.py: make_vcf_inputs)
VCF_CASES = [
    ("vcf_cmd2", "test.ped", "vcf_in_full.vcf.gz", "ref_vcf_cmd2.vcf.gz"),
    ("vcf_mix_edge", "test.mix.ped", "vcf_in_edge.vcf.gz", "ref_vcf_mix_edge.vcf.gz"),
    ("vcf_single_family_edge", "single.ped", "vcf_in_edge.vcf.gz", "ref_vcf_single_family_edge.vcf.gz"),
    ("vcf_ext_edge", "ext.ped", "vcf_in_edge.vcf.gz", "ref_vcf_ext_edge.vcf.gz"),
    ("vcf_quartets_gl", "test.ped", "vcf_in_gl.vcf.gz", "ref_vcf_quartets_gl.vcf.gz"),
]
# the same inputs declared to be chrX / chrY / MT (--chrX 1 ...): every family with non-founders is peeled there and the
# labels turn haploid / "." (5th field: extra arguments)
VCF_NONAUTO_CASES = [
    ("vcf_x_quartets", "test.ped", "vcf_in_full.vcf.gz", "ref_vcf_x_quartets.vcf.gz", ["--chrX", "1"]),
    ("vcf_y_mix", "test.mix.ped", "vcf_in_full.vcf.gz", "ref_vcf_y_mix.vcf.gz", ["--chrY", "1"]),
    ("vcf_mt_quartets", "test.ped", "vcf_in_full.vcf.gz", "ref_vcf_mt_quartets.vcf.gz", ["--MT", "1"]),
    ("vcf_x_ext_edge", "ext.ped", "vcf_in_edge.vcf.gz", "ref_vcf_x_ext_edge.vcf.gz", ["--chrX", "1"]),
    ("vcf_y_single_edge", "single.ped", "vcf_in_edge.vcf.gz", "ref_vcf_y_single_edge.vcf.gz", ["--chrY", "1"]),
    ("vcf_mt_mix_edge", "test.mix.ped", "vcf_in_edge.vcf.gz", "ref_vcf_mt_mix_edge.vcf.gz", ["--MT", "1"]),
    ("vcf_y_ext_edge", "ext.ped", "vcf_in_edge.vcf.gz", "ref_vcf_y_ext_edge.vcf.gz", ["--chrY", "1"]),
]


def check_vcf_case(exe, tmpdir, case, gz_input=False, extra=()):
    name, ped, vin, golden = case[:4]
    extra = list(extra) + (list(case[4]) if len(case) > 4 else [])
    src = os.path.join(GOLDEN, vin)
    if gz_input:
        inp = src                                   # the reader is gz-transparent
    else:
        inp = os.path.join(tmpdir, vin[:-3])
        with gzip.open(src, "rb") as f, open(inp, "wb") as g:
            g.write(f.read())
    out = os.path.join(tmpdir, name + ".out.vcf")
    cmd = [exe, "-p", os.path.join(GOLDEN, "peds", ped), "-d", os.path.join(GOLDEN, "peds", "test.dat"), "--in_vcf", inp, "--out_vcf", out] + list(extra)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1800)
    log = p.stdout.decode(errors="replace")
    assert p.returncode == 0, log[-2000:]
    got = body(open(out, "rb").read())
    want = golden_text(golden)
    assert got == want, f"{name}: " + first_diff(got, want)
    return log


SLOW_FOR_ORACLE = {"ext_denovo", "ceph_denovo", "x_ext_dn", "y_ext_dn", "mt_ext_dn"}  # minutes of CPU in the oracle; covered on the GPU and by make_golden runs


def check_case(exe, glfdir, tmpdir, case):
    name, ped, extra, golden = case
    got, log = run_cli(exe, glfdir, os.path.join(GOLDEN, "peds", ped), extra, os.path.join(tmpdir, name + ".vcf"))
    flat = {pos for (c, pos) in FLAT_OBJECTIVE_ROWS if c == name}
    tied = {pos for (c, pos) in TIED_GENOTYPE_ROWS if c == name}
    if (flat or tied) and not golden.endswith(".sha"):
        def strip(text):
            out = []
            for l in text.splitlines(keepends=True):
                t = l.split(b"\t", 9)
                if len(t) > 9 and not l.startswith(b"#") and int(t[1]) in flat:
                    l = b"\t".join(t[:7]) + b"\t<flat objective: AF, GQ not compared>\n"   # CHROM..FILTER still have to agree
                elif len(t) > 9 and not l.startswith(b"#") and int(t[1]) in tied:
                    l = b"\t".join(t[:9]) + b"\t<tied genotype posterior: sample columns not compared>\n"
                out.append(l)
            return b"".join(out)
        want = strip(golden_text(golden))
        got = strip(got)
        assert got == want, f"{name}: " + first_diff(got, want)
        return log
    if golden.endswith(".sha"):
        if sha_of(got) != golden_sha(golden):
            dump = os.path.join(ROOT, "gpurun_out")   # keep the text for a post-mortem against the reference's
            if os.path.isdir(dump):
                with gzip.open(os.path.join(dump, f"fail_{name}.vcf.gz"), "wb") as f:
                    f.write(got)
        assert sha_of(got) == golden_sha(golden), f"{name}: sha/line-count mismatch {sha_of(got)} vs {golden_sha(golden)}"
    else:
        want = golden_text(golden)
        assert got == want, f"{name}: " + first_diff(got, want)
    return log


# BASELINE.json configs 2, 4, 5 in miniature (tests/baseline_shapes.py): seeded synthetic inputs, outputs of the unmodified
# reference committed as sha256 + line count (+ the sha256 of the generated input)
BASELINE_SHAPES = ["cfg2_trios1000_dn", "cfg5_mixed100_ba", "cfg5_mixed100_dn", "cfg4_vcf200x5"]


def check_baseline_shape(exe, tmpdir, name, extra=()):
    import resource
    import baseline_shapes as B
    soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
    if soft < 4096:
        resource.setrlimit(resource.RLIMIT_NOFILE, (min(4096, hard), hard))
    sha, lines, input_sha = open(os.path.join(GOLDEN, f"ref_{name}.sha")).read().split()
    argv, digest = B.materialise(name, tmpdir)
    assert digest == input_sha, f"{name}: the synthetic generator no longer produces the input the reference output was made from"
    out = os.path.join(tmpdir, name + ".out.vcf")
    p = subprocess.run([exe] + argv + ["--out_vcf", out] + list(extra), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1800)
    assert p.returncode == 0, p.stdout.decode(errors="replace")[-2000:]
    got = body(open(out, "rb").read())
    if sha_of(got) != (sha, int(lines)):
        head = golden_text(f"ref_{name}.head.vcf.gz").splitlines()
        mine = [b"\t".join(l.split(b"\t")[:14]) for l in got.splitlines()[:41]]
        diff = next((f"row {i}: got {a[:200]!r} expected {b[:200]!r}" for i, (a, b) in enumerate(zip(mine, head)) if a != b), "first 40 rows agree")
        raise AssertionError(f"{name}: sha/line count {sha_of(got)} vs reference {(sha, int(lines))}; {diff}")
