#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference tree (run in the build container, where
/root/reference exists and oracle/build_ref.sh has produced oracle/_ref/polymutt).

  example12.pmpk.gz          the 12 example GLFs (81,016 sites) merged as the engine receives them for
                             12 unrelated people (column i = GLF i), packed by `pm-tools pack`
  golden_cmd1.vcf.gz ...     the reference's own shipped goldens for example/run.sh commands 1, 3, 4
                             (non-## lines only)
  ref_<case>.vcf.gz / .sha   output of the UNMODIFIED reference binary on pedigrees the shipped
                             goldens do not cover (extended pedigrees, --denovo on them, --all_sites),
                             same GLFs; big outputs are kept as sha256 + line count only
  peds/*.ped                 the pedigrees (ext.ped and ceph.ped are ours; GLF_Index re-uses the 12 files)
"""
import gzip
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/example"
REFBIN = os.path.join(ROOT, "oracle", "_ref", "polymutt")
TOOLS = os.path.join(ROOT, "polymutt_b200", "bin", "pm-tools")

EXT_PED = """famA	1	0	0	1	1
famA	2	0	0	2	2
famA	3	1	2	1	3
famA	4	0	0	2	4
famA	5	3	4	1	5
famA	6	3	4	2	6
famA	7	1	2	2	7
famB	9	0	0	1	9
famB	10	0	0	2	10
famB	11	9	10	1	11
famB	12	9	10	1	12
"""
CEPH_PED = "".join(
    ["ceph\tg1\t0\t0\t1\t1\nceph\tg2\t0\t0\t2\t2\nceph\tg3\t0\t0\t1\t3\nceph\tg4\t0\t0\t2\t4\n",
     "ceph\tp1\tg1\tg2\t1\t5\nceph\tp2\tg3\tg4\t2\t6\n"] +
    [f"ceph\tk{k}\tp1\tp2\t{1 + (k + 1) % 2}\t{7 + (k - 1) % 6 if k <= 6 else (k - 6)}\n" for k in range(1, 15)])

CASES = {  # name: (ped file, extra args, keep full text?)
    "ext_ba": ("ext.ped", [], True),
    "ext_denovo": ("ext.ped", ["--denovo"], True),
    "ceph_ba": ("ceph.ped", [], True),
    "ceph_denovo": ("ceph.ped", ["--denovo"], False),
    "mix_all_sites": ("test.mix.ped", ["--all_sites"], False),
    "mix_denovo_loose": ("test.mix.ped", ["--denovo", "--rate_denovo", "1e-4", "--minLLR_denovo", "1e-3", "--tstv_denovo", "1.0"], True),
    "mix_strict": ("test.mix.ped", ["-c", "0.99", "--minMapQuality", "50", "--minPercSampleWithData", "90", "--theta", "0.01", "--poly_tstv", "3.0"], True),
}
# an extended family next to more than eight nuclear / single units (the wide kernel's ES instances); GLFs re-used
MIXEXT_PED = EXT_PED + "".join(f"s{i}\ts{i}\t0\t0\t{1 + i % 2}\t{1 + (i * 5) % 12}\n" for i in range(1, 11))
CASES.update({
    "mixext_ba": ("mixext.ped", [], False),
    "mixext_dn": ("mixext.ped", ["--denovo", "--rate_denovo", "1.5e-07"], True),
    "mixext_x": ("mixext.ped", ["--chrX", "1"], True),   # full text: see TIED_GENOTYPE_ROWS in cli_util.py
})
# --pos (pos_300.txt: 300 random positions of section "1"; pos_dup.txt: 50 of them, one twice, one on another chromosome)
CASES.update({
    "pos_mix": ("test.mix.ped", ["--pos", os.path.join(HERE, "pos_300.txt")], False),
    "pos_quartets_dn": ("test.ped", ["--pos", os.path.join(HERE, "pos_300.txt"), "--denovo", "--rate_denovo", "1.5e-07"], False),
    "pos_ext_dup": ("ext.ped", ["--pos", os.path.join(HERE, "pos_dup.txt")], False),
    "pos_single_c099": ("single.ped", ["--pos", os.path.join(HERE, "pos_300.txt"), "-c", "0.99"], False),
})
# --quick_call (the everybody-unrelated pre-pass of main.cpp:354-437)
CASES.update({
    "q_quartets": ("test.ped", ["--quick_call"], False),
    "q_mix": ("test.mix.ped", ["--quick_call"], False),
    "q_single": ("single.ped", ["--quick_call"], False),
    "q_ext": ("ext.ped", ["--quick_call"], False),
    "q_quartets_dn": ("test.ped", ["--quick_call", "--denovo", "--rate_denovo", "1.5e-07"], True),
    "q_mix_c099": ("test.mix.ped", ["--quick_call", "-c", "0.99"], False),
    "q_x_quartets": ("test.ped", ["--quick_call", "--chrX", "1"], False),
})
# chrX / chrY / MT rules: the example's only section is labelled "1", so `--chrX 1` (--chrY 1, --MT 1) makes the
# reference treat the same data as that chromosome.  Three cases keep their text, the rest sha256 + line count.
DN = ["--denovo", "--rate_denovo", "1.5e-07"]
for _c, _flag in (("x", "--chrX"), ("y", "--chrY"), ("mt", "--MT")):
    for _p, _ped in (("quartets", "test.ped"), ("mix", "test.mix.ped"), ("single", "single.ped"), ("ext", "ext.ped")):
        CASES[f"{_c}_{_p}_ba"] = (_ped, [_flag, "1"], (_c, _p) in (("x", "quartets"), ("y", "mix"), ("mt", "ext")))
        CASES[f"{_c}_{_p}_dn"] = (_ped, [_flag, "1"] + DN, (_c, _p) == ("y", "ext"))  # y_ext_dn: see FLAT_OBJECTIVE_ROWS in cli_util.py


VCF_CASES = {  # name: (ped file, input vcf)
    "vcf_cmd2": ("test.ped", "vcf_in_full.vcf.gz"),
    "vcf_mix_edge": ("test.mix.ped", "vcf_in_edge.vcf.gz"),
    "vcf_single_family_edge": ("single.ped", "vcf_in_edge.vcf.gz"),
    "vcf_ext_edge": ("ext.ped", "vcf_in_edge.vcf.gz"),
    "vcf_quartets_gl": ("test.ped", "vcf_in_gl.vcf.gz"),
}
# chrX / chrY / MT under --in_vcf: name -> (ped file, input vcf, extra args)
VCF_NONAUTO_CASES = {
    "vcf_x_quartets": ("test.ped", "vcf_in_full.vcf.gz", ["--chrX", "1"]),
    "vcf_y_mix": ("test.mix.ped", "vcf_in_full.vcf.gz", ["--chrY", "1"]),
    "vcf_mt_quartets": ("test.ped", "vcf_in_full.vcf.gz", ["--MT", "1"]),
    "vcf_x_ext_edge": ("ext.ped", "vcf_in_edge.vcf.gz", ["--chrX", "1"]),
    "vcf_y_single_edge": ("single.ped", "vcf_in_edge.vcf.gz", ["--chrY", "1"]),
    "vcf_mt_mix_edge": ("test.mix.ped", "vcf_in_edge.vcf.gz", ["--MT", "1"]),
    "vcf_y_ext_edge": ("ext.ped", "vcf_in_edge.vcf.gz", ["--chrY", "1"]),
}
SINGLE_PED = "fam1\t1\t0\t0\t1\t1\nfam1\t2\t0\t0\t2\t2\nfam1\t3\t1\t2\t2\t3\nfam1\t4\t1\t2\t1\t4\n"


def make_vcf_inputs():
    """Derived VCF inputs: the shipped example input (command 2 of run.sh) as is, an edge-case variant of its first
    1,600 records (multi-allelic, REF==ALT, indels, missing / empty / all-zero PL fields, missing DP, a sample that is
    not in the pedigree) and a GL-typed variant."""
    src = open(os.path.join(REF, "testvcf.in.vcf")).read().split("\n")
    with gzip.GzipFile(os.path.join(HERE, "vcf_in_full.vcf.gz"), "wb", 9, mtime=0) as f:
        f.write("\n".join(src).encode())
    head = [l for l in src if l.startswith("#")]
    recs = [l for l in src if l and not l.startswith("#")][:1600]
    edge, gl = list(head[:-1]), list(head[:-1])
    cols = head[-1].split("\t")
    edge.append("\t".join(cols + ["stranger"]))
    gl.append(head[-1])
    for n, line in enumerate(recs):
        t = line.split("\t")
        g = list(t)
        g[8] = "GT:GQ:DP:DS:GL"
        for i in range(9, len(t)):
            f = t[i].split(":")
            pl = [int(x) for x in f[4].split(",")]
            f[4] = ",".join("%.2f" % (-x / 10.0) if x else "0" for x in pl)
            if n % 17 == 5 and i == 12:
                f[4] = "-30.1,-0.004,-12"          # beyond the 255 cap, fractional phred
            g[i] = ":".join(f)
        gl.append("\t".join(g))
        e = list(t) + ["0/0:10:9:0.00:0,20,200"]
        if n % 7 == 3:
            e[4] = "T,G"                            # not bi-allelic: dropped
        elif n % 11 == 4:
            e[4] = e[3]                             # REF == ALT: dropped
        elif n % 13 == 6:
            e[3], e[4] = "CA", "C"                  # indel
        elif n % 19 == 8:
            e[3], e[4] = "G", "A"                   # a transition the reference's isTs() does not recognise
        if n % 23 == 9:
            e[12] = ":".join(e[12].split(":")[:3])  # PL absent for the 4th sample: the samples after it are ignored
        if n % 29 == 10:
            f = e[10].split(":"); f[4] = ""; e[10] = ":".join(f)   # empty PL
        if n % 31 == 11:
            for i in range(9, len(e)):
                f = e[i].split(":")
                if len(f) > 4:
                    f[4] = "0,0,0"                  # nobody has data: stale output
                e[i] = ":".join(f)
        if n % 37 == 12:
            f = e[15].split(":"); f[2] = ""; e[15] = ":".join(f)   # missing DP
        if n % 41 == 13:
            f = e[9].split(":"); f[4] = "300,0,999"; e[9] = ":".join(f)   # beyond the cap
        edge.append("\t".join(e))
    for name, lines in (("vcf_in_edge.vcf.gz", edge), ("vcf_in_gl.vcf.gz", gl)):
        with gzip.GzipFile(os.path.join(HERE, name), "wb", 9, mtime=0) as f:
            f.write(("\n".join(lines) + "\n").encode())


def body(path):
    with open(path, "rb") as f:
        return b"".join(l for l in f if not l.startswith(b"##"))


def make_baseline_shapes(only):
    """BASELINE.json configs 2, 4, 5 in miniature (tests/baseline_shapes.py): the unmodified reference on seeded synthetic
    inputs of those shapes.  ref_<name>.sha = sha256 and line count of the non-## output + sha256 of the generated input;
    the first 40 data rows, cut to their first 14 columns, are kept as text for post-mortems."""
    import resource
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import baseline_shapes as B
    soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
    resource.setrlimit(resource.RLIMIT_NOFILE, (min(max(4096, soft), hard), hard))
    for name in B.SHAPES:
        if only and name not in only and "cfg" not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            argv, digest = B.materialise(name, tmp)
            out = os.path.join(tmp, name + ".vcf")
            subprocess.run([REFBIN] + argv + ["--nthreads", "1", "--out_vcf", out], check=True, stdout=subprocess.DEVNULL)
            text = body(out)
        with open(os.path.join(HERE, f"ref_{name}.sha"), "w") as f:
            f.write("%s %d %s\n" % (hashlib.sha256(text).hexdigest(), text.count(b"\n"), digest))
        head = b"".join(b"\t".join(l.split(b"\t")[:14]) + b"\n" for l in text.splitlines()[:41])
        with gzip.GzipFile(os.path.join(HERE, f"ref_{name}.head.vcf.gz"), "wb", 9, mtime=0) as dst:
            dst.write(head)
        print(name, len(text), "bytes", text.count(b"\n"), "lines")


def make_glf_variants(only):
    """Goldens on GLF inputs derived from the example (tests/fixtures_util.py): three sections per file (ms_*) and
    repeated positions (rep_*).  pos_rep.txt lists 40 of the repeated positions (and 20 others), so that the rows of
    those sites -- both of them where a position has two -- are printed whatever their quality."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import cli_util as U
    import fixtures_util as F
    from polymutt_b200 import load_pmpk
    ex = load_pmpk(os.path.join(HERE, "example12.pmpk.gz"))
    rep = F.repeat_sites(len(ex.hdr), ex.recs.shape[1])
    rng = np.random.default_rng(7)
    sites = sorted(set(int(x) for g in sorted(rep) for x in rep[g][::4]) | set(int(x) for x in rng.integers(0, len(ex.hdr), 20)))
    open(os.path.join(HERE, "pos_rep.txt"), "w").write("".join("%s\t%d\n" % (ex.label, int(ex.hdr["pos"][i]) + 1) for i in sites))
    with tempfile.TemporaryDirectory() as tmp:
        for kind, writer, cases in (("ms", F.write_multisection_glfs, U.MS_CASES), ("rep", F.write_repeat_glfs, U.REP_CASES)):
            if only and kind not in only and not any(c[0] in only for c in cases):
                continue
            d = os.path.join(tmp, kind)
            writer(ex, d)
            for name, ped, extra, golden in cases:
                if only and kind not in only and name not in only:
                    continue
                out = os.path.join(tmp, name + ".vcf")
                subprocess.run([REFBIN, "-p", os.path.join(HERE, "peds", ped), "-d", os.path.join(d, "dat"), "-g", os.path.join(d, "gif"),
                                "--out_vcf", out] + list(extra), check=True, stdout=subprocess.DEVNULL)
                text = body(out)
                if golden.endswith(".vcf.gz"):
                    with gzip.GzipFile(os.path.join(HERE, golden), "wb", 9, mtime=0) as dst:
                        dst.write(text)
                else:
                    with open(os.path.join(HERE, golden), "w") as f:
                        f.write("%s %d\n" % (hashlib.sha256(text).hexdigest(), text.count(b"\n")))
                print(name, len(text), "bytes", text.count(b"\n"), "lines")


def main():
    if sys.argv[1:] and all(a.startswith("cfg") for a in sys.argv[1:]):
        return make_baseline_shapes(set(sys.argv[1:]))
    if sys.argv[1:] and all(a.startswith("ms") or a.startswith("rep") for a in sys.argv[1:]):
        return make_glf_variants(set(sys.argv[1:]))
    os.makedirs(os.path.join(HERE, "peds"), exist_ok=True)
    for name in ("test.ped", "test.mix.ped", "test.dat"):
        shutil.copy(os.path.join(REF, name), os.path.join(HERE, "peds", name))
    open(os.path.join(HERE, "peds", "ext.ped"), "w").write(EXT_PED)
    open(os.path.join(HERE, "peds", "ceph.ped"), "w").write(CEPH_PED)
    open(os.path.join(HERE, "peds", "mixext.ped"), "w").write(MIXEXT_PED)
    with tempfile.TemporaryDirectory() as tmp:
        for f in os.listdir(REF):
            if f.endswith(".glf") or f in ("test.gif", "test.dat"):
                shutil.copy(os.path.join(REF, f), tmp)
        for f in os.listdir(os.path.join(HERE, "peds")):
            shutil.copy(os.path.join(HERE, "peds", f), tmp)
        # 12 unrelated people, person i <-> GLF i: the fixture holds every stream, in GLF-index order
        open(os.path.join(tmp, "all12.ped"), "w").write("".join(f"u{i}\t{i}\t0\t0\t1\t{i}\n" for i in range(1, 13)))
        subprocess.run([TOOLS, "pack", "-p", "all12.ped", "-d", "test.dat", "-g", "test.gif", "-o", "example12.pmpk"], cwd=tmp, check=True)
        with open(os.path.join(tmp, "example12.pmpk"), "rb") as src, gzip.GzipFile(os.path.join(HERE, "example12.pmpk.gz"), "wb", 9, mtime=0) as dst:
            shutil.copyfileobj(src, dst)
        for cmd, golden in (("cmd1", "test.out.vcf"), ("cmd3", "test.out.vcfa"), ("cmd4", "test.denovo.out.vcf")):
            with gzip.GzipFile(os.path.join(HERE, f"golden_{cmd}.vcf.gz"), "wb", 9, mtime=0) as dst:
                dst.write(body(os.path.join(REF, golden)))
        only = set(sys.argv[1:])
        if not only or any(o.startswith("vcf") for o in only):
            open(os.path.join(HERE, "peds", "single.ped"), "w").write(SINGLE_PED)
            shutil.copy(os.path.join(HERE, "peds", "single.ped"), tmp)
            make_vcf_inputs()
            for name, spec in list(VCF_CASES.items()) + list(VCF_NONAUTO_CASES.items()):
                ped, vin = spec[0], spec[1]
                more = spec[2] if len(spec) > 2 else []
                if only and name not in only and "vcf" not in only:
                    continue
                raw = os.path.join(tmp, vin[:-3])
                with gzip.open(os.path.join(HERE, vin), "rb") as src, open(raw, "wb") as dst:
                    dst.write(src.read())
                out = os.path.join(tmp, name + ".vcf")
                subprocess.run([REFBIN, "-p", ped, "-d", "test.dat", "--in_vcf", raw, "--out_vcf", out] + more, cwd=tmp, check=True, stdout=subprocess.DEVNULL)
                text = body(out)
                with gzip.GzipFile(os.path.join(HERE, f"ref_{name}.vcf.gz"), "wb", 9, mtime=0) as dst:
                    dst.write(text)
                print(name, len(text), "bytes", text.count(b"\n"), "lines")
        for name, (ped, extra, keep) in CASES.items():
            if only and name not in only:
                continue
            out = os.path.join(tmp, name + ".vcf")
            subprocess.run([REFBIN, "-p", ped, "-d", "test.dat", "-g", "test.gif", "--out_vcf", out] + extra, cwd=tmp, check=True, stdout=subprocess.DEVNULL)
            text = body(out)
            if keep:
                with gzip.GzipFile(os.path.join(HERE, f"ref_{name}.vcf.gz"), "wb", 9, mtime=0) as dst:
                    dst.write(text)
            with open(os.path.join(HERE, f"ref_{name}.sha"), "w") as f:
                f.write("%s %d\n" % (hashlib.sha256(text).hexdigest(), text.count(b"\n")))
            print(name, len(text), "bytes", text.count(b"\n"), "lines")
        if not only:
            make_baseline_shapes(set())
            make_glf_variants(set())


if __name__ == "__main__":
    main()
