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


def body(path):
    with open(path, "rb") as f:
        return b"".join(l for l in f if not l.startswith(b"##"))


def main():
    os.makedirs(os.path.join(HERE, "peds"), exist_ok=True)
    for name in ("test.ped", "test.mix.ped", "test.dat"):
        shutil.copy(os.path.join(REF, name), os.path.join(HERE, "peds", name))
    open(os.path.join(HERE, "peds", "ext.ped"), "w").write(EXT_PED)
    open(os.path.join(HERE, "peds", "ceph.ped"), "w").write(CEPH_PED)
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


if __name__ == "__main__":
    main()
