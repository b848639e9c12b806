"""End to end on the GPU box: the drop-in executable (CUDA engine) must print the reference's VCFs.

Same goldens as tests/test_oracle_golden.py: the reference's shipped example outputs and outputs of the
unmodified reference on extended pedigrees / --denovo / --all_sites.  Byte-for-byte on non-## lines;
the formats print 2-4 significant decimals of doubles that agree to ~1e-12, so a difference means a
real bug or a knife-edge rounding, which is reported with the offending line."""
import os
import subprocess

import pytest

import cli_util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def glfdir(tools_built, tmp_path_factory):
    subprocess.run(["make", "-s", "cli"], cwd=U.ROOT, check=True)
    return U.unpack_example(str(tmp_path_factory.mktemp("example")))


@pytest.mark.parametrize("case", U.CASES, ids=lambda c: c[0])
def test_product_cli_matches_reference_golden(case, glfdir, tmp_path):
    if not os.path.exists(os.path.join(U.GOLDEN, case[3])):
        pytest.skip("golden not generated")
    log = U.check_case(U.PRODUCT_CLI, glfdir, str(tmp_path), case)
    assert "Summary of reference -- 1" in log


@pytest.mark.parametrize("case", U.NONAUTO_CASES, ids=lambda c: c[0])
def test_product_cli_matches_reference_on_sex_chromosomes_and_mt(case, glfdir, tmp_path):
    """--chrX / --chrY / --MT: all 24 reference outputs (4 pedigree shapes x 3 chromosome classes x bi-allelic / --denovo)."""
    U.check_case(U.PRODUCT_CLI, glfdir, str(tmp_path), case)


@pytest.fixture(scope="module")
def msdir(example12, tools_built, tmp_path_factory):
    import fixtures_util as F
    d = str(tmp_path_factory.mktemp("multisection"))
    F.write_multisection_glfs(example12, d)
    return d


@pytest.mark.parametrize("case", U.MS_CASES, ids=lambda c: c[0])
def test_product_cli_multi_section_glfs(case, msdir, glfdir, tmp_path):
    U.check_case(U.PRODUCT_CLI, msdir, str(tmp_path), case)


@pytest.mark.parametrize("case", U.REP_CASES, ids=lambda c: c[0])
def test_product_cli_glfs_with_repeated_positions(case, example12, tools_built, tmp_path):
    """GLF streams that repeat a position (offset 0): one more site at that position, made of the repeats alone."""
    import fixtures_util as F
    d = str(tmp_path / "repeats")
    F.write_repeat_glfs(example12, d)
    U.check_case(U.PRODUCT_CLI, d, str(tmp_path), case)


@pytest.mark.parametrize("case", U.POS_CASES, ids=lambda c: c[0])
def test_product_cli_pos_list_matches_reference(case, glfdir, tmp_path):
    U.check_case(U.PRODUCT_CLI, glfdir, str(tmp_path), case)


@pytest.mark.parametrize("case", U.MIXEXT_CASES, ids=lambda c: c[0])
def test_product_cli_extended_family_among_many_units(case, glfdir, tmp_path):
    U.check_case(U.PRODUCT_CLI, glfdir, str(tmp_path), case)


@pytest.mark.parametrize("case", U.QUICK_CASES, ids=lambda c: c[0])
def test_product_cli_quick_call_matches_reference(case, glfdir, tmp_path):
    U.check_case(U.PRODUCT_CLI, glfdir, str(tmp_path), case)


@pytest.mark.parametrize("case", U.VCF_CASES, ids=lambda c: c[0])
def test_product_cli_vcf_input_matches_reference(case, glfdir, tmp_path):
    log = U.check_vcf_case(U.PRODUCT_CLI, str(tmp_path), case, gz_input=(case[0] == "vcf_cmd2"))
    assert "Total samples in both VCF and PED files" in log


@pytest.mark.parametrize("case", U.VCF_NONAUTO_CASES, ids=lambda c: c[0])
def test_product_cli_vcf_input_on_sex_chromosomes_and_mt(case, glfdir, tmp_path):
    U.check_vcf_case(U.PRODUCT_CLI, str(tmp_path), case)


def test_product_cli_small_batches_and_multi_gpu(glfdir, tmp_path):
    """Many small batches through the double-buffered host entry point; with more than one GPU on the box the batches
    are also sharded over all of them (--gpus N) and concatenated in order."""
    import torch
    n = max(1, torch.cuda.device_count())
    case = ("cmd3_batched", "test.mix.ped", ["--gpus", str(min(n, 8)), "--batch_sites", "3000"], "golden_cmd3.vcf.gz")
    U.check_case(U.PRODUCT_CLI, glfdir, str(tmp_path), case)
    case = ("cmd4_batched", "test.ped", ["--denovo", "--rate_denovo", "1.5e-07", "--batch_sites", "1000"], "golden_cmd4.vcf.gz")
    U.check_case(U.PRODUCT_CLI, glfdir, str(tmp_path), case)


def test_product_cli_vcf_input_on_all_gpus(tmp_path):
    """--in_vcf --gpus N (every GPU of the box; one context on a one-GPU box): each chunk's records are cut into N
    contiguous ranges, one per GPU; small chunks, so that some hold fewer computed records than there are GPUs."""
    import torch
    n = min(max(1, torch.cuda.device_count()), 8)
    for case in U.VCF_CASES:
        U.check_vcf_case(U.PRODUCT_CLI, str(tmp_path), case, gz_input=(case[0] == "vcf_cmd2"), extra=["--gpus", str(n), "--batch_sites", "5"])


def test_cli_reports_missing_inputs(tmp_path):
    p = subprocess.run([U.PRODUCT_CLI, "-p", "nope.ped", "-d", "nope.dat", "-g", "nope.gif", "--out_vcf", str(tmp_path / "o.vcf")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert p.returncode == 1 and b"FATAL ERROR" in p.stdout


@pytest.mark.parametrize("name", U.BASELINE_SHAPES)
def test_product_cli_matches_reference_on_baseline_shapes(name, glfdir, tmp_path):
    """The drop-in executable (block-per-site kernel, GLF ingest / VCF tokeniser, writers) against outputs of the
    unmodified reference on BASELINE.json's configs 2, 4 and 5 in miniature."""
    U.check_baseline_shape(U.PRODUCT_CLI, str(tmp_path), name)
