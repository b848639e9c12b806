import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _make(target):
    subprocess.run(["make", "-s", target], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle_built():
    """The CPU oracle (test infrastructure) — built on demand with gcc."""
    _make("oracle")
    return os.path.join(ROOT, "oracle", "_build")


@pytest.fixture(scope="session")
def tools_built():
    _make("tools")
    return os.path.join(ROOT, "polymutt_b200", "bin", "pm-tools")


@pytest.fixture(scope="session")
def example12():
    from polymutt_b200 import load_pmpk
    return load_pmpk(os.path.join(GOLDEN, "example12.pmpk.gz"))


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
