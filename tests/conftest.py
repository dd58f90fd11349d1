import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def smf():
    """The product package, with the in-tree library built if it is missing."""
    import subprocess
    so = os.path.join(ROOT, "sparse_matrix_with_flops_b200", "libb200spgemm.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "sparse_matrix_with_flops_b200", "csrc")])
    import sparse_matrix_with_flops_b200 as pkg
    return pkg


@pytest.fixture(scope="session")
def gpu(smf):
    """Initialised GPU context; the C-ABI fails loudly when there is no device."""
    smf.init(0)
    return smf


@pytest.fixture
def b200_options(gpu, monkeypatch):
    """Set B200_* developer switches for one test: the library reads them once (b200_init), so
    they are re-read after every change and again after the test's environment is restored."""
    def set_options(**kw):
        for k, v in kw.items():
            monkeypatch.setenv(k, str(v))
        gpu.reload_options()
    yield set_options
    monkeypatch.undo()
    gpu.reload_options()
