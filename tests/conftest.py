import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Parity tolerances of BASELINE.json's north star.
TOL_BLOCH = 1e-9      # |mx,my,mz - reference| absolute, fp64
TOL_SLR = 1e-9        # |alpha,beta - reference| absolute, fp64


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_files(pattern):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, pattern)))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (restatement + compiled reference where present).  Test infrastructure only."""
    from oracle import ref
    ref.build()
    return ref


@pytest.fixture(scope="session")
def mbrf():
    """The product package with the C-ABI library loaded; GPU tests call through it."""
    import multiband_rf_pulse_design_b200 as m
    m.lib()
    return m
