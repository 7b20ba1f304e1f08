import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def solver_cases():
    return np.load(os.path.join(GOLDEN, "solver_cases.npz"))


@pytest.fixture(scope="session")
def ukbb():
    return np.load(os.path.join(GOLDEN, "ukbb_subset.npz"))


@pytest.fixture(scope="session")
def ref_synth():
    return np.load(os.path.join(GOLDEN, "ref_synthetic.npz"))


SOLVER_CASE_NAMES = [
    "s1_pos", "s1_signed", "s2_pos", "s2_signed", "s2_iso", "s2_dup", "s3_pos_iso",
    "s3_signed", "s3_pos_ear", "s3_dup", "s4_pos", "s4_signed", "s5_pos",
]
