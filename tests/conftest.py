import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def ckpt_att2s():
    return load_npz("ckpt_att2s_v3.npz")


@pytest.fixture(scope="session")
def ckpt_aggr():
    return load_npz("ckpt_aggr_v2p.npz")


@pytest.fixture(scope="session")
def golden_synth():
    return load_npz("att2s_synth.npz")


@pytest.fixture(scope="session")
def golden_edge():
    return load_npz("att2s_edge.npz")


@pytest.fixture(scope="session")
def golden_seeded():
    return load_npz("att2s_seeded.npz")


@pytest.fixture(scope="session")
def golden_batchloop():
    return load_npz("att2s_batchloop.npz")


@pytest.fixture(scope="session")
def golden_aggr():
    return load_npz("aggr_synth.npz")
