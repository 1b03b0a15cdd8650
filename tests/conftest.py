"""Shared fixtures.  `-m "not gpu"` = oracle vs golden vectors, host logic, ABI symbol check;
`-m gpu` = parity of the CUDA path (through the C-ABI) against the oracle and the goldens."""
import glob
import json
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


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def big_case_names():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "big_*.npz")))


def unpack_mask_bits(bits, shape):
    """Inverse of np.packbits(mask.ravel()) for a mask of `shape`."""
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape)


def e2e_case_names():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "e2e_*.npz")))


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "huffman_kats.json")) as f:
        return json.load(f)


STREAMS = ("indices_coarse", "indices_medium", "indices_fine", "mask_coarse", "mask_medium")
