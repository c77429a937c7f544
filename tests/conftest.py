"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs on a CPU-only box (oracle vs golden vectors, host logic,
C-ABI symbol checks, gloo world_size-2 sharding); ``-m gpu`` needs a B200 and
calls the CUDA path through the C-ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
BANDS = ["Sub-Bass", "Bass", "Low Mids", "High Mids", "Presence", "Brilliance"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def gains_dict(values, keys=BANDS):
    return {k: float(v) for k, v in zip(keys, values)}


@pytest.fixture(scope="session")
def golden_src():
    return load_golden("src.npz")


@pytest.fixture(scope="session")
def golden_eq():
    return load_golden("eq.npz")


@pytest.fixture(scope="session")
def golden_fft():
    return load_golden("fft.npz")


@pytest.fixture(scope="session")
def golden_spectrum():
    return load_golden("spectrum.npz")


@pytest.fixture(scope="session")
def golden_design():
    return load_golden("design.npz")


def c1_input(g):
    """C1's stand-in clip (1 channel x 30 s @ 44.1 kHz), remade from the seed the golden file records
    (the file does not carry 5 MB of noise) and checked against the fingerprint make_golden.py stored."""
    x = np.random.default_rng(int(g["seed"])).uniform(-1, 1, int(g["n"])).astype(np.float32)
    x = x / np.max(np.abs(x))
    assert np.allclose([np.sum(x.astype(np.float64)), x[0], x[-1]], g["x_check"], rtol=0, atol=1e-9)
    return x


@pytest.fixture(scope="session")
def golden_chain():
    g = dict(load_golden("chain_c1.npz"))
    g["x"] = c1_input(g)
    return g


@pytest.fixture(scope="session")
def golden_app():
    return load_golden("app_helpers.npz")
