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


def golden_model(name):
    g = np.load(os.path.join(GOLDEN, "model_%s.npz" % name))
    return g["theta"], g["pi"], g["T"], g["E"]


def example_symbols():
    return np.load(os.path.join(GOLDEN, "example_pair.npz"))["symbols"]


def random_hmm(rng, K, S=3, missing_col=True):
    """A random reversible-or-not HMM with strictly positive entries."""
    pi = rng.dirichlet(np.ones(K))
    T = rng.dirichlet(np.ones(K) * 0.7, size=K)
    E = rng.dirichlet(np.ones(S), size=K)
    if missing_col and S == 3:
        E[:, 2] = 1.0
    return pi, T, E


def synthetic_sequence(rng, pi, T, E, L, missing=0.04, mean_run=100):
    """Sample a hidden path from (pi, T), symbols from E[:, :2], overlay missing-data runs."""
    K = pi.size
    cum_T = np.cumsum(T, axis=1)
    states = np.empty(L, dtype=np.int64)
    s = rng.choice(K, p=pi / pi.sum())
    u = rng.random(L)
    for t in range(L):
        states[t] = s
        s = min(int(np.searchsorted(cum_T[s], u[t])), K - 1)
    p1 = E[states, 1] / (E[states, 0] + E[states, 1])
    obs = (rng.random(L) < p1).astype(np.uint8)
    if missing > 0:
        t = 0
        while t < L:
            gap = rng.geometric(missing / (mean_run * (1 - missing)))
            t += gap
            run = rng.geometric(1.0 / mean_run)
            obs[t:t + run] = 2
            t += run
    return obs


@pytest.fixture(scope="session")
def have_gpu():
    import imcoalhmm_b200 as m
    import ctypes
    n = ctypes.c_int()
    m._lib.load().imc_device_count(ctypes.byref(n))
    return n.value > 0
