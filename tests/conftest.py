import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def hm():
    """The product package (ctypes mirror over libhmmcuda.so)."""
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libhmmcuda.so")):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    o = ge.load_oracle()
    o.build()
    return o


TEMPLATE_PARAMS = [(3.0, 0.8, 0.2), (4.0, 0.3, 0.2), (2.0, 0.5, 0.3), (2.5, 0.6, 0.25), (3.5, 0.4, 0.15),
                   (1.8, 0.9, 0.3), (2.8, 0.2, 0.1), (2.2, 0.7, 0.12)]
RATES = [0.003, 0.001, 0.002, 0.0015, 0.0025, 0.001, 0.002, 0.0012]


def make_case(hm, N, K, T, seed, sigma=0.3, rate_scale=1.0):
    """Synthetic recording + true model (SURVEY 8d style)."""
    temps = np.stack([hm.create_spike_template(K, *TEMPLATE_PARAMS[i]) for i in range(N)], axis=1)
    pp = np.array(RATES[:N]) * rate_scale
    S = hm.create_signal(T, sigma, pp, temps, hm.make_rng(seed))
    mu = np.asfortranarray(temps.copy())
    mu[0, :] = 0.0
    lA = hm.StateMatrix(N, K, np.log(pp), False)
    return S, lA, mu, sigma


@pytest.fixture(scope="session")
def case_factory(hm):
    return lambda *a, **k: make_case(hm, *a, **k)
