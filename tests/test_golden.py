"""Committed fixtures (tests/golden/oracle_small.npz, written by tests/golden/make_golden.py): the oracle must
keep reproducing them (CPU), and the CUDA path must match them through the C ABI (GPU) -- state sequences bit
for bit, ll within 1e-9 relative, one E/M step within 1e-8."""
import hashlib
import os

import numpy as np
import pytest

from conftest import make_case
from golden.make_golden import CASES

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_small.npz"))


def _sha(x):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest(), dtype=np.uint8)


@pytest.mark.parametrize("name,N,K,T,seed,rs", CASES)
def test_oracle_reproduces_golden(hm, O, name, N, K, T, seed, rs):
    S, lA, mu, sig = make_case(hm, N, K, T, seed, rate_scale=rs)
    x, ll = O.viterbi(S, lA, mu, sig)
    assert np.array_equal(_sha(x), G[f"{name}_x_sha"]) and np.array_equal(x[:4096], G[f"{name}_x_head"])
    assert abs(ll - float(G[f"{name}_ll"])) <= 1e-12 * abs(ll)
    lp, pp, mu1, s1, llk = O.em_step(S, O.OracleStateMatrix(N, K, np.log(np.full(N, 0.01)), False),
                                     np.asfortranarray(0.7 * mu), float(np.std(S)))
    assert np.abs(lp - G[f"{name}_em_lp"]).max() < 1e-10 and np.abs(mu1 - G[f"{name}_em_mu"]).max() < 1e-10
    assert abs(s1 - float(G[f"{name}_em_sigma"])) < 1e-12 and abs(llk - float(G[f"{name}_em_loglik"])) <= 1e-12 * abs(llk)


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,K,T,seed,rs", CASES)
def test_cuda_matches_golden(hm, name, N, K, T, seed, rs):
    S, lA, mu, sig = make_case(hm, N, K, T, seed, rate_scale=rs)
    for mode in ("ring", "faithful"):
        x, ll = hm.viterbi(S, lA, mu, sig, mode=mode)
        assert np.array_equal(_sha(x), G[f"{name}_x_sha"]), mode
        assert abs(ll - float(G[f"{name}_ll"])) <= 1e-9 * abs(ll), mode
    lA0 = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), False)
    lp, pp, mu1, s1, llk = hm.em_step(S, lA0, np.asfortranarray(0.7 * mu), float(np.std(S)))[:5]
    assert np.abs(lp - G[f"{name}_em_lp"]).max() < 1e-8 and np.abs(mu1 - G[f"{name}_em_mu"]).max() < 1e-8
    assert abs(s1 - float(G[f"{name}_em_sigma"])) < 1e-8
    assert abs(llk - float(G[f"{name}_em_loglik"])) <= 1e-9 * abs(llk)
