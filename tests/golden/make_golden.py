"""Regenerates tests/golden/oracle_small.npz.

The reference is a Julia package and cannot run in this environment (no Julia toolchain), and its own tests hold
no output vectors for the hot path, so these fixtures are ORACLE-generated, not reference-generated:
they freeze what `oracle/hmm_oracle.c` (the literal restatement of src/viterbi.jl, src/baumwelch.jl,
src/reconstruction.jl) produced when it was checked against the reference's known-answer tests
(tests/test_oracle.py).  They guard the oracle against drift (compiler, flags) and give the GPU parity tests a
target that does not depend on building the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from conftest import make_case  # noqa: E402

CASES = [  # name, N, K, T, seed, rate_scale
    ("n3k60", 3, 60, 40_000, 101, 1.0),
    ("n2k10", 2, 10, 6_000, 102, 4.0),
    ("n5k60", 5, 60, 50_000, 103, 1.0),
    ("n4k48", 4, 48, 36_000, 104, 1.0),
]


def main():
    hm = ge.load_package()  # only its pure-numpy synthetic-data helpers are used here (no device needed)
    O = ge.load_oracle()
    O.build()
    out = {}
    for name, N, K, T, seed, rs in CASES:
        S, lA, mu, sig = make_case(hm, N, K, T, seed, rate_scale=rs)
        x, ll = O.viterbi(S, lA, mu, sig)
        lp0 = np.log(np.full(N, 0.01))
        mu0 = np.asfortranarray(0.7 * mu)
        s0 = float(np.std(S))
        lp, pp, mu1, s1, llk = O.em_step(S, O.OracleStateMatrix(N, K, lp0, False), mu0.copy(order="F"), s0)
        out[f"{name}_x_sha"] = np.frombuffer(__import__("hashlib").sha256(np.ascontiguousarray(x).tobytes()).digest(),
                                             dtype=np.uint8)
        out[f"{name}_x_head"] = x[:4096].copy()
        out[f"{name}_ll"] = np.float64(ll)
        out[f"{name}_em_lp"] = lp
        out[f"{name}_em_mu"] = mu1
        out[f"{name}_em_sigma"] = np.float64(s1)
        out[f"{name}_em_loglik"] = np.float64(llk)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
