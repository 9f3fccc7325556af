"""FP32 mode (north_star: "T1 and log-likelihoods must agree within ... 1e-4 in FP32 mode"): Float32 recording, the
FIR of the ring decode in FP32, everything else FP64.  The oracle decodes the same Float32-rounded values in FP64.
ll must agree to 1e-4 relative; x is not promised bit-exact -- the mismatch rate is measured and bounded."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LL_RTOL_F32 = 1e-4


@pytest.mark.parametrize("N,K,T,seed", [(3, 60, 400_000, 91), (4, 48, 300_000, 92), (5, 60, 300_000, 93), (2, 20, 100_000, 94)])
def test_fp32_decode_against_oracle(hm, O, case_factory, N, K, T, seed):
    S, lA, mu, sig = case_factory(N, K, T, seed)
    S32 = S.astype(np.float32)
    x, ll, info = hm.viterbi_f32(S32, lA, mu, sig, mode="ring", return_info=True)
    xo, llo = O.viterbi(S32.astype(np.float64), lA, mu, sig)
    assert info["engine"] == 2
    assert abs(ll - llo) <= LL_RTOL_F32 * abs(llo), (ll, llo)
    rate = float(np.mean(x != xo))
    assert rate < 2e-3, f"x mismatch rate {rate:.2e} in FP32 mode"
    # FP64 decode of the same values through the same entry point family: bit-exact again
    x64, ll64 = hm.viterbi(S32.astype(np.float64), lA, mu, sig, mode="ring")
    assert np.array_equal(x64, xo) and abs(ll64 - llo) <= 1e-9 * abs(llo)


def test_set_precision_switches_the_f64_entry_points(hm, O, case_factory):
    S, lA, mu, sig = case_factory(3, 60, 300_000, 95)
    xo, llo = O.viterbi(S, lA, mu, sig)
    try:
        hm.set_precision("f32")
        x, ll = hm.viterbi(S, lA, mu, sig, mode="ring")
        assert abs(ll - llo) <= LL_RTOL_F32 * abs(llo) and float(np.mean(x != xo)) < 2e-3
        # back to FP64: bit-exact again
        hm.set_precision("f64")
        x2, ll2 = hm.viterbi(S, lA, mu, sig, mode="ring")
        assert np.array_equal(x2, xo)  # (ll is a function of (x, y, model) only: an identical path gives the identical ll)
    finally:
        hm.set_precision("f64")
