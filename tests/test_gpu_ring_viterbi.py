"""Parity of the time-parallel ring Viterbi engine against the CPU oracle, through
the C ABI.  Bars (BASELINE.json north_star): x bit-exact; ll within 1e-9 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9


def _check(hm, O, S, lA, mu, sig, **kw):
    x, ll, info = hm.viterbi(S, lA, mu, sig, mode="ring", return_info=True)
    xo, llo = O.viterbi(S, lA, mu, sig)
    assert info["engine"] == 2
    bad = np.nonzero(x != xo)[0]
    assert bad.size == 0, f"{bad.size} mismatches, first at {bad[:5]} (chunks={info['n_chunks']})"
    assert abs(ll - llo) <= LL_RTOL * abs(llo), (ll, llo)
    return info


@pytest.mark.parametrize("N,K,T,seed", [(3, 60, 20000, 1234), (3, 60, 300000, 2), (4, 48, 200000, 7),
                                        (5, 60, 150000, 5), (1, 20, 50000, 3), (2, 10, 30000, 1), (7, 60, 60000, 9),
                                        (2, 4, 5000, 12), (6, 33, 40000, 13), (3, 97, 50000, 15)])
def test_ring_viterbi_matches_oracle(hm, O, case_factory, N, K, T, seed):
    hm.set_ring_params(0, 0)
    S, lA, mu, sig = case_factory(N, K, T, seed)
    _check(hm, O, S, lA, mu, sig)


@pytest.mark.parametrize("chunk,warm", [(1024, 256), (2048, 512), (4096, 256), (512, 512)])
@pytest.mark.parametrize("N,K,seed", [(3, 60, 31), (4, 48, 32), (2, 12, 33)])
def test_ring_viterbi_many_small_chunks(hm, O, case_factory, N, K, seed, chunk, warm):
    """Hundreds of chunk boundaries: exercises speculation, verification and repair."""
    S, lA, mu, sig = case_factory(N, K, 120000, seed)
    try:
        hm.set_ring_params(chunk, warm)
        info = _check(hm, O, S, lA, mu, sig)
        assert info["n_chunks"] >= 120000 // max(chunk, 4 * 256) - 1
    finally:
        hm.set_ring_params(0, 0)


def test_ring_viterbi_dense_spiking_forces_repairs(hm, O, case_factory):
    """High firing rates leave few quiet gaps, so speculative starts fail to
    coalesce and the repair path must produce the exact answer."""
    S, lA, mu, sig = case_factory(3, 60, 100000, 41, rate_scale=8.0)
    try:
        hm.set_ring_params(1024, 256)
        info = _check(hm, O, S, lA, mu, sig)
    finally:
        hm.set_ring_params(0, 0)
    assert info["n_chunks"] > 50


@pytest.mark.parametrize("N,K,seed", [(3, 60, 61), (5, 60, 62), (2, 12, 63)])
def test_ring_viterbi_wrong_speculation_is_detected_and_repaired(hm, O, case_factory, monkeypatch, N, K, seed):
    """The exactness argument IS the verify/repair path, so it must be seen to run: with the warm-up / look-ahead
    forced to 0 (HMMCUDA_DEBUG_WARMUP) every chunk starts from an empty state and every traceback chunk ends in
    an assumed noise state -- wrong wherever a spike straddles a boundary.  Verification has to catch each of
    them and the sequential repair has to restore the oracle's path."""
    S, lA, mu, sig = case_factory(N, K, 150000, seed, rate_scale=3.0)
    monkeypatch.setenv("HMMCUDA_DEBUG_WARMUP", "0")
    try:
        hm.set_ring_params(2048, 256)
        info = _check(hm, O, S, lA, mu, sig)
    finally:
        hm.set_ring_params(0, 0)
    assert info["n_chunks"] >= 70
    assert info["fwd_repaired"] > 0 and info["bwd_repaired"] > 0, info


def test_ring_viterbi_forced_flags_exercise_partial_repair(hm, O, case_factory, monkeypatch):
    """HMMCUDA_DEBUG_FLAG_EVERY=3: the boundary checks also flag every third chunk, so repaired (exact-start) and
    accepted (speculative) chunks alternate; the decode must not change."""
    S, lA, mu, sig = case_factory(3, 60, 200000, 71)
    monkeypatch.setenv("HMMCUDA_DEBUG_FLAG_EVERY", "3")
    try:
        hm.set_ring_params(4096, 512)
        info = _check(hm, O, S, lA, mu, sig)
    finally:
        hm.set_ring_params(0, 0)
    assert info["fwd_repaired"] >= info["n_chunks"] // 3 - 1 and info["bwd_repaired"] > 0, info


def test_pipelined_host_decode_repairs(hm, O, case_factory, monkeypatch):
    """The host-pointer pipeline (segments = time shards with ghost chunks, api.cu viterbi_host_pipelined) with forced
    flags: repairs inside the segments and the right-to-left traceback re-link must give the oracle's path."""
    S, lA, mu, sig = case_factory(3, 60, (1 << 22) + 12345, 81)
    hm.set_ring_params(0, 0)
    monkeypatch.setenv("HMMCUDA_DEBUG_FLAG_EVERY", "5")
    x, ll, info = hm.viterbi(S, lA, mu, sig, mode="ring", return_info=True)
    monkeypatch.delenv("HMMCUDA_DEBUG_FLAG_EVERY")
    assert info["fwd_repaired"] > 0 and info["bwd_repaired"] > 0, info
    xo, llo = O.viterbi(S, lA, mu, sig)
    assert np.array_equal(x, xo) and abs(ll - llo) <= LL_RTOL * abs(llo)


def test_ring_viterbi_fitted_like_model(hm, O, case_factory):
    """A mis-specified model (scaled templates, wrong sigma, mu row 1 non-zero)."""
    S, lA, mu, sig = case_factory(3, 60, 80000, 51)
    mu2 = np.asfortranarray(mu * 0.8)
    mu2[0, :] = [0.01, -0.02, 0.005]
    _check(hm, O, S, lA, mu2, 0.35)


def test_ring_viterbi_auto_and_batch(hm, O, case_factory):
    cases = [case_factory(4, 48, 40000, 200 + c) for c in range(6)]
    Y = np.asfortranarray(np.stack([c[0] for c in cases], axis=1))
    models = [(c[1], np.asfortranarray(c[2] * (1 + 0.05 * i)), 0.3 + 0.01 * i) for i, c in enumerate(cases)]
    x, ll, info = hm.viterbi_batch(Y, models, mode="auto", return_info=True)
    assert info["engine"] == 2  # ring-structured and T >= 2048 -> ring engine
    for c in range(6):
        xo, llo = O.viterbi(Y[:, c], models[c][0], models[c][1], models[c][2])
        assert np.array_equal(x[:, c], xo)
        assert abs(ll[c] - llo) <= LL_RTOL * abs(llo)


def test_ring_range_limits_fall_back(hm, O, case_factory):
    """N = 8 is outside the ring engine (7 neurons max): explicit ring mode refuses, auto mode decodes with the
    time-parallel per-state engine (T >= 4096) or, for short sequences, the sequential one."""
    S, lA, mu, sig = case_factory(8, 40, 40000, 14)
    with pytest.raises(hm.HmmError) as ei:
        hm.viterbi(S, lA, mu, sig, mode="ring")
    assert ei.value.code == hm._lib.HMM_EUNSUPPORTED
    x, ll, info = hm.viterbi(S, lA, mu, sig, mode="auto", return_info=True)
    xo, llo = O.viterbi(S, lA, mu, sig)
    assert info["engine"] == 3 and np.array_equal(x, xo) and abs(ll - llo) <= LL_RTOL * abs(llo)
    x, ll, info = hm.viterbi(S[:3000], lA, mu, sig, mode="auto", return_info=True)
    xo, llo = O.viterbi(S[:3000], lA, mu, sig)
    assert info["engine"] == 1 and np.array_equal(x, xo) and ll == llo


def test_ring_rejects_overlap_model(hm):
    lA = hm.StateMatrix(2, 6, np.log([0.01, 0.02]), True)
    mu = np.asfortranarray(np.zeros((6, 2)))
    with pytest.raises(hm.HmmError) as ei:
        hm.viterbi(np.zeros(5000), lA, mu, 0.3, mode="ring")
    assert ei.value.code == hm._lib.HMM_EUNSUPPORTED


def test_ring_equals_faithful_1m(hm, case_factory):
    """A 1 M-sample cut of config 2: ring engine == faithful engine (bit-exact x)."""
    S, lA, mu, sig = case_factory(3, 60, 1_000_000, 2)
    xr, llr, info = hm.viterbi(S, lA, mu, sig, mode="ring", return_info=True)
    xf, llf = hm.viterbi(S, lA, mu, sig, mode="faithful")
    assert np.array_equal(xr, xf)
    assert abs(llr - llf) <= LL_RTOL * abs(llf)


@pytest.mark.parametrize("N,K,T,seed,chunk", [(3, 60, 70_001, 501, 0), (4, 48, 50_123, 502, 2048), (5, 60, 33_333, 503, 1024),
                                              (7, 97, 40_000, 504, 0), (1, 4, 9_999, 505, 512), (2, 33, 262_144 + 17, 506, 0)])
def test_no_kernel_writes_outside_its_buffers(hm, O, case_factory, monkeypatch, N, K, T, seed, chunk):
    """compute-sanitizer is not available on this GPU pool, so the decode checks itself: with HMMCUDA_DEBUG_GUARD=1
    every buffer of the decode plan lies between two guard zones that are verified after the run (the call fails if a
    kernel wrote out of bounds), on ragged lengths, the smallest and the largest models, forced repairs included."""
    S, lA, mu, sig = case_factory(N, K, T, seed)
    monkeypatch.setenv("HMMCUDA_DEBUG_GUARD", "1")
    monkeypatch.setenv("HMMCUDA_DEBUG_FLAG_EVERY", "4")
    try:
        hm.set_ring_params(chunk, 0)
        for _ in range(2):  # the second call re-uses the cached program (CUDA graph)
            _check(hm, O, S, lA, mu, sig)
    finally:
        hm.set_ring_params(0, 0)
        hm.lib().hmm_release_workspace()
