"""BASELINE.json full-size configurations through the C ABI, against the CPU oracle AT the BASELINE sizes:
  * config 2: the whole 18 M-sample decode, x bit-exact and ll to 1e-9, plus the near-tie screen of the input;
  * config 4: two full 18 M-sample channels of a batched (N=4, K=48) decode;
  * config 5: a 20 M-sample prefix of the N=5 recording decoded by both, and the full 108 M-sample decode
    against that prefix away from the cut;
  * config 3: one E/M iteration at T = 1.8 M and 20 iterations at T = 200 k, fits within 1e-6;
  * config 1 (README example) end to end.
The oracle takes tens of seconds per case at these sizes (it is the reference's sequential algorithm); independent
oracle calls run on host threads.  Size-independent properties (chunk-geometry invariance, path structure, an
independent numpy recomputation of ll) are kept as well."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9     # north_star: T1 / log-likelihoods within 1e-9 relative (FP64)
FIT_ATOL = 1e-6    # north_star: fitted mu / sigma / lA within 1e-6
EPS64 = 64 * 2.220446049250313e-16


def _c2(hm, T, seed):
    K, N = 60, 3
    temps = np.stack([hm.create_spike_template(K, 3.0, 0.8, 0.2), hm.create_spike_template(K, 4.0, 0.3, 0.2),
                      hm.create_spike_template(K, 2.0, 0.5, 0.3)], axis=1)
    pp = np.array([0.003, 0.001, 0.002])
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(seed))
    mu = np.asfortranarray(temps.copy())
    mu[0, :] = 0.0
    return S, hm.StateMatrix(N, K, np.log(pp), False), mu, 0.3


def _ll_numpy(y, x, lA, mu, sigma):
    """sum_{i=2..T} T1[x[i], i] along the decoded path (src/viterbi.jl:92-96), vectorised."""
    ns = lA.nstates
    W = np.full((ns, ns), np.nan)
    W[lA.transitions["src"] - 1, lA.transitions["dst"] - 1] = lA.transitions["lp"]
    m = np.array([sum(mu[lA.states[l, j] - 1, l] for l in range(lA.N)) for j in range(ns)])
    xs = x.astype(np.int64) - 1
    w = W[xs[:-1], xs[1:]]
    assert not np.isnan(w).any(), "decoded path uses a transition that is not in the StateMatrix"
    q = -0.9189385332046727 - np.log(sigma) - (y - m[xs]) ** 2 / (2 * sigma * sigma)
    p0 = 0.0 if xs[0] == 0 else q[0]
    inc = w + q[1:]
    T = y.size
    return (T - 1) * p0 + float(np.dot(np.arange(T - 1, 0, -1, dtype=np.float64), inc))


def _check_structure(x, lA):
    """Inside a chain the state index increases by one per step; chains start at a head."""
    L = lA.K - 1
    xs = x.astype(np.int64) - 1
    ph = np.where(xs > 0, (xs - 1) % L, -1)  # 0-based phase, -1 for noise
    nxt_ok = (ph[:-1] < 0) | (ph[:-1] == L - 1) | (xs[1:] == xs[:-1] + 1)
    assert nxt_ok.all()
    starts_ok = (ph[1:] != 0) | (ph[:-1] < 0) | (ph[:-1] == L - 1)
    assert starts_ok.all()


def _same_path(x, xo, what):
    bad = np.nonzero(x != xo)[0]
    assert bad.size == 0, f"{what}: {bad.size} of {x.size} states differ from the oracle, first at {bad[:5]}"


def test_config2_full_size_vs_oracle(hm, O):
    """The BASELINE config-2 decode (18 M samples, N=3, K=60) against the oracle's decode of the SAME 18 M samples."""
    T = 18_000_000
    S, lA, mu, sig = _c2(hm, T, 2)
    hm.set_ring_params(0, 0)
    with ThreadPoolExecutor(2) as ex:  # the oracle runs on host threads while the GPU decodes
        f_dec = ex.submit(O.viterbi, S, lA, mu, sig)
        f_scr = ex.submit(O.viterbi_screen, S, lA, mu, sig)
        x, ll, info = hm.viterbi(S, lA, mu, sig, mode="ring", return_info=True)
        xo, llo = f_dec.result()
        scr = f_scr.result()
    assert info["engine"] == 2 and x.shape == (T,)
    # the input has no decision that the reference's own rounding noise would decide (SURVEY 8d screen)
    assert scr["n_below_rel_b"] == 0 and scr["min_rel_margin"] > EPS64, scr
    _same_path(x, xo, "config 2")
    assert abs(ll - llo) <= LL_RTOL * abs(llo), (ll, llo)
    try:
        hm.set_ring_params(12288, 1024)
        x2, ll2, info2 = hm.viterbi(S, lA, mu, sig, mode="ring", return_info=True)
    finally:
        hm.set_ring_params(0, 0)
    assert info2["n_chunks"] != info["n_chunks"]
    assert np.array_equal(x, x2) and abs(ll - ll2) <= 1e-12 * abs(ll)
    _check_structure(x, lA)
    assert abs(ll - _ll_numpy(S, x, lA, mu, sig)) <= 1e-9 * abs(ll)
    # prefix vs the faithful engine (bit-exact reference order); the decode of a prefix can
    # only differ from the full decode near the cut
    Tp = 2_000_000
    xf, _ = hm.viterbi(S[:Tp], lA, mu, sig, mode="faithful")
    assert np.array_equal(xf[:Tp - 4096], x[:Tp - 4096])
    Y = hm.reconstruct_signal(x, lA, mu, sig)
    assert 0.5 < 1 - np.std(Y - S) / np.std(S) < 0.7


def test_config5_long_sequence_vs_oracle(hm, O):
    """Config 5: single channel, 1 h at 30 kHz (108 M samples), N=5 x K=60."""
    T, N, K = 108_000_000, 5, 60
    params = [(3.0, 0.8, 0.2), (4.0, 0.3, 0.2), (2.0, 0.5, 0.3), (2.5, 0.6, 0.25), (3.5, 0.4, 0.15)]
    temps = np.stack([hm.create_spike_template(K, *q) for q in params], axis=1)
    pp = np.array([0.003, 0.001, 0.002, 0.0015, 0.0025])
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(5))
    mu = np.asfortranarray(temps.copy())
    mu[0, :] = 0.0
    lA = hm.StateMatrix(N, K, np.log(pp), False)
    x, ll, info = hm.viterbi(S, lA, mu, 0.3, mode="ring", return_info=True)
    try:
        hm.set_ring_params(65536, 512)
        x2, ll2, _ = hm.viterbi(S, lA, mu, 0.3, mode="ring", return_info=True)
    finally:
        hm.set_ring_params(0, 0)
    assert np.array_equal(x, x2)
    del x2
    _check_structure(x, lA)
    # a 20 M-sample prefix decoded by the oracle and by the GPU; the full decode agrees with it away from the cut
    Tq = 20_000_000
    with ThreadPoolExecutor(2) as ex:
        f_dec = ex.submit(O.viterbi, S[:Tq], lA, mu, 0.3)
        f_scr = ex.submit(O.viterbi_screen, S[:Tq], lA, mu, 0.3)
        xq, llq = hm.viterbi(S[:Tq], lA, mu, 0.3, mode="ring")
        xo, llo = f_dec.result()
        scr = f_scr.result()
    assert scr["n_below_rel_b"] == 0, scr
    _same_path(xq, xo, "config 5 prefix")
    assert abs(llq - llo) <= LL_RTOL * abs(llo), (llq, llo)
    assert np.array_equal(x[:Tq - 8192], xo[:Tq - 8192])
    assert abs(llq - _ll_numpy(S[:Tq], xq, lA, mu, 0.3)) <= 1e-9 * abs(llq)


def _c4_channels(hm, C, T, seed0=1000):
    N, K = 4, 48
    rng = np.random.default_rng(seed0)
    models, cols = [], []
    for c in range(C):
        prm = [(rng.uniform(2, 4), rng.uniform(0.3, 0.9), rng.uniform(0.1, 0.3)) for _ in range(N)]
        temps = np.stack([hm.create_spike_template(K, *q) for q in prm], axis=1)
        pp = rng.uniform(0.0005, 0.004, size=N)
        cols.append(hm.create_signal(T, 0.3, pp, temps, hm.make_rng(seed0 + c)))
        mu = np.asfortranarray(temps.copy())
        mu[0, :] = 0.0
        models.append((hm.StateMatrix(N, K, np.log(pp), False), mu, 0.3))
    return np.asfortranarray(np.stack(cols, axis=1)), models


def test_config4_two_full_size_channels_vs_oracle(hm, O):
    """Config 4 at the BASELINE length: two independent 18 M-sample channels (N=4, K=48, own templates and
    rates) through the batch entry point, each against the oracle's decode of the same 18 M samples."""
    T = 18_000_000
    Y, models = _c4_channels(hm, 2, T)
    with ThreadPoolExecutor(2) as ex:
        fo = [ex.submit(O.viterbi, Y[:, c], *models[c]) for c in range(2)]
        x, ll, info = hm.viterbi_batch(Y, models, mode="ring", return_info=True)
        ref = [f.result() for f in fo]
    assert info["engine"] == 2
    for c in range(2):
        _same_path(x[:, c], ref[c][0], f"config 4 channel {c}")
        assert abs(ll[c] - ref[c][1]) <= LL_RTOL * abs(ref[c][1])


def test_config3_one_iteration_full_size_vs_oracle(hm, O):
    """Config 3 at the BASELINE length: one E/M step on T = 1.8 M samples (the oracle materialises alpha, beta
    and gamma: 7.7 GB of host memory, about a minute)."""
    T, N, K = 1_800_000, 3, 60
    S, _, mu_true, _ = _c2(hm, T, 3)
    lp0 = np.log(np.full(N, 0.01))
    mu0 = np.asfortranarray(0.7 * mu_true)
    s0 = float(np.std(S))
    with ThreadPoolExecutor(1) as ex:
        fo = ex.submit(O.em_step, S, O.OracleStateMatrix(N, K, lp0, False), mu0.copy(order="F"), s0)
        lp, pp, mu1, s1, ll, info = hm.em_step(S, hm.StateMatrix(N, K, lp0, False), mu0.copy(order="F"), s0,
                                               mode="ring", return_info=True)
        lpo, ppo, muo, so, llo = fo.result()
    assert info["engine"] == 2
    assert np.abs(mu1 - muo).max() < FIT_ATOL and abs(s1 - so) < FIT_ATOL and np.abs(lp - lpo).max() < FIT_ATOL
    assert abs(ll - llo) <= LL_RTOL * abs(llo), (ll, llo)
    fin = np.isfinite(ppo)
    assert np.array_equal(np.isfinite(pp), fin) and np.abs(pp[fin] - ppo[fin]).max() < FIT_ATOL


def test_config3_twenty_iterations_vs_oracle(hm, O):
    """BASELINE config 3's 20 Baum-Welch iterations (on T = 200 k so that the oracle finishes in about a minute):
    every iteration's fit is compared, not only the last one."""
    T, N, K = 200_000, 3, 60
    S, _, mu_true, _ = _c2(hm, T, 3)
    lp0 = np.log(np.full(N, 0.01))
    mu0 = np.asfortranarray(0.7 * mu_true)
    s0 = float(np.std(S))
    seen = []
    mu = mu0.copy(order="F")
    lA_fit, mu_fit, s_fit = hm.train_model(S, hm.StateMatrix(N, K, lp0, False), mu, s0, 20,
                                           lambda m: seen.append(m.copy()))
    smo, muo, so = O.OracleStateMatrix(N, K, lp0, False), mu0.copy(order="F"), s0
    for it in range(20):
        assert np.abs(seen[it] - muo).max() < FIT_ATOL, it
        lpo, ppo, muo, so, llo = O.em_step(S, smo, muo, so)
        smo = O.OracleStateMatrix(N, K, lpo, False)
    assert np.abs(mu_fit - muo).max() < FIT_ATOL and abs(s_fit - so) < FIT_ATOL
    assert np.abs(lA_fit.transitions["lp"] - smo.transitions["lp"]).max() < FIT_ATOL


def test_config4_channel_batch(hm, O):
    """Config 4 style: independent per-channel HMMs (N=4, K=48), batched decode; two of the
    channels are checked against the oracle on a prefix-sized recording."""
    C, T, N, K = 6, 400_000, 4, 48
    rng = np.random.default_rng(1000)
    models, cols = [], []
    for c in range(C):
        prm = [(rng.uniform(2, 4), rng.uniform(0.3, 0.9), rng.uniform(0.1, 0.3)) for _ in range(N)]
        temps = np.stack([hm.create_spike_template(K, *q) for q in prm], axis=1)
        pp = rng.uniform(0.0005, 0.004, size=N)
        cols.append(hm.create_signal(T, 0.3, pp, temps, hm.make_rng(1000 + c)))
        mu = np.asfortranarray(temps.copy())
        mu[0, :] = 0.0
        models.append((hm.StateMatrix(N, K, np.log(pp), False), mu, 0.3))
    Y = np.asfortranarray(np.stack(cols, axis=1))
    x, ll, info = hm.viterbi_batch(Y, models, mode="ring", return_info=True)
    assert info["engine"] == 2
    for c in (0, C - 1):
        xo, llo = O.viterbi(Y[:, c], *models[c])
        assert np.array_equal(x[:, c], xo) and abs(ll[c] - llo) <= 1e-9 * abs(llo)
    for c in range(C):
        _check_structure(x[:, c], models[c][0])


def test_config1_readme_example(hm, O):
    """Config 1: create_signal(20_000, 0.3, [0.003, 0.001]) from two 60-sample templates;
    10 E/M steps of a 3-neuron K=60 non-overlap model from an explicit seeded init (the
    reference's own init uses Julia's RNG), then viterbi + reconstruct_signal --
    every stage against the oracle."""
    K, N, T = 60, 3, 20_000
    temps2 = np.stack([hm.create_spike_template(K, 3.0, 0.8, 0.2), hm.create_spike_template(K, 4.0, 0.3, 0.2)], 1)
    S = hm.create_signal(T, 0.3, np.array([0.003, 0.001]), temps2, hm.make_rng(1234))
    rng = np.random.default_rng(7)
    sig0 = float(np.std(S))
    mu0 = np.asfortranarray(np.stack([hm.create_spike_template(K, 3 * sig0 * rng.random(), 0.5 + 0.1 * rng.normal(),
                                                               1.5 * rng.random()) for _ in range(N)], 1))
    mu0[0, :] = 0.0
    p0 = 2.0 ** (-3 * K / 2)  # src/baumwelch.jl:311
    lA = hm.StateMatrix(N, K, np.log(np.full(N, p0)), False)
    mu = mu0.copy(order="F")
    lA_fit, mu_fit, s_fit = hm.train_model(S, lA, mu, sig0, 10)
    smo, muo, so = O.OracleStateMatrix(N, K, np.log(np.full(N, p0)), False), mu0.copy(order="F"), sig0
    for _ in range(10):
        lpo, ppo, muo, so, _ = O.em_step(S, smo, muo, so)
        smo = O.OracleStateMatrix(N, K, lpo, False)
    assert np.abs(mu_fit - muo).max() < 1e-6 and abs(s_fit - so) < 1e-6
    assert np.abs(lA_fit.transitions["lp"] - smo.transitions["lp"]).max() < 1e-6
    x, ll = hm.viterbi(S, lA_fit, mu_fit, s_fit)
    xo, llo = O.viterbi(S, smo, muo, so)
    assert np.array_equal(x, xo) and abs(ll - llo) <= 1e-9 * abs(llo)
    Y = hm.reconstruct_signal(x, lA_fit, mu_fit, s_fit)
    assert np.allclose(Y, O.reconstruct_signal(xo, smo, muo), atol=1e-6)


def test_config1_end_to_end_with_merge_and_prune(hm, O):
    """BASELINE config 1 replayed through the reference's WHOLE training driver (src/baumwelch.jl:324-354): 10 E/M
    steps, condense_templates -> remove_sparse -> remove_small, 5 more E/M steps, then viterbi + reconstruct_signal.
    The host-side merge / prune code is the literal restatement in tests/train_harness.py (quirks included); the
    E/M steps come once from the CPU oracle and once from libhmmcuda.  Same discrete decisions, fits within 1e-6,
    identical decoded path."""
    import train_harness as th

    K, N, T = 60, 3, 20_000
    temps2 = np.stack([hm.create_spike_template(K, 3.0, 0.8, 0.2), hm.create_spike_template(K, 4.0, 0.3, 0.2)], 1)
    S = hm.create_signal(T, 0.3, np.array([0.003, 0.001]), temps2, hm.make_rng(1234))
    rng = np.random.default_rng(7)
    sig0 = float(np.std(S))
    mu0 = np.asfortranarray(np.stack([hm.create_spike_template(K, 3 * sig0 * rng.random(), 0.5 + 0.1 * rng.normal(),
                                                               1.5 * rng.random()) for _ in range(N)], 1))
    mu0[0, :] = 0.0
    sm0 = hm.StateMatrix(N, K, np.log(np.full(N, 2.0 ** (-3 * K / 2))), False)  # p0 of src/baumwelch.jl:311

    def em_gpu(X, sm, mu, sigma):
        return hm.em_step(X, sm, np.asfortranarray(mu), sigma)

    def em_cpu(X, sm, mu, sigma):
        return O.em_step(X, sm, np.asfortranarray(mu), sigma)

    sm_g, mu_g, s_g, log_g = th.train_model(em_gpu, hm.StateMatrix, S, sm0, mu0, sig0, 10)
    sm_c, mu_c, s_c, log_c = th.train_model(em_cpu, hm.StateMatrix, S, sm0, mu0, sig0, 10)
    # the discrete decisions of the merge / prune phase
    assert log_g["after_condense"][0] == log_c["after_condense"][0]
    assert log_g["after_sparse"] == log_c["after_sparse"] and log_g["after_small"] == log_c["after_small"]
    assert np.abs(log_g["after_phase1"][1] - log_c["after_phase1"][1]).max() < FIT_ATOL
    assert sm_g is not None and sm_g.N == sm_c.N and mu_g.shape == mu_c.shape
    assert np.abs(mu_g - mu_c).max() < FIT_ATOL and abs(s_g - s_c) < FIT_ATOL
    assert np.abs(sm_g.transitions["lp"] - sm_c.transitions["lp"]).max() < FIT_ATOL
    x, ll = hm.viterbi(S, sm_g, mu_g, s_g)
    xo, llo = O.viterbi(S, sm_c, mu_c, s_c)
    assert np.array_equal(x, xo) and abs(ll - llo) <= LL_RTOL * abs(llo)
    Y = hm.reconstruct_signal(x, sm_g, mu_g, s_g)
    assert np.allclose(Y, O.reconstruct_signal(xo, sm_c, mu_c), atol=1e-6)
    assert 0.2 < 1 - np.std(Y - S) / np.std(S) < 0.7  # sanity only: from this seeded init the literal driver keeps one template
