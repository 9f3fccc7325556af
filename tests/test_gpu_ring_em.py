"""Parity of the fused ring E/M step against the CPU oracle's
forward -> backward -> update (src/baumwelch.jl:362-370), through the C ABI.
Bars (BASELINE.json north_star): log-likelihood within 1e-9 relative; fitted
mu / sigma / lA (the noise->head log-probabilities lp that define lA) within 1e-6
after N iterations from an explicit (mu0, sigma0, lp0)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIT_ATOL = 1e-6
LL_RTOL = 1e-9


def _start(hm, S, mu_true, N, K, p0=0.01):
    lA = hm.StateMatrix(N, K, np.log(np.full(N, p0)), False)
    mu0 = np.asfortranarray(0.7 * mu_true)
    return lA, mu0, float(np.std(S))


def _compare(r, o, tol=FIT_ATOL):
    lp, pp, mu, sig, ll = r[:5]
    lpo, ppo, muo, sigo, llo = o
    assert np.abs(lp - lpo).max() < tol, ("lp", lp, lpo)
    assert np.abs(mu - muo).max() < tol, ("mu", np.abs(mu - muo).max())
    assert abs(sig - sigo) < tol, ("sigma", sig, sigo)
    assert abs(ll - llo) <= LL_RTOL * abs(llo), ("loglik", ll, llo)
    fin = np.isfinite(ppo) & (ppo > -600)
    assert np.abs(pp[fin] - ppo[fin]).max() < 1e-6 * (1 + np.abs(ppo[fin]).max()), "pp"


@pytest.mark.parametrize("N,K,T,seed", [(3, 60, 30000, 3), (2, 10, 4000, 1), (3, 20, 9000, 2), (4, 48, 20000, 7),
                                        (1, 30, 6000, 4), (5, 60, 12000, 5), (7, 60, 8000, 9), (2, 4, 3000, 6),
                                        (3, 97, 8000, 8)])
def test_em_step_matches_oracle(hm, O, case_factory, N, K, T, seed):
    hm.set_ring_params(0, 0)
    S, lA_true, mu_true, sig = case_factory(N, K, T, seed, rate_scale=min(4.0, 60.0 / K))
    lA, mu0, s0 = _start(hm, S, mu_true, N, K)
    r = hm.em_step(S, lA, mu0, s0, mode="ring", return_info=True)
    assert r[5]["engine"] == 2
    o = O.em_step(S, O.OracleStateMatrix(N, K, np.log(np.full(N, 0.01)), False), mu0, s0)
    _compare(r, o, tol=1e-9)
    assert np.all(r[2][0, :] == 0.0)


@pytest.mark.parametrize("chunk,warm", [(1024, 256), (512, 512), (2048, 256)])
def test_em_step_many_chunks(hm, O, case_factory, chunk, warm):
    S, lA_true, mu_true, sig = case_factory(3, 60, 40000, 11)
    lA, mu0, s0 = _start(hm, S, mu_true, 3, 60)
    try:
        hm.set_ring_params(chunk, warm)
        r = hm.em_step(S, lA, mu0, s0, mode="ring", return_info=True)
    finally:
        hm.set_ring_params(0, 0)
    assert r[5]["n_chunks"] >= 40000 // max(chunk, 4 * 256) - 1
    o = O.em_step(S, O.OracleStateMatrix(3, 60, np.log(np.full(3, 0.01)), False), mu0, s0)
    _compare(r, o, tol=1e-9)


def test_em_true_model_and_mu_row1(hm, O, case_factory):
    """E/M step from the true model, sigma fixed, and with mu row 1 non-zero."""
    S, lA, mu, sig = case_factory(3, 60, 25000, 12)
    r = hm.em_step(S, lA, mu, sig, mode="ring")
    o = O.em_step(S, lA, mu, sig)
    _compare(r, o, tol=1e-9)
    mu2 = np.asfortranarray(mu.copy())
    mu2[0, :] = [0.02, -0.01, 0.03]
    r = hm.em_step(S, lA, mu2, sig, mode="ring")
    o = O.em_step(S, lA, mu2, sig)
    _compare(r, o, tol=1e-9)


def test_baum_welch_iterations(hm, O, case_factory):
    """5 E/M iterations through train_model's loop (device-resident X), compared with
    the oracle iterating forward/backward/update + the StateMatrix rebuild
    (src/baumwelch.jl:265,325-335).  mu is updated in place and the callback sees it
    (SURVEY D7)."""
    N, K, T = 3, 60, 30000
    S, lA_true, mu_true, sig = case_factory(N, K, T, 21)
    lA, mu0, s0 = _start(hm, S, mu_true, N, K)
    seen = []
    mu = mu0.copy(order="F")
    lA_fit, mu_fit, s_fit = hm.train_model(S, lA, mu, s0, 5, lambda m: seen.append(m.copy()))
    assert mu_fit is mu and len(seen) == 5 and np.array_equal(seen[0], mu0)
    smo = O.OracleStateMatrix(N, K, np.log(np.full(N, 0.01)), False)
    muo, so = mu0.copy(order="F"), s0
    for it in range(5):
        lpo, ppo, muo, so, llo = O.em_step(S, smo, muo, so)
        smo = O.OracleStateMatrix(N, K, lpo, False)
        if it < 4:
            assert np.abs(seen[it + 1] - muo).max() < FIT_ATOL
    assert np.abs(mu_fit - muo).max() < FIT_ATOL
    assert abs(s_fit - so) < FIT_ATOL
    lp_fit, _ = lA_fit.get_lp()
    # get_lp (src/types.jl:42-61) returns the noise -> head_i weights lp_i + (N-1) lpz of the rebuilt StateMatrix
    tro = smo.transitions
    assert np.abs(lp_fit - tro["lp"][(tro["src"] == 1) & (tro["dst"] > 1)]).max() < FIT_ATOL
    assert np.abs(lA_fit.transitions["lp"] - smo.transitions["lp"]).max() < FIT_ATOL
    assert abs(s_fit - 0.3) < 0.01  # converges to the truth (SURVEY appendix B probe)


@pytest.mark.parametrize("overlap", [False, True])
def test_train_loop_inside_the_library_equals_the_host_loop(hm, O, case_factory, overlap):
    """hmm_train_run: the E/M loop of src/baumwelch.jl:325-335 in one C call (no callback) -- each step's lp becomes the
    next step's transition weights inside the library.  It must walk the same sequence of models as the host loop
    (per-step call + StateMatrix rebuild), for a ring model and for an overlap model (generic E/M path)."""
    if overlap:
        N, K, T = 2, 12, 6000
        temps = np.stack([hm.create_spike_template(K, 3.0, 0.8, 0.2), hm.create_spike_template(K, 4.0, 0.3, 0.2)], 1)
        S = hm.create_signal(T, 0.3, np.array([0.01, 0.005]), temps, hm.make_rng(5))
        mu_true = np.asfortranarray(temps)
        lA = hm.StateMatrix(N, K, np.log(np.array([0.01, 0.01, 0.0001])), True)
    else:
        N, K, T = 3, 60, 30000
        S, _, mu_true, _ = case_factory(N, K, T, 21)
        lA = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), False)
    mu0, s0 = np.asfortranarray(0.7 * mu_true), float(np.std(S))
    steps = 4
    with hm.TrainContext(S) as ctx:
        lA_run, mu_run, s_run, ll_run, info = ctx.run(lA, mu0, s0, steps, return_info=True)
        lA_h, mu_h, s_h, ll_h = lA, mu0.copy(order="F"), s0, []
        for _ in range(steps):
            lp, pp, mu_h, s_h, ll = ctx.em_step(lA_h, mu_h, s_h)
            lA_h = hm.StateMatrix.from_states(lA_h.states, pp, K, lp, overlap)
            ll_h.append(ll)
    assert ll_run.size == steps and info["kernel_launches"] >= 4 * steps
    assert np.abs(mu_run - mu_h).max() < 1e-12 and abs(s_run - s_h) < 1e-12
    assert np.allclose(ll_run, ll_h, rtol=1e-13, atol=0)
    assert np.abs(lA_run.transitions["lp"] - lA_h.transitions["lp"]).max() < 1e-12
    assert np.array_equal(lA_run.transitions["src"], lA_h.transitions["src"])
    # and train_model without a callback takes that route
    mu = mu0.copy(order="F")
    lA_t, mu_t, s_t = hm.train_model(S, lA, mu, s0, steps)
    assert mu_t is mu and np.abs(mu - mu_h).max() < 1e-12 and abs(s_t - s_h) < 1e-12


def test_one_step_train_model_in_place(hm, O, case_factory):
    S, lA_true, mu_true, sig = case_factory(2, 30, 8000, 22, rate_scale=2.0)
    lA, mu0, s0 = _start(hm, S, mu_true, 2, 30)
    mu = mu0.copy(order="F")
    lA2, mu_out, s2 = hm.train_model(S, lA, mu, s0)
    assert mu_out is mu and not np.array_equal(mu, mu0)
    o = O.em_step(S, O.OracleStateMatrix(2, 30, np.log(np.full(2, 0.01)), False), mu0, s0)
    assert np.abs(mu - o[2]).max() < 1e-9 and abs(s2 - o[3]) < 1e-9
    smo = O.OracleStateMatrix(2, 30, o[0], False)
    assert np.abs(lA2.transitions["lp"] - smo.transitions["lp"]).max() < 1e-9
    assert np.allclose(lA2.pi[np.isfinite(o[1])], o[1][np.isfinite(o[1])], atol=1e-6)


def test_em_silent_neuron_and_high_snr(hm, O, case_factory, monkeypatch):
    """Two numerically awkward corners of the E/M step:
    (a) a neuron whose prior is 2^-90 (the reference's initial lp, src/baumwelch.jl:311): its statistics are sums of
        terms around e^-100 whose RELATIVE accuracy sets the new lp -- nothing representable may be dropped;
    (b) sigma ten times smaller than the data's (a spike gains thousands of nats): a single 32-step window spans
        more than e^600 and the live windows fall back from the linear to the log domain.
    Both against the oracle, and the linear-domain live windows against the log-domain ones (HMMCUDA_EM_DBG=1)."""
    N, K, T = 3, 60, 24000
    S, lA_true, mu_true, sig = case_factory(N, K, T, 21, rate_scale=2.0)
    mu0 = np.asfortranarray(0.7 * mu_true)
    # (a)
    lp0 = np.log(np.array([0.01, 2.0 ** -90, 0.003]))
    lA = hm.StateMatrix(N, K, lp0, False)
    r = hm.em_step(S, lA, mu0.copy(order="F"), float(np.std(S)), mode="ring", return_info=True)
    o = O.em_step(S, O.OracleStateMatrix(N, K, lp0, False), mu0.copy(order="F"), float(np.std(S)))
    _compare(r, o, tol=1e-8)
    assert r[0][1] < -15  # the silent neuron stays rare (p < 1e-6), and its lp is still accurate to 1e-8
    # (b)
    lA = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), False)
    s_small = 0.1 * sig
    r_lin = hm.em_step(S, lA, mu_true.copy(order="F"), s_small, mode="ring", return_info=True)
    o = O.em_step(S, O.OracleStateMatrix(N, K, np.log(np.full(N, 0.01)), False), mu_true.copy(order="F"), s_small)
    _compare(r_lin, o, tol=1e-8)
    monkeypatch.setenv("HMMCUDA_EM_DBG", "1")  # log-domain live windows everywhere
    r_log = hm.em_step(S, lA, mu_true.copy(order="F"), s_small, mode="ring", return_info=True)
    monkeypatch.delenv("HMMCUDA_EM_DBG")
    _compare(r_log, o, tol=1e-8)
    assert abs(r_lin[4] - r_log[4]) <= 1e-12 * abs(r_log[4])
    assert np.abs(r_lin[2] - r_log[2]).max() < 1e-10


def test_neuron_whose_mass_sits_at_the_end_of_the_recording(hm, O):
    """tests/golden/em_end_mass_case.npz (N=5, K=80, T=22 477; found by tools/fuzz_parity.py, regenerated by
    tests/golden/make_end_mass_case.py): a nearly silent neuron (0.17 expected spikes) whose posterior mass sits in a
    spike cut off by the end of the recording.  The per-phase occupancies S0[i][s] of its late phases are sums over
    very few entry times; they must be accumulated as sums of positive terms (chains that fit + the last L-1 entry
    times per phase), never as `total - those that do not reach the phase`, which cancels catastrophically here
    (mu was off by 1.5).  Checked against the stored oracle step and against the oracle run now."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "em_end_mass_case.npz"))
    N, K = int(g["N"]), int(g["K"])
    lA = hm.StateMatrix(N, K, g["lp0"], False)
    r = hm.em_step(g["S"], lA, np.asfortranarray(g["mu0"]).copy(order="F"), float(g["sigma0"]), mode="ring")
    o = O.em_step(g["S"], lA, np.asfortranarray(g["mu0"]).copy(order="F"), float(g["sigma0"]))
    assert np.array_equal(o[0], g["lp"]) and np.array_equal(o[2], g["mu"])  # the oracle still says what the fixture says
    tol = 1e-9 + 1e-11 * np.exp(-o[0])  # statistics of a neuron with n expected spikes carry 1e-12 T / n (see tools/fuzz_parity.py)
    assert np.all(np.abs(r[0] - o[0]) <= tol), (r[0], o[0])
    assert np.all(np.abs(r[2] - o[2]).max(axis=0) <= tol), np.abs(r[2] - o[2]).max(axis=0)
    assert abs(r[3] - o[3]) < 1e-9 and abs(r[4] - o[4]) <= LL_RTOL * abs(o[4])


def test_high_snr_boundary_is_not_accepted_on_its_large_entries_alone(hm, O):
    """tests/golden/em_high_snr_boundary_case.npz (N=6, K=81, T=45 482, template amplitudes of 12 sigma; found by
    tools/fuzz_parity.py seed 5): a chunk boundary falls inside a spike whose pending chain entry is more than e^745
    above the noise score.  The boundary check used to skip entries that far below the maximum and accepted a
    speculative start whose noise score had not converged -- log-likelihood off by 1.2e-3, sigma NaN.  Every finite
    entry is compared now: the chunk is repaired and the step is the oracle's."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "em_high_snr_boundary_case.npz"))
    N, K = int(g["N"]), int(g["K"])
    lA = hm.StateMatrix(N, K, g["lp0"], False)
    out = hm.em_step(g["S"], lA, np.asfortranarray(g["mu0"]).copy(order="F"), float(g["sigma0"]), mode="ring", return_info=True)
    r, info = out[:5], out[5]
    assert info["fwd_repaired"] + info["bwd_repaired"] > 0, info  # the default chunking does hit the case
    assert np.abs(r[0] - g["lp"]).max() < 1e-9 and np.abs(r[2] - g["mu"]).max() < 1e-9
    assert abs(r[3] - float(g["sigma"])) < 1e-9 and abs(r[4] - float(g["loglik"])) <= LL_RTOL * abs(float(g["loglik"]))
