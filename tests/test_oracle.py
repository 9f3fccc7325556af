"""Pins the CPU oracle (oracle/hmm_oracle.c) to everything the reference's own
tests fix for this path, plus brute force, an independent dense formulation and
invariants.  No GPU needed."""
import itertools

import numpy as np
import pytest

LOG2PI = 0.9189385332046727


def funcl(x, mu, sigma):
    return -LOG2PI - np.log(sigma) - (x - mu) * (x - mu) / (2 * sigma * sigma)


def test_unroll_known_answer(hm, O):
    """test/runtests.jl:36-42 -- exact state layout of generate_states (overlap model)."""
    for ctor in (O.OracleStateMatrix, hm.StateMatrix):
        sm = ctor(2, 5, np.log([0.01, 0.004]))
        assert sm.nstates == 25 and sm.transitions.size == 36
        mlseq = np.array([1, 1, 1, 2, 3, 4, 5, 1, 6, 7, 8, 9, 1, 10, 15, 20, 25, 1], dtype=np.int16)
        u = O.unroll_mlseq(mlseq, sm)
        assert u[0].tolist() == [1, 1, 1, 2, 3, 4, 5, 1, 1, 1, 1, 1, 1, 2, 3, 4, 5, 1]
        assert u[1].tolist() == [1, 1, 1, 1, 1, 1, 1, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 1]


def test_template_golden(hm):
    """test/runtests.jl:49-55 -- find_best_overlap of temp1 against itself shifted by 4
    is sum(temp1[1:56].^2) ~ 100.66411692920131; pins create_spike_template."""
    t1 = hm.create_spike_template(60, 3.0, 0.8, 0.2)
    assert t1[0] == 0.0
    assert np.isclose(float((t1[:56] ** 2).sum()), 100.66411692920131, rtol=1e-12)


@pytest.mark.parametrize("N,K,ov", [(3, 60, False), (4, 48, False), (5, 60, False), (7, 60, False), (2, 5, True),
                                    (3, 6, True), (2, 60, True)])
def test_statematrix_counts_and_mirror(hm, O, N, K, ov):
    """SURVEY appendix B probe counts; host mirror == oracle restatement, bit for bit."""
    lp = np.log(np.linspace(0.001, 0.004, N))
    a, b = O.OracleStateMatrix(N, K, lp, ov), hm.StateMatrix(N, K, lp, ov)
    assert np.array_equal(a.states, b.states)
    assert a.transitions.tobytes() == b.transitions.tobytes()
    expect = {(3, 60, False): (178, 187), (4, 48, False): (189, 205), (5, 60, False): (296, 321),
              (7, 60, False): (414, 463), (2, 5, True): (25, 36), (3, 6, True): (91, 157), (2, 60, True): (3600, 3721)}
    assert (a.nstates, a.transitions.size) == expect[(N, K, ov)]
    src, dst = a.transitions["src"], a.transitions["dst"]
    assert np.all(np.diff(src * (a.nstates + 1) + dst) > 0)  # sorted by (src, dst)


def test_ring_weights(O):
    """the non-overlap transition weights, incl. the
    product-not-sum quirk lpz = log(1 - prod p_i) (src/types.jl:96)."""
    N, K = 3, 6
    p = np.array([0.01, 0.02, 0.03])
    sm = O.OracleStateMatrix(N, K, np.log(p), False)
    W = {(int(r["src"]), int(r["dst"])): float(r["lp"]) for r in sm.transitions}
    lpz = np.log1p(-np.exp(np.log(p).sum()))
    assert np.isclose(lpz, np.log(1 - p.prod()))
    L = K - 1
    head = lambda i: 2 + i * L
    tail = lambda i: 1 + (i + 1) * L
    assert W[(1, 1)] == (0.0 + lpz + lpz) + lpz
    for i in range(N):
        assert np.isclose(W[(1, head(i))], np.log(p[i]) + 2 * lpz, rtol=1e-15)
        assert np.isclose(W[(tail(i), 1)], 2 * lpz)
        for s in range(L - 1):
            assert np.isclose(W[(head(i) + s, head(i) + s + 1)], 2 * lpz)
        for j in range(N):
            if i != j:
                assert np.isclose(W[(tail(j), head(i))], np.log(p[i]) + lpz)
        assert (tail(i), head(i)) not in W
    assert sorted(d for (s, d) in W if s == 1) == [1, 2, 7, 12]


def _tiny(hm, N=2, K=3, T=6, seed=0):
    rng = np.random.default_rng(seed)
    mu = np.asfortranarray(rng.normal(size=(K, N)))
    mu[0, :] = 0.0
    y = rng.normal(size=T)
    lp = np.log(rng.uniform(0.05, 0.3, size=N))
    return y, mu, lp


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("ov", [False, True])
def test_viterbi_bruteforce(hm, O, seed, ov):
    """Enumerate every state path of a tiny model: the oracle's path must attain the
    maximum of init + sum(lp + emission) under src/viterbi.jl:55-63,74-87 semantics,
    and ll must be the sum of the cumulative scores along it (:92-96)."""
    N, K, T = 2, 3, 6
    y, mu, lp = _tiny(hm, N, K, T, seed)
    sm = O.OracleStateMatrix(N, K, lp, ov)
    sig = 0.7
    W = {(int(r["src"]) - 1, int(r["dst"]) - 1): float(r["lp"]) for r in sm.transitions}
    m = np.array([sum(mu[sm.states[l, j] - 1, l] for l in range(N)) for j in range(sm.nstates)])
    q = np.array([[funcl(y[t], m[j], sig) for j in range(sm.nstates)] for t in range(T)])
    best, best_path = -np.inf, None
    for path in itertools.product(range(sm.nstates), repeat=T):
        s = 0.0 if path[0] == 0 else q[0, path[0]]
        ok = True
        for t in range(1, T):
            w = W.get((path[t - 1], path[t]))
            if w is None:
                ok = False
                break
            s += w + q[t, path[t]]
        if ok and s > best:
            best, best_path = s, path
    x, ll, T2, T1 = O.viterbi(y, sm, mu, sig, trellis=True)
    xs = x.astype(int) - 1
    cum = [0.0 if xs[0] == 0 else q[0, xs[0]]]
    for t in range(1, T):
        cum.append(cum[-1] + W[(xs[t - 1], xs[t])] + q[t, xs[t]])
    assert np.isclose(cum[-1], best, rtol=1e-12)
    assert tuple(xs) == best_path
    assert np.isclose(ll, sum(cum[1:]), rtol=1e-12)
    assert np.isclose(T1[xs[-1], -1], best, rtol=1e-12)
    # streaming (block-recompute) form == dense form, bit for bit
    x2, ll2 = O.viterbi(y, sm, mu, sig)
    assert np.array_equal(x, x2) and ll == ll2


def test_viterbi_blockwise_equals_dense(hm, O, case_factory):
    """The oracle's checkpoint/recompute variant is bit-identical to the dense-trellis
    form of src/viterbi.jl:52-53 across block boundaries (block = 4096)."""
    S, lA, mu, sig = case_factory(2, 20, 9000, 11)
    x, ll, T2, T1 = O.viterbi(S, lA, mu, sig, trellis=True)
    x2, ll2 = O.viterbi(S, lA, mu, sig)
    assert np.array_equal(x, x2) and ll == ll2
    # backtrack the dense trellis by hand
    xb = np.empty_like(x)
    xb[-1] = np.argmax(T1[:, -1]) + 1
    acc = 0.0
    for i in range(S.size - 1, 0, -1):
        xb[i - 1] = T2[xb[i] - 1, i]
        acc += T1[xb[i] - 1, i]
    assert np.array_equal(xb, x) and acc == ll
    assert T1[0, 0] == 0.0 and np.all(T2[:, 0] == 1)


def test_forward_backward_invariants(hm, O, case_factory):
    """LSE_j(alpha+beta) is the same at every t; sum_j exp(gamma) = 1."""
    S, lA, mu, sig = case_factory(2, 12, 1500, 5, rate_scale=4.0)
    a, b = O.forward(S, lA, mu, sig), O.backward(S, lA, mu, sig)
    g = np.logaddexp.reduce(a + b, axis=0)
    assert np.ptp(g) < 1e-9 * abs(g[0])
    assert np.isclose(O.loglik(a), g[-1], rtol=1e-13)
    lp, pp, mu2, s2, gam = O.update(a, b, lA, mu, sig, S, want_gamma=True)
    assert np.allclose(np.exp(gam).sum(axis=0), 1.0, atol=1e-12)
    assert np.array_equal(pp, gam[:, 0])
    assert np.all(mu2[0, :] == 0.0) and s2 > 0


def test_forward_dense_formulation(hm, O):
    """N=1: the StateMatrix forward (src/baumwelch.jl:25-51) equals the legacy dense
    single-ring forward (src/baumwelch.jl:1-23 with prepA :376-385) when the latter
    is given a zero log-prior -- an independent second formulation."""
    K, T, p, sig = 8, 300, 0.05, 0.5
    rng = np.random.default_rng(3)
    temp = hm.create_spike_template(K, 2.0, 0.8, 0.2)
    V = hm.create_signal(T, sig, [p], temp[:, None], hm.make_rng(4))
    sm = O.OracleStateMatrix(1, K, np.log([p]), False)
    mu = np.asfortranarray(temp[:, None].copy())
    mu[0, 0] = 0.0
    a = O.forward(V, sm, mu, sig)
    n = K
    lA = np.full((n, n), -np.inf)
    lA[0, 0] = np.log(1 - p)
    lA[0, 1] = np.log(p)
    for i in range(1, n - 1):
        lA[i, i + 1] = 0.0
    lA[n - 1, 0] = 0.0
    m = mu[:, 0]
    d = np.zeros((n, T))
    d[:, 0] = funcl(V[0], m, sig)
    for i in range(1, T):
        for j in range(n):
            aa = -np.inf
            for k in range(n):
                if np.isfinite(lA[k, j]):
                    aa = np.logaddexp(aa, d[k, i - 1] + lA[k, j])
            d[j, i] = funcl(V[i], m[j], sig) + aa
    assert np.allclose(a, d, rtol=1e-12, atol=1e-12)


def test_em_monotone_and_converges(hm, O, case_factory):
    """Baum-Welch sanity (SURVEY appendix B probe): log-likelihood is non-decreasing
    and sigma / templates approach the truth from mu0 = 0.7*truth."""
    S, lA_true, mu_true, sig = case_factory(2, 30, 6000, 8, rate_scale=2.0)
    N, K = 2, 30
    sm = O.OracleStateMatrix(N, K, np.log(np.full(N, 0.01)), False)
    mu = np.asfortranarray(0.7 * mu_true)
    s = float(np.std(S))
    lls = []
    for _ in range(4):
        lp, pp, mu, s, ll = O.em_step(S, sm, mu, s)
        sm = O.OracleStateMatrix(N, K, lp, False)
        lls.append(ll)
    assert all(b >= a - 1e-6 for a, b in zip(lls, lls[1:]))
    assert abs(s - sig) < 0.02
    assert np.abs(mu - mu_true).max() < 0.25


def test_reconstruct_and_chunked(hm, O, case_factory):
    S, lA, mu, sig = case_factory(3, 20, 12000, 2)
    x, ll = O.viterbi(S, lA, mu, sig)
    Y = O.reconstruct_signal(x, lA, mu)
    m = np.array([sum(mu[lA.states[l, j] - 1, l] for l in range(lA.N)) for j in range(lA.nstates)])
    assert np.array_equal(Y, m[x - 1])
    score = 1 - np.std(Y - S) / np.std(S)
    assert 0.3 < score < 0.8  # test/runtests.jl:31-34 analogue (statistical window)
    ml, llc = O.fit_chunked(S, lA, mu, sig, 3000)
    assert ml.size == S.size and (ml != x).mean() < 0.05
    with pytest.raises(ValueError):
        O.reconstruct_signal(np.array([0], dtype=np.int16), lA, mu)


@pytest.mark.parametrize("N,K,T,seed", [(2, 5, 70, 31), (1, 6, 90, 32), (3, 4, 50, 33)])
def test_em_step_against_50_digit_arithmetic(hm, O, N, K, T, seed):
    """An independent statement of src/baumwelch.jl:25-51, 73-98, 205-309 in the LINEAR domain with mpmath at 50
    digits (plain sum-product, no log-sum-exp, no operation-order assumptions): alpha, beta, gamma, xi, then the
    M-step formulas.  The oracle's one-step train_model must agree to ~1e-12 -- this pins the oracle's E/M
    arithmetic beyond its own invariants."""
    import mpmath as mp

    mp.mp.dps = 50
    temps = np.stack([hm.create_spike_template(K, 2.0 + i, 0.7 - 0.1 * i, 0.25) for i in range(N)], axis=1)
    p = np.array([0.05, 0.03, 0.04][:N])
    S = hm.create_signal(T, 0.4, p, temps, hm.make_rng(seed))
    sm = O.OracleStateMatrix(N, K, np.log(np.full(N, 0.04)), False)
    mu0 = np.asfortranarray(0.8 * temps)
    mu0[0, :] = 0.0
    s0 = 0.5
    lp, pp, mu1, s1, llk = O.em_step(S, sm, mu0.copy(order="F"), s0)

    ns, states, tr = sm.nstates, np.asarray(sm.states), sm.transitions
    m = [mp.fsum(mp.mpf(float(mu0[states[l, j] - 1, l])) for l in range(N)) for j in range(ns)]
    sig = mp.mpf(s0)
    y = [mp.mpf(float(v)) for v in S]

    def b(j, t):  # funcl, src/utils.jl:3-4
        d = y[t] - m[j]
        return mp.exp(-mp.log(2 * mp.pi) / 2 - mp.log(sig) - d * d / (2 * sig * sig))

    edges = [(int(e["src"]) - 1, int(e["dst"]) - 1, mp.exp(mp.mpf(float(e["lp"])))) for e in tr]
    al = [[mp.mpf(0)] * ns for _ in range(T)]
    be = [[mp.mpf(0)] * ns for _ in range(T)]
    for j in range(ns):
        al[0][j] = b(j, 0)  # no prior, noise not forced (src/baumwelch.jl:30-37)
        be[T - 1][j] = mp.mpf(1)
    for t in range(1, T):
        for (k, j, a) in edges:
            al[t][j] += al[t - 1][k] * a * b(j, t)
    for t in range(T - 2, -1, -1):
        for (k, j, a) in edges:
            be[t][k] += a * b(j, t + 1) * be[t + 1][j]
    Z = mp.fsum(al[T - 1])
    assert abs(mp.log(Z) - llk) <= 1e-12 * abs(llk)
    gam = [[al[t][j] * be[t][j] / Z for j in range(ns)] for t in range(T)]
    # transition posteriors out of the noise state, list order (src/baumwelch.jl:226-253)
    from_noise = [(j, a) for (k, j, a) in edges if k == 0]
    xx = [mp.fsum(al[t][0] * a * b(j, t + 1) * be[t + 1][j] / Z for t in range(T - 1)) for (j, a) in from_noise]
    bb = mp.fsum(gam[t][0] for t in range(T - 1))
    lp_ref = [mp.log(x / bb) for x in xx][1:]
    assert max(abs(float(a2) - b2) for a2, b2 in zip(lp_ref, lp)) < 1e-11
    # templates: states with exactly one active neuron (src/baumwelch.jl:266-287)
    mu_ref = np.zeros_like(mu0)
    for l in range(N):
        for s_ in range(2, K + 1):
            js = [j for j in range(ns) if states[l, j] == s_ and all(states[q, j] == 1 for q in range(N) if q != l)]
            num = mp.fsum(y[t] * gam[t][j] for t in range(T) for j in js)
            den = mp.fsum(gam[t][j] for t in range(T) for j in js)
            mu_ref[s_ - 1, l] = float(num / den)
    assert np.abs(mu_ref - mu1).max() < 1e-11
    # sigma with the NEW means over all states (src/baumwelch.jl:288-307)
    m1 = [mp.fsum(mp.mpf(float(mu_ref[states[l, j] - 1, l])) for l in range(N)) for j in range(ns)]
    num = mp.fsum((y[t] - m1[j]) ** 2 * gam[t][j] for t in range(T) for j in range(ns))
    den = mp.fsum(gam[t][j] for t in range(T) for j in range(ns))
    assert abs(float(mp.sqrt(num / den)) - s1) < 1e-12
    assert max(abs(float(mp.log(gam[0][j])) - pp[j]) for j in range(ns) if gam[0][j] > mp.mpf("1e-250")) < 1e-9


def _isvalid_transition_literal(states0, K, lp, j1, j2):
    """src/types.jl:94-113, line for line (0-based j1/j2): sum(lp) runs over the WHOLE vector."""
    lpt = 0.0
    s = 0.0
    for v in lp:
        s += float(v)
    lpz = float(np.log1p(-np.exp(s)))
    for i in range(states0.shape[0]):
        s1, s2 = int(states0[i, j1]), int(states0[i, j2])
        if s1 == 0 and s2 == 0:
            lpt += lpz
        elif s1 == 0 and s2 == 1:
            lpt += float(lp[i])
        elif (s2 - s1 == 1) or (s1 == K - 1 and s2 == 0):
            lpt += 0.0
        else:
            return -np.inf
    return lpt


def test_overlap_rebuild_uses_the_full_lp_vector(hm, O):
    """update() rebuilds an overlap StateMatrix from xb[2:end], which has one entry per transition out of the
    silent state (N + N(N-1)/2 = 3 for N = 2), not N (src/baumwelch.jl:226,264-265); sum(lp) at
    src/types.jl:96 covers all of them.  Both mirrors must follow the literal restatement."""
    N, K = 2, 4
    base = hm.StateMatrix(N, K, np.log([0.01, 0.02]), True)
    assert int((base.transitions["src"] == 1).sum()) - 1 == 3
    lp3 = np.log([0.011, 0.019, 0.0004])
    s0 = np.asarray(base.states, dtype=np.int16) - 1
    expect = []
    for i in range(base.nstates):
        for j in range(base.nstates):
            a = _isvalid_transition_literal(s0, K, lp3, i, j)
            if np.isfinite(a):
                expect.append((i + 1, j + 1, a))
    pp = np.full(base.nstates, -np.log(base.nstates))
    got_h = hm.StateMatrix.from_states(base.states, pp, K, lp3, True).transitions
    got_o = O.OracleStateMatrix(N, K, lp3, True, states0=np.asfortranarray(s0)).transitions
    for got in (got_h, got_o):
        assert got.size == len(expect)
        assert [(int(r["src"]), int(r["dst"])) for r in got] == [(a, b) for a, b, _ in expect]
        assert np.array_equal(got["lp"], np.array([w for _, _, w in expect]))
    # the truncated sum (first N entries only) gives different noise->noise weights: the test can tell
    trunc = np.log1p(-np.exp(lp3[:N].sum()))
    assert got_h["lp"][0] != N * trunc


@pytest.mark.parametrize("N,K,overlap", [(1, 2, False), (1, 5, False), (3, 12, False), (7, 9, False), (2, 2, True), (2, 7, True),
                                         (3, 5, True), (4, 4, True), (3, 2, True), (2, 24, True)])
def test_transition_enumeration_equals_the_all_pairs_scan(hm, O, N, K, overlap):
    """The host mirror finds the finite transitions of a layout by enumerating each state's admissible successors
    instead of testing all pairs (src/types.jl:114-127 is O(nstates^2 N): 45 s in numpy for the CLI's 21 123-state
    model).  Same records, same order, as the literal scan and as the oracle's constructor."""
    sm = hm.statematrix
    st0 = sm.generate_states(N, K, overlap)
    a, b = sm._enumerate_topology(st0, K), sm._scan_topology(st0, K)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    lp = np.log(np.linspace(0.002, 0.01, N))
    tr = hm.StateMatrix(N, K, lp, overlap).transitions
    tro = O.OracleStateMatrix(N, K, lp, overlap).transitions
    assert np.array_equal(tr["src"], tro["src"]) and np.array_equal(tr["dst"], tro["dst"])
    assert np.allclose(tr["lp"], tro["lp"], rtol=1e-14, atol=0)
