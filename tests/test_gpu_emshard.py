"""Time-sharded Baum-Welch (hmm_emshard_*): one recording cut into time shards with ghost chunks; the
sufficient statistics of the shards' main spans are summed, the boundary vectors of neighbours are compared.  Here all
shards run on one GPU (on a multi-GPU box the same code runs one shard per rank over NCCL: bench.py --workload c3).
Fits must agree with the CPU oracle's whole-recording E/M step to 1e-6 (north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FIT_ATOL = 1e-6


def _setup(hm, case_factory, N, K, T, seed, n, chunk_len):
    import torch

    ts = hm.timeshard
    S, lA_true, mu_true, sig = case_factory(N, K, T, seed, rate_scale=2.0)
    dev = torch.device("cuda", 0)
    spans = ts.shard_plan(T, n, chunk_len, 256)
    xs = [torch.from_numpy(np.ascontiguousarray(S[sp[0]:sp[1]])).to(dev) for sp in spans]
    shards = [ts.EmShard(x.data_ptr(), False, sp, T, chunk_len) for x, sp in zip(xs, spans)]
    lA = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), False)
    em = ts.EmSharded(shards, N, K, lA.nstates, dev)
    return S, lA, np.asfortranarray(0.7 * mu_true), float(np.std(S)), em, xs


@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_em_step_matches_oracle(hm, O, case_factory, n):
    N, K, T = 3, 60, 160_000
    S, lA, mu0, s0, em, keep = _setup(hm, case_factory, N, K, T, 401, n, 4096)
    try:
        lp, pp, mu, sig, ll = em.em_step(lA, mu0, s0)
    finally:
        em.close()
    o = O.em_step(S, O.OracleStateMatrix(N, K, np.log(np.full(N, 0.01)), False), mu0.copy(order="F"), s0)
    assert em.last_check is True
    assert np.abs(lp - o[0]).max() < FIT_ATOL and np.abs(mu - o[2]).max() < FIT_ATOL and abs(sig - o[3]) < FIT_ATOL
    assert abs(ll - o[4]) <= 1e-9 * abs(o[4]), (ll, o[4])
    fin = np.isfinite(o[1])
    assert np.array_equal(np.isfinite(pp), fin) and np.abs(pp[fin] - o[1][fin]).max() < FIT_ATOL


def test_sharded_iterations_follow_the_single_gpu_fit(hm, O, case_factory):
    """Five iterations, N=4 / K=48, 5 shards: every iteration against the single-GPU fused step and the last one
    against the oracle."""
    N, K, T = 4, 48, 120_000
    S, lA, mu, sig, em, keep = _setup(hm, case_factory, N, K, T, 402, 5, 4096)
    lA1, mu1, sig1 = lA, mu.copy(order="F"), sig
    try:
        for it in range(5):
            lp, pp, mu, sig, ll = em.em_step(lA, mu, sig)
            lA = hm.StateMatrix.from_states(lA.states, pp, K, lp, False)
            r = hm.em_step(S, lA1, mu1, sig1)
            lA1, mu1, sig1 = hm.StateMatrix.from_states(lA1.states, r[1], K, r[0], False), r[2], r[3]
            assert np.abs(mu - mu1).max() < FIT_ATOL and abs(sig - sig1) < FIT_ATOL and abs(ll - r[4]) <= 1e-9 * abs(r[4])
    finally:
        em.close()


def test_last_shard_of_one_partial_chunk(hm, O, case_factory):
    """T = 12 chunks + 1000 samples, 7 shards: the last main span ends in a partial chunk, so the shard before the last
    has a right ghost shorter than a chunk -- complete all the same, because it ends where the recording ends
    (hmm_emshard_create used to refuse it; tools/fuzz_parity.py ran into that)."""
    N, K, cl = 3, 40, 4096
    T = 12 * cl + 1000
    S, lA, mu0, s0, em, keep = _setup(hm, case_factory, N, K, T, 403, 7, cl)  # 2 chunks per shard, the last one: the partial chunk
    try:
        lp, pp, mu, sig, ll = em.em_step(lA, mu0, s0)
    finally:
        em.close()
    o = O.em_step(S, O.OracleStateMatrix(N, K, np.log(np.full(N, 0.01)), False), mu0.copy(order="F"), s0)
    assert em.last_check is True
    assert np.abs(lp - o[0]).max() < FIT_ATOL and np.abs(mu - o[2]).max() < FIT_ATOL and abs(sig - o[3]) < FIT_ATOL
    assert abs(ll - o[4]) <= 1e-9 * abs(o[4]), (ll, o[4])
