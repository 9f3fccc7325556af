"""Parity of the time-parallel per-state engine (generic_parallel.cu: any StateMatrix -- overlap models, N > 7) against
the CPU oracle, through the C ABI.  Bars: x bit-exact; ll within 1e-9 relative (the path score is re-summed in parallel;
mode="faithful" keeps the reference's own rounding of ll)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
LL_RTOL = 1e-9


def _overlap_case(hm, K, T, seed, pp=(0.003, 0.001)):
    temps = np.stack([hm.create_spike_template(K, 3.0, 0.8, 0.2), hm.create_spike_template(K, 4.0, 0.3, 0.2)], 1)
    pp = np.array(pp)
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(seed))
    return S, hm.StateMatrix(2, K, np.log(pp), True), np.asfortranarray(temps), 0.3


def _check(hm, O, S, lA, mu, sig, mode="generic"):
    x, ll, info = hm.viterbi(S, lA, mu, sig, mode=mode, return_info=True)
    xo, llo = O.viterbi(S, lA, mu, sig)
    assert info["engine"] == 3, info
    assert x.dtype == np.int16 and np.array_equal(x, xo), int(np.sum(x != xo))
    assert abs(ll - llo) <= LL_RTOL * abs(llo)
    return info


def test_reference_testset_overlap_model_time_parallel(hm, O):
    """test/runtests.jl:17-34 -- two K=60 templates, allow_overlaps=true (3 600 states), T=20 000: the automatic engine
    for a model the ring engine cannot take is now the time-parallel one."""
    S, lA, mu, sig = _overlap_case(hm, 60, 20000, 1234)
    assert lA.nstates == 3600
    info = _check(hm, O, S, lA, mu, sig, mode="auto")
    assert info["n_chunks"] >= 8 and info["kernel_launches"] >= 5


def test_overlap_model_many_chunks_and_real_repairs(hm, O):
    """A warm-up far shorter than a template (8 samples) makes most speculative starts wrong: the verification must catch
    every one of them and the repaired decode must be the oracle's."""
    S, lA, mu, sig = _overlap_case(hm, 24, 150000, 5, pp=(0.01, 0.006))
    try:
        hm.set_ring_params(512, 8)
        info = _check(hm, O, S, lA, mu, sig)
    finally:
        hm.set_ring_params(0, 0)
    assert info["n_chunks"] >= 290
    assert info["fwd_repaired"] > 0 and info["bwd_repaired"] > 0, info


def test_forced_flags_alternate_repaired_and_accepted_chunks(hm, O, monkeypatch):
    S, lA, mu, sig = _overlap_case(hm, 30, 120000, 9)
    monkeypatch.setenv("HMMCUDA_DEBUG_FLAG_EVERY", "3")
    try:
        hm.set_ring_params(2048, 256)
        info = _check(hm, O, S, lA, mu, sig)
    finally:
        hm.set_ring_params(0, 0)
    assert info["fwd_repaired"] >= info["n_chunks"] // 3 - 1 and info["bwd_repaired"] > 0, info


@pytest.mark.parametrize("T", [4096, 4097, 65536 + 1, 100003])
def test_ragged_lengths(hm, O, T):
    S, lA, mu, sig = _overlap_case(hm, 16, T, T % 97)
    try:
        hm.set_ring_params(1000, 100)  # chunk length not a power of two, last chunk ragged
        _check(hm, O, S, lA, mu, sig)
    finally:
        hm.set_ring_params(0, 0)


def test_ring_model_through_the_generic_engine(hm, O, case_factory):
    S, lA, mu, sig = case_factory(3, 60, 200000, 71)
    info = _check(hm, O, S, lA, mu, sig)
    assert info["fwd_repaired"] == 0 and info["bwd_repaired"] == 0, info


def test_eight_neurons_is_outside_the_ring_engine_and_goes_generic(hm, O, case_factory):
    """N = 8 > RING_MAX_N: the automatic choice is the time-parallel per-state engine."""
    S, lA, mu, sig = case_factory(8, 30, 60000, 13)
    _check(hm, O, S, lA, mu, sig, mode="auto")


def test_batch_of_channels(hm, O):
    cases = [_overlap_case(hm, 20, 30000, 40 + c) for c in range(3)]
    Y = np.asfortranarray(np.stack([c[0] for c in cases], axis=1))
    lA = cases[0][1]
    mus = [c[2] for c in cases]
    x, ll, info = hm.viterbi_batch(Y, [(lA, mus[c], 0.3) for c in range(3)], mode="generic", return_info=True)
    assert info["engine"] == 3
    for c in range(3):
        xo, llo = O.viterbi(cases[c][0], lA, mus[c], 0.3)
        assert np.array_equal(x[:, c], xo) and abs(ll[c] - llo) <= LL_RTOL * abs(llo)


def test_short_sequences_one_chunk_in_generic_mode_and_sequential_engine_in_auto(hm, O):
    S, lA, mu, sig = _overlap_case(hm, 12, 3000, 3)
    for T in (1, 2, 3000):
        _check(hm, O, S[:T], lA, mu, sig)
    x, ll, info = hm.viterbi(S, lA, mu, sig, return_info=True)
    xo, llo = O.viterbi(S, lA, mu, sig)
    assert info["engine"] == 1 and np.array_equal(x, xo) and ll == llo


@pytest.mark.parametrize("kb,direct", [(150, 0), (70, 1), (40, 0), (4, 1)])
def test_every_shared_memory_placement(hm, O, monkeypatch, kb, direct):
    """Models of 10 000+ states keep part of their tables / the score columns in global memory.  A reduced budget
    (HMMCUDA_DEBUG_GEN_SMEM_KB) walks the 3 600-state reference test model through every placement -- columns in shared
    memory with the tables in L2, everything in L2 -- and HMMCUDA_DEBUG_GEN_DIRECT_TRACE through the traceback that
    follows backpointers straight from global memory; forced flags keep the repair paths in play."""
    S, lA, mu, sig = _overlap_case(hm, 60, 30000, 77)
    monkeypatch.setenv("HMMCUDA_DEBUG_GEN_SMEM_KB", str(kb))
    monkeypatch.setenv("HMMCUDA_DEBUG_GEN_DIRECT_TRACE", str(direct))
    monkeypatch.setenv("HMMCUDA_DEBUG_FLAG_EVERY", "4")
    info = _check(hm, O, S, lA, mu, sig)
    assert info["fwd_repaired"] > 0 and info["bwd_repaired"] > 0, info


def test_cli_sized_overlap_model_three_templates(hm, O):
    """src/hmmsort.jl:54 builds an overlap model from up to four templates: three K=60 templates are 10 621 states --
    beyond what either per-state engine could hold in shared memory before.  Auto mode decodes it on the
    time-parallel engine (score columns in shared memory, tables in L2) at any length."""
    K, T = 60, 9000
    temps = np.stack([hm.create_spike_template(K, 3.0, 0.8, 0.2), hm.create_spike_template(K, 4.0, 0.3, 0.2),
                      hm.create_spike_template(K, 2.0, 0.5, 0.3)], 1)
    pp = np.array([0.004, 0.002, 0.003])
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(11))
    lA = hm.StateMatrix(3, K, np.log(pp), True)
    assert lA.nstates == 10621
    mu = np.asfortranarray(temps)
    _check(hm, O, S, lA, mu, 0.3, mode="auto")
    _check(hm, O, S[:2500], lA, mu, 0.3, mode="auto")  # short: still this engine (the sequential one cannot hold it)


def test_cli_sized_overlap_model_four_templates(hm, O):
    """Four K=60 templates with overlaps (src/hmmsort.jl:54, max_templates = 4): 21 123 states.  Neither the score
    columns nor the tables fit shared memory -- columns in a per-CTA global scratch, tables read from L2, traceback
    straight from global memory (939 multi-predecessor states)."""
    K, T = 60, 5000
    pars = [(3.0, 0.8, 0.2), (4.0, 0.3, 0.2), (2.0, 0.5, 0.3), (2.5, 0.6, 0.25)]
    temps = np.stack([hm.create_spike_template(K, *p) for p in pars], 1)
    pp = np.array([0.004, 0.002, 0.003, 0.002])
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(12))
    lA = hm.StateMatrix(4, K, np.log(pp), True)
    assert lA.nstates == 21123
    _check(hm, O, S, lA, np.asfortranarray(temps), 0.3, mode="auto")
