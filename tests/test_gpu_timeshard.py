"""Time-sharded decode of one recording (config 5 mechanism): the recording is cut into
shards with ghost chunks, decoded shard by shard and stitched with verified boundary
exchanges.  Here all shards run on one GPU (messages through host memory); on a
multi-GPU box the same protocol runs over NCCL (`viterbi_time_sharded_dist`, bench.py
--workload c5).  The stitched result must be bit-identical to the unsharded decode."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_time_sharded_equals_unsharded(hm, O, case_factory, n_shards):
    S, lA, mu, sig = case_factory(3, 60, 400_000, 61)
    x_ref, ll_ref = O.viterbi(S, lA, mu, sig)
    x, ll, info = hm.viterbi_time_sharded(S, lA, mu, sig, n_shards, chunk_len=4096, warmup=512, return_info=True)
    assert np.array_equal(x, x_ref)
    assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
    assert info["fwd_rounds"] >= 1 and info["trace_rounds"] >= 1


def test_time_sharded_default_chunking_and_models(hm, case_factory):
    """Default chunking (one wave over n_shards GPUs) on a longer recording, N=5, K=60."""
    S, lA, mu, sig = case_factory(5, 60, 3_000_000, 62)
    x1, ll1 = hm.viterbi(S, lA, mu, sig, mode="ring")
    x, ll, info = hm.viterbi_time_sharded(S, lA, mu, sig, 4, return_info=True)
    assert np.array_equal(x, x1) and abs(ll - ll1) <= 1e-9 * abs(ll1)


def test_time_sharded_dense_spiking_repairs_across_shards(hm, O, case_factory):
    """Short warm-up + dense firing: speculative shard starts may fail and must be repaired
    through the exchanged boundary vectors."""
    S, lA, mu, sig = case_factory(3, 60, 200_000, 63, rate_scale=8.0)
    x_ref, ll_ref = O.viterbi(S, lA, mu, sig)
    x, ll, info = hm.viterbi_time_sharded(S, lA, mu, sig, 5, chunk_len=1024, warmup=256, return_info=True)
    assert np.array_equal(x, x_ref) and abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
