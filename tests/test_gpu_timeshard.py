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


def test_one_collective_protocol_summaries_and_judge(hm, O, case_factory):
    """hmm_vshard_summary_dev / hmm_vshard_judge_dev (bench.py --workload c5): every shard decodes on its own,
    the boundary summaries are gathered (here: concatenated) and one kernel checks every shard boundary.  The
    stitched result must equal the oracle's; a corrupted summary must be detected."""
    import torch

    ts = hm.timeshard
    T, n = 400_000, 4
    S, lA, mu, sig = case_factory(3, 60, T, 64)
    x_ref, ll_ref = O.viterbi(S, lA, mu, sig)
    dev = torch.device("cuda", 0)
    spans = ts.shard_plan(T, n, 4096)
    xs, summaries, shards = [], [], []
    for span in spans:
        y_loc = torch.from_numpy(np.ascontiguousarray(S[span[0]:span[1]])).to(dev)
        sh = ts.Shard(y_loc.data_ptr(), False, span, T, 4096, 512, lA, mu, sig)
        x_main = torch.empty(span[3] - span[2], dtype=torch.int16, device=dev)
        summ = torch.zeros(sh.summary_len, dtype=torch.float64, device=dev)
        sh.forward()
        sh.fwd_verify(count=False)
        sh.trace()
        sh.trace_verify(count=False)
        sh.summary_dev(x_main.data_ptr(), summ.data_ptr())
        torch.cuda.synchronize()
        xs.append(x_main.cpu().numpy())
        summaries.append(summ)
        shards.append((sh, y_loc))
    gath = torch.cat(summaries).contiguous()
    res = torch.zeros(2, dtype=torch.float64, device=dev)
    shards[0][0].judge_dev(gath.data_ptr(), n, res.data_ptr())
    torch.cuda.synchronize()
    ll, bad = res.tolist()
    assert bad == 0
    assert np.array_equal(np.concatenate(xs), x_ref)
    assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
    # a wrong traceback state at a shard boundary, and a wrong forward vector, are both caught
    slen = shards[0][0].summary_len
    bvec = shards[0][0].bvec
    g2 = gath.clone()
    g2[1 * slen + 2 * bvec + 0] += 8.0
    shards[0][0].judge_dev(g2.data_ptr(), n, res.data_ptr())
    torch.cuda.synchronize()
    assert res.tolist()[1] == 1
    g3 = gath.clone()
    g3[2 * slen + 5] += 1e-3  # end vector of shard 2 (entry 5 of the forward vector)
    shards[0][0].judge_dev(g3.data_ptr(), n, res.data_ptr())
    torch.cuda.synchronize()
    assert res.tolist()[1] == 1
    for sh, _ in shards:
        sh.close()


def test_dist_decoder_single_rank(hm, O, case_factory):
    """DistDecoder (the torch.distributed driver of bench.py --workload c5) with one rank and no process group:
    the whole one-collective path -- local decode, summary, judge, single read -- against the oracle, twice (a
    decoder is re-run per recording), on torch's default stream (the decoder then creates its own)."""
    import torch

    ts = hm.timeshard
    T = 300_000
    S, lA, mu, sig = case_factory(5, 60, T, 65)
    x_ref, ll_ref = O.viterbi(S, lA, mu, sig)
    dev = torch.device("cuda", 0)
    span = ts.shard_plan(T, 1, 4096)[0]
    y_loc = torch.from_numpy(np.ascontiguousarray(S[span[0]:span[1]])).to(dev)
    x_main = torch.zeros(span[3] - span[2], dtype=torch.int16, device=dev)
    dec = ts.DistDecoder(y_loc.data_ptr(), span, T, 4096, 512, lA, mu, sig, x_main.data_ptr(), dev)
    try:
        for _ in range(2):
            x_main.zero_()
            ll = dec.decode()
            torch.cuda.synchronize()
            assert np.array_equal(x_main.cpu().numpy(), x_ref)
            assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
        assert dec.stats["fallbacks"] == 0 and dec.stats["decodes"] == 2
    finally:
        dec.close()


def test_judge_sees_what_a_locally_repaired_chunk_was_started_from(hm, O, case_factory):
    """A shard whose first main chunk was REPAIRED locally from a wrong ghost-chunk end vector must not pass the
    judge on the strength of its stale speculative start vector (which may well equal the neighbour's true end
    vector): after an exact re-run the start vector on record is the one the chunk was really computed from.
    Here the ghost's end vector of shard 1 is overwritten with a perturbed copy before verification."""
    import torch

    ts = hm.timeshard
    T, n = 300_000, 3
    S, lA, mu, sig = case_factory(3, 60, T, 66)
    dev = torch.device("cuda", 0)
    spans = ts.shard_plan(T, n, 4096)
    summaries, shards = [], []
    for r, span in enumerate(spans):
        y_loc = torch.from_numpy(np.ascontiguousarray(S[span[0]:span[1]])).to(dev)
        sh = ts.Shard(y_loc.data_ptr(), False, span, T, 4096, 512, lA, mu, sig)
        summ = torch.zeros(sh.summary_len, dtype=torch.float64, device=dev)
        sh.forward()
        if r == 1:
            # what the left neighbour would send, but wrong: shard 0's true end vector with one entry moved
            v = shards[0][0].fwd_get()
            v[1 + 7] += 0.5
            sh.fwd_set(v)
            assert sh.fwd_verify(count=True) >= 1  # the first main chunk is re-run from the wrong vector
        else:
            sh.fwd_verify(count=False)
        sh.trace()
        sh.trace_verify(count=False)
        sh.summary_dev(0, summ.data_ptr())
        summaries.append(summ)
        shards.append((sh, y_loc))
    gath = torch.cat(summaries).contiguous()
    res = torch.zeros(2, dtype=torch.float64, device=dev)
    shards[0][0].judge_dev(gath.data_ptr(), n, res.data_ptr())
    torch.cuda.synchronize()
    assert res.tolist()[1] >= 1, "the judge accepted a shard that was started from a wrong boundary vector"
    for sh, _ in shards:
        sh.close()


def test_time_sharded_forced_repairs(hm, O, case_factory, monkeypatch):
    """Exchange/verify rounds with every chunk of every shard flagged (HMMCUDA_DEBUG_FLAG_EVERY=1): the first main
    chunk of each shard is first repaired from its ghost chunk, then again from the neighbour's true vector."""
    S, lA, mu, sig = case_factory(3, 60, 250_000, 67, rate_scale=2.0)
    x_ref, ll_ref = O.viterbi(S, lA, mu, sig)
    monkeypatch.setenv("HMMCUDA_DEBUG_FLAG_EVERY", "1")
    x, ll, info = hm.viterbi_time_sharded(S, lA, mu, sig, 4, chunk_len=4096, warmup=512, return_info=True)
    assert np.array_equal(x, x_ref) and abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)
    assert info["fwd_repaired"] > 0 and info["trace_repaired"] > 0, info


def test_shard_plan_short_tail_is_merged(hm, O, case_factory):
    """T = k * chunk_len + a remainder shorter than the look-back: the short final span joins the previous shard
    instead of becoming a shard of its own (hmm_vshard_create would reject its neighbour's ghost)."""
    ts = hm.timeshard
    T = 10 * 4096 + 100
    plan = ts.shard_plan(T, 5, 4096)
    assert plan[-1][3] == T and plan[-1][3] - plan[-1][2] == 2 * 4096 + 100 and all(p[3] - p[2] >= 4096 for p in plan)
    S, lA, mu, sig = case_factory(3, 60, T, 68)
    x_ref, ll_ref = O.viterbi(S, lA, mu, sig)
    x, ll = hm.viterbi_time_sharded(S, lA, mu, sig, 5, chunk_len=4096, warmup=512)
    assert np.array_equal(x, x_ref) and abs(ll - ll_ref) <= 1e-9 * abs(ll_ref)


@pytest.mark.parametrize("n", [1, 3, 5])
def test_peer_memory_protocol_single_process(hm, O, case_factory, n):
    """hmm_vshard_p2p_*: every shard stores its summary straight into every other shard's exchange block and raises
    a flag; each shard's judge waits for all flags and checks every boundary.  Here the shards are driven from one
    process on one device (block pointers instead of CUDA-IPC handles); three decodes -- the second and third
    re-launch the captured CUDA graph and use the other half of the double-buffered exchange block."""
    import torch

    ts = hm.timeshard
    T = 300_000
    S, lA, mu, sig = case_factory(3, 60, T, 69)
    x_ref, ll_ref = O.viterbi(S, lA, mu, sig)
    dev = torch.device("cuda", 0)
    spans = ts.shard_plan(T, n, 4096)
    shards, ys, xs, ptrs = [], [], [], []
    for r, span in enumerate(spans):
        y_loc = torch.from_numpy(np.ascontiguousarray(S[span[0]:span[1]])).to(dev)
        sh = ts.Shard(y_loc.data_ptr(), False, span, T, 4096, 512, lA, mu, sig)
        _, ptr = sh.p2p_init(r, n)
        shards.append(sh); ys.append(y_loc); ptrs.append(ptr)
        xs.append(torch.zeros(span[3] - span[2], dtype=torch.int16, device=dev))
    for sh in shards:
        sh.p2p_attach(block_ptrs=ptrs)
    for it in range(3):
        for x in xs:
            x.zero_()
        for sh, x in zip(shards, xs):
            sh.p2p_launch(x.data_ptr())
        verdicts = [sh.p2p_finish() for sh in shards]
        for ll, bad in verdicts:
            assert bad == 0 and abs(ll - ll_ref) <= 1e-9 * abs(ll_ref), (it, ll, bad)
        assert len({v[0] for v in verdicts}) == 1  # every shard computes the same total, bit for bit
        assert np.array_equal(np.concatenate([x.cpu().numpy() for x in xs]), x_ref)
    for sh in shards:
        sh.close()
