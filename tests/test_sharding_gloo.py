"""N>1 host logic on CPU: world_size-2 `gloo` process group.  Channel sharding,
result gathering and the max-over-ranks timing reduction are exercised; the
per-shard decode is done by the CPU oracle here (on a GPU box it is
hm.viterbi_batch -- see bench.py)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions(hm):
    sh = hm.sharding
    for n in (0, 1, 7, 16, 128, 129):
        for world in (1, 2, 3, 8):
            spans = [sh.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sh.shard_range(128, 8, 3) == (48, 64)  # config 4: 16 channels per GPU
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def test_shard_time(hm):
    sh = hm.sharding
    T = 108_000_000  # config 5
    spans = [sh.shard_time(T, 8, r, halo=4096) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == T
    for (a, b, lo, hi), (a2, b2, lo2, hi2) in zip(spans, spans[1:]):
        assert b == a2 and a % 256 == 0 and lo2 == a2 - 4096 and hi == b + 4096
    assert spans[0][2] == 0 and spans[-1][3] == T


def test_time_shard_plan(hm):
    """Config 5 style time sharding: whole chunks per shard, one ghost chunk on either side."""
    plan = hm.timeshard.shard_plan(108_000_000, 8, 7680)
    assert plan[0][0] == 0 and plan[0][2] == 0 and plan[-1][3] == 108_000_000 and plan[-1][1] == 108_000_000
    for (lb, le, mb, me), (lb2, le2, mb2, me2) in zip(plan, plan[1:]):
        assert me == mb2 and lb2 == mb2 - 7680 and le == me + 7680 and mb % 7680 == 0
    sizes = [me - mb for _, _, mb, me in plan]
    assert max(sizes) - min(sizes) <= 2 * 7680  # one chunk of imbalance + the partial last chunk
    with pytest.raises(ValueError):
        hm.timeshard.shard_plan(10_000, 8, 4096)


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    import __graft_entry__ as ge
    hm = ge.load_package(); O = ge.load_oracle()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["MASTER_PORT"], rank=rank, world_size=world)
    C, T, K, N = 5, 3000, 12, 2
    temps = np.stack([hm.create_spike_template(K, 3.0, 0.8, 0.2), hm.create_spike_template(K, 4.0, 0.3, 0.2)], 1)
    pp = np.array([0.01, 0.005])
    Y = np.asfortranarray(np.stack([hm.create_signal(T, 0.3, pp, temps, hm.make_rng(100 + c)) for c in range(C)], 1))
    mu = np.asfortranarray(temps.copy()); mu[0, :] = 0
    models = [(hm.StateMatrix(N, K, np.log(pp), False), mu * (1 + 0.05 * c), 0.3) for c in range(C)]
    def decode(Ys, ms):   # stand-in for hm.viterbi_batch on a GPU rank
        xs = [O.viterbi(Ys[:, k], *ms[k]) for k in range(len(ms))]
        return np.asfortranarray(np.stack([a for a, _ in xs], 1)), np.array([b for _, b in xs])
    x_all, ll_all = hm.sharding.decode_channels_sharded(decode, Y, models)
    x_loc, ll_loc, span = hm.sharding.decode_channels_sharded(decode, Y, models, gather=False)
    tmax = hm.sharding.all_reduce_max(1.0 + rank)
    ref = [O.viterbi(Y[:, c], *models[c]) for c in range(C)]
    ok = all(np.array_equal(x_all[:, c], ref[c][0]) and ll_all[c] == ref[c][1] for c in range(C))
    ok = ok and span == hm.sharding.shard_range(C, world, rank) and x_loc.shape[1] == span[1] - span[0]
    print(json.dumps({{"rank": rank, "ok": bool(ok), "tmax": tmax, "span": span}}))
    dist.destroy_process_group()
""")


def test_boundary_vectors_agree_rule(hm):
    """Shard-level boundary check of the time-sharded E/M step: equal up to a constant over EVERY finite entry.  An entry
    far below the maximum still counts (at high SNR a pending chain's entry is multiplied by e^(+1000s) later), and the
    pattern of -inf entries must be the same."""
    agree = hm.timeshard._vectors_agree
    a = np.array([0.0, -3.0, -900.0, -np.inf, 5.0])
    ok, c = agree(a + 7.25, a)
    assert ok and c == 7.25
    b = a.copy()
    b[2] = -870.0  # 900 below the maximum in `a`: used to be ignored
    assert not agree(a, b)[0]
    b = a.copy()
    b[3] = -2000.0  # finite where the other vector has -inf
    assert not agree(a, b)[0]
    b = a.copy()
    b[1] += 1e-6
    assert not agree(a, b)[0]
    b[1] = a[1] + 1e-13
    assert agree(a, b)[0]


def test_channel_sharding_world2_gloo(hm, O, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    import json
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-2000:]
        outs.append(json.loads(out.strip().splitlines()[-1]))
    assert all(o["ok"] for o in outs)
    assert all(o["tmax"] == 2.0 for o in outs)
    assert sorted(tuple(o["span"]) for o in outs) == [(0, 3), (3, 5)]


DIST_WORKER = textwrap.dedent("""
    import os, sys, json, ctypes
    import numpy as np
    sys.path.insert(0, {root!r})
    import torch, torch.distributed as dist
    import __graft_entry__ as ge
    hm = ge.load_package()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["MASTER_PORT"], rank=rank, world_size=world)

    def view(ptr, n, ct):   # the decoder passes raw addresses, as it does to the C ABI
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,))

    class FakeShard:
        # Stand-in with hmm_vshard's interface: "decodes" nothing, but produces and consumes the protocol's
        # messages so that their routing, ordering and termination can be checked on CPU.
        bvec, summary_len = 7, 2 * 7 + 4
        def __init__(self):
            self.first, self.last = rank == 0, rank == world - 1
            self.calls, self.fwd_in, self.trace_in = [], None, None
            self.force_bad, self.fwd_repairs_left = False, 0
        def forward(self): self.calls.append("forward")
        def trace(self): self.calls.append("trace")
        def fwd_verify(self, count=True):
            self.calls.append("fwd_verify")
            if count and self.fwd_repairs_left > 0:
                self.fwd_repairs_left -= 1
                return 1
            return 0
        def trace_verify(self, count=True):
            self.calls.append("trace_verify")
            return 0
        def summary_dev(self, x_ptr, summ_ptr):
            s = view(summ_ptr, self.summary_len, ctypes.c_double)
            s[:] = 0.0
            s[2 * self.bvec + 2] = 10.0 + rank          # partial ll
        def judge_dev(self, gath_ptr, n, out_ptr):
            g = view(gath_ptr, n * self.summary_len, ctypes.c_double).reshape(n, self.summary_len)
            out = view(out_ptr, 2, ctypes.c_double)
            out[0] = g[:, 2 * self.bvec + 2].sum()
            out[1] = 1.0 if self.force_bad else 0.0
        def fwd_get(self, out_ptr): view(out_ptr, self.bvec, ctypes.c_double)[:] = 100.0 + rank
        def fwd_set(self, in_ptr): self.fwd_in = float(view(in_ptr, self.bvec, ctypes.c_double)[0])
        def trace_get(self, out_ptr): view(out_ptr, 1, ctypes.c_int64)[0] = 1000 + rank
        def trace_set(self, in_ptr): self.trace_in = int(view(in_ptr, 1, ctypes.c_int64)[0])
        def finish(self, x_ptr=None): return 10.0 + rank
        def close(self): pass

    sh = FakeShard()
    dec = hm.timeshard.DistDecoder(0, (0, 0, 0, 0), 0, 0, 0, None, None, 0.0, 0, torch.device("cpu"), shard=sh)
    want = sum(10.0 + r for r in range(world))
    ll_fast = dec.decode()                       # fast path: one all-gather, verdict 0
    fast_calls = list(sh.calls)
    sh.calls.clear()
    sh.force_bad = True                          # verdict != 0 on every rank -> iterative fallback
    sh.fwd_repairs_left = 1 if rank == 1 else 0  # rank 1 repairs once: everyone must go a second forward round
    ll_slow = dec.decode()
    ok = ll_fast == want and ll_slow == want and dec.stats["fallbacks"] == 1
    ok = ok and fast_calls == ["forward", "fwd_verify", "trace", "trace_verify"]
    ok = ok and sh.calls.count("fwd_verify") == 1 + 2 and sh.calls.count("trace_verify") == 1 + 1
    ok = ok and (sh.first or sh.fwd_in == 100.0 + rank - 1) and (sh.last or sh.trace_in == 1000 + rank + 1)
    print(json.dumps({{"rank": rank, "ok": bool(ok), "calls": sh.calls, "stats": dec.stats}}))
    dist.destroy_process_group()
""")


def test_dist_decoder_protocol_world3_gloo(hm, tmp_path):
    """DistDecoder's two protocols over a real (gloo) process group of three ranks, with a stand-in shard: the fast
    path is one all-gather and no point-to-point traffic; a non-zero verdict runs the iterative fallback, whose
    boundary vectors travel right and traceback states left, and whose rounds end on all ranks together."""
    script = tmp_path / "dist_worker.py"
    script.write_text(DIST_WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(3):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="3", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    import json
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-3000:]
        o = json.loads(out.strip().splitlines()[-1])
        assert o["ok"], o
