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
