"""Per-phase host timings of the time-sharded E/M step (all shards on one GPU)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench, torch
hm = ge.load_package(); ts = hm.timeshard
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T = 1_800_000
S, lA_true, mu_true, _ = bench.make_c2(hm, 3, T=T)
dev = torch.device("cuda", 0)
chunk_len, _w = ts.em_default_chunking(T, n, 3, 60)
spans = ts.shard_plan(T, n, chunk_len, 256)
xs = [torch.from_numpy(np.ascontiguousarray(S[sp[0]:sp[1]])).to(dev) for sp in spans]
shards = [ts.EmShard(x.data_ptr(), False, sp, T, chunk_len) for x, sp in zip(xs, spans)]
lA = hm.StateMatrix(3, 60, np.log(np.full(3, 0.01)), False)
em = ts.EmSharded(shards, 3, 60, lA.nstates, dev)
mu = np.asfortranarray(0.7 * mu_true); sigma = float(np.std(S))
for it in range(6):
    t0 = time.perf_counter()
    for sh, st, bd in zip(em.shards, em.stats, em.bnd):
        sh.estep(lA, mu, sigma, st.data_ptr(), bd.data_ptr())
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    r = em.em_step(lA, mu, sigma); t3 = time.perf_counter()
    print(f"n={n} chunk {chunk_len}: estep issue {1e3*(t1-t0):.3f} ms, sync {1e3*(t2-t1):.3f} ms, full em_step {1e3*(t3-t2):.3f} ms", flush=True)
