"""Small end-to-end case for compute-sanitizer (racecheck / memcheck / synccheck): ring decode (warp-specialised forward
kernel with mbarrier hand-off and named barriers, fused verify kernels, traceback), the fused E/M step, a two-shard
peer-memory decode and a two-shard E/M step.  Sizes are tiny because the sanitizer slows kernels down ~100x."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import __graft_entry__ as ge
from conftest import make_case
import torch
hm = ge.load_package(); O = ge.load_oracle(); ts = hm.timeshard
T = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000
for (N, K) in ((3, 60), (5, 60)):
    S, lA, mu, sig = make_case(hm, N, K, T, 7)
    hm.set_ring_params(2048, 512)
    x, ll = hm.viterbi(S, lA, mu, sig, mode="ring")
    hm.set_ring_params(0, 0)
    xo, llo = O.viterbi(S, lA, mu, sig)
    assert np.array_equal(x, xo) and abs(ll - llo) <= 1e-9 * abs(llo)
    print("viterbi ring ok", N, K, flush=True)
S, lA, mu, sig = make_case(hm, 3, 60, T, 8, rate_scale=2.0)
lA0 = hm.StateMatrix(3, 60, np.log(np.full(3, 0.01)), False)
r = hm.em_step(S, lA0, np.asfortranarray(0.7 * mu), float(np.std(S)), mode="ring")
o = O.em_step(S, O.OracleStateMatrix(3, 60, np.log(np.full(3, 0.01)), False), np.asfortranarray(0.7 * mu), float(np.std(S)))
assert np.abs(r[2] - o[2]).max() < 1e-6
print("em_step ok", flush=True)
dev = torch.device("cuda", 0)
spans = ts.shard_plan(T, 2, 4096, 512)
ys = [torch.from_numpy(np.ascontiguousarray(S[sp[0]:sp[1]])).to(dev) for sp in spans]
shards = [ts.Shard(y.data_ptr(), False, sp, T, 4096, 512, lA, mu, sig) for y, sp in zip(ys, spans)]
ptrs = [sh.p2p_init(k, 2)[1] for k, sh in enumerate(shards)]
xs = [torch.zeros(sp[3] - sp[2], dtype=torch.int16, device=dev) for sp in spans]
for sh in shards: sh.p2p_attach(block_ptrs=ptrs)
for sh, xx in zip(shards, xs): sh.p2p_launch(xx.data_ptr())
v = [sh.p2p_finish() for sh in shards]
xo, llo = O.viterbi(S, lA, mu, sig)
assert v[0][1] == 0 and np.array_equal(np.concatenate([xx.cpu().numpy() for xx in xs]), xo)
for sh in shards: sh.close()
print("p2p shards ok", flush=True)
es = [ts.EmShard(y.data_ptr(), False, sp, T, 4096) for y, sp in zip(ys, spans)]
em = ts.EmSharded(es, 3, 60, lA0.nstates, dev)
r2 = em.em_step(lA0, np.asfortranarray(0.7 * mu), float(np.std(S)))
assert np.abs(r2[2] - o[2]).max() < 1e-6
em.close()
print("em shards ok", flush=True)
