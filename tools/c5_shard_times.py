"""Config 5 at 8 ranks, emulated on ONE GPU: the eight shards of the 108 M-sample recording are created in one process
(peer-memory protocol with block pointers) and each shard's local decode + summary exchange is timed on its own with
CUDA events -- what one rank does per step, without the other ranks' GPUs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench, torch
hm = ge.load_package(); ts = hm.timeshard
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = 108_000_000
S, lA, mu, sig = bench.make_c5(hm, T=T)
dev = torch.device("cuda", 0)
chunk_len, warm = ts.default_chunking(T, n, lA.N, lA.K)
if n > 1: warm = 256
spans = ts.shard_plan(T, n, chunk_len, warm)
shards, ys, xs, ptrs = [], [], [], []
for r, span in enumerate(spans):
    y_loc = torch.from_numpy(np.ascontiguousarray(S[span[0]:span[1]])).to(dev)
    sh = ts.Shard(y_loc.data_ptr(), False, span, T, chunk_len, warm, lA, mu, sig)
    _, ptr = sh.p2p_init(r, n)
    shards.append(sh); ys.append(y_loc); ptrs.append(ptr)
    xs.append(torch.zeros(span[3] - span[2], dtype=torch.int16, device=dev))
for sh in shards: sh.p2p_attach(block_ptrs=ptrs)
work = torch.cuda.Stream(device=dev); torch.cuda.set_stream(work)
hm.lib().hmm_set_stream(__import__("ctypes").c_void_p(work.cuda_stream))
times = np.zeros((6, n)); tj = np.zeros(6)
for it in range(6):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for r, (sh, x) in enumerate(zip(shards, xs)):
        evs[r][0].record(); sh.p2p_launch(x.data_ptr()); evs[r][1].record()
    t0 = time.perf_counter()
    v = [sh.p2p_finish() for sh in shards]
    tj[it] = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    times[it] = [a.elapsed_time(b) for a, b in evs]
print(f"n={n} chunk_len={chunk_len} warm={warm}; per-shard local decode + exchange (ms, median of last 4):", np.round(np.median(times[2:], axis=0), 4))
print("judge+sync per shard (ms, host):", np.round(np.median(tj[2:]) * 1e3, 4), "verdict", v[0])
hm.lib().hmm_set_stream(None)
