"""One rank's share of config 5 at 8 GPUs (13.5 M samples, N=5, K=60, chunk 11520, warm-up 256) decoded on one GPU:
for an ncu launch list of the kernels a rank runs per step."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench
hm = ge.load_package()
T = 13_500_000
S, lA, mu, sigma = bench.make_c5(hm, T=T)
os.environ["HMMCUDA_NO_PIPELINE"] = "1"
hm.set_ring_params(11520, 256)
for _ in range(4):
    t0 = time.perf_counter()
    x, ll, info = hm.viterbi(S, lA, mu, sigma, mode="ring", return_info=True)
    dt = time.perf_counter() - t0
print(f"kernels {info['kernel_ms']:.3f} ms top {info['top_kernel_ms']:.3f} ms chunks {info['n_chunks']} wall {dt*1e3:.1f} ms")
