"""Per-phase cycle counts of the forward kernel's producer / consumer warps (library built with -DHMM_PHASE_TIMING)
and the decode-step timing of the default build.  Usage: LIBHMMCUDA=tools/alt/libhmmcuda_pt.so python tools/fwd_phase.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench
hm = ge.load_package()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 18_000_000
S, lA, mu, sig = bench.make_c2(hm, 2, T=T)
os.environ["HMMCUDA_NO_PIPELINE"] = "1"
hm.lib().hmm_set_profiling(1)
for it in range(3):
    x, ll, info = hm.viterbi(S, lA, mu, sig, mode="ring", return_info=True)
    print("top %.4f ms kernels %.4f ms chunks %d" % (info["top_kernel_ms"], info["kernel_ms"], info["n_chunks"]), flush=True)
