"""Repair counts of the config-2 decode over several recording seeds (a repaired chunk is re-run sequentially by one
warp and costs ~0.3 ms: a boundary tolerance that is too tight shows up here as false mismatches)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench
hm = ge.load_package()
os.environ["HMMCUDA_NO_PIPELINE"] = "1"
for seed in range(2, 12):
    S, lA, mu, sig = bench.make_c2(hm, seed, T=18_000_000)
    x, ll, info = hm.viterbi(S, lA, mu, sig, mode="ring", return_info=True)
    print(seed, "chunks", info["n_chunks"], "repaired fwd/bwd", info["fwd_repaired"], info["bwd_repaired"], flush=True)
