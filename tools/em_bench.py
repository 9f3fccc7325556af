import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
hm = ge.load_package()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1_800_000
S, lA_true, mu_true, _ = bench.make_c2(hm, seed=3, T=T)
N, K = 3, 60
lA = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), False)
mu = np.asfortranarray(0.7 * mu_true); sigma = float(np.std(S))
with hm.TrainContext(S) as ctx:
    for it in range(12):
        t0 = time.perf_counter()
        lp, pp, mu, sigma, ll, info = ctx.em_step(lA, mu, sigma, return_info=True)
        t1 = time.perf_counter()
        lA = hm.StateMatrix.from_states(lA.states, pp, K, lp, False)
        t2 = time.perf_counter()
        if it >= 8:
            print(f"iter {it}: em_step wall {1e3*(t1-t0):.3f} ms (device {info['device_ms']:.3f} ms, top {info['top_kernel_ms']:.3f}) "
                  f"rebuild {1e3*(t2-t1):.3f} ms chunks {info['n_chunks']} rep {info['fwd_repaired']}/{info['bwd_repaired']} ll {ll:.3f} sigma {sigma:.6f}")
