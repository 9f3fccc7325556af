"""Where one config-3 Baum-Welch iteration spends its time: the C call vs the host-side StateMatrix rebuild, and the
per-kernel stage times (HMMCUDA_EM_TIMING=1 prints them on stderr)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

hm = ge.load_package()
T = 1_800_000
S, lA_true, mu_true, _ = bench.make_c2(hm, seed=3, T=T)
N, K = 3, 60
lA = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), False)
mu = np.asfortranarray(0.7 * mu_true)
sigma = float(np.std(S))
with hm.TrainContext(S) as ctx:
    for _ in range(3):
        ctx.em_step(lA, mu, sigma)
    tc = tr = 0.0
    dev = ker = 0.0
    n = 20
    for _ in range(n):
        t0 = time.perf_counter()
        lp, pp, mu, sigma, ll, info = ctx.em_step(lA, mu, sigma, return_info=True)
        t1 = time.perf_counter()
        lA = hm.StateMatrix.from_states(lA.states, pp, K, lp, False)
        t2 = time.perf_counter()
        tc += t1 - t0
        tr += t2 - t1
        dev += info["device_ms"]
    print(f"per iteration: C call {tc / n * 1e6:.1f} us (device events {dev / n * 1e3:.1f} us), rebuild {tr / n * 1e6:.1f} us", file=sys.stderr)
    lA0, mu0 = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), False), np.asfortranarray(0.7 * mu_true)
    for _ in range(2):
        t0 = time.perf_counter()
        lA_r, mu_r, s_r, lls, info = ctx.run(lA0, mu0, float(np.std(S)), 20, return_info=True)
        dt = time.perf_counter() - t0
        print(f"hmm_train_run: {dt / 20 * 1e6:.1f} us per iteration ({20 / dt:.1f} iters/s), device events {info['device_ms'] / 20 * 1e3:.1f} us, chunks {info['n_chunks']}, repaired {info['fwd_repaired']}/{info['bwd_repaired']}", file=sys.stderr)
