"""Forward-kernel and whole-decode timing of the device-resident ring decode for the three BASELINE model shapes.
Usage: python tools/fwd_time.py [T]   (LIBHMMCUDA=... selects an alternative build)"""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import make_case
hm = ge.load_package(); L = hm.lib()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 18_000_000
dev = torch.device("cuda", 0)
p = lambda a: a.ctypes.data_as(C.c_void_p)
for (N, K) in ((3, 60), (4, 48), (5, 60)):
    S, lA, mu, sig = make_case(hm, N, K, T, 2)
    st = np.asfortranarray(lA.states); tr = np.ascontiguousarray(lA.transitions); sg = np.array([sig])
    y = torch.from_numpy(S).to(dev); x = torch.empty(T, dtype=torch.int16, device=dev)
    info = hm.HmmInfo(); ll = C.c_double(0)
    def step():
        hm._lib.check(L.hmm_viterbi_dev_f64(C.c_void_p(y.data_ptr()), C.c_int64(T), C.c_int32(1), p(st), C.c_int32(1), C.c_int32(N), C.c_int32(K),
              C.c_int32(lA.nstates), p(tr), C.c_int64(tr.size), p(mu), p(sg), C.c_void_p(x.data_ptr()), C.byref(ll), C.c_int32(2), C.byref(info)))
    L.hmm_set_profiling(1)
    tops = []
    for _ in range(6):
        step(); tops.append(info.top_kernel_ms)
    L.hmm_set_profiling(0)
    for _ in range(4): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 30
    for _ in range(n): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    dfma = (T + info.n_chunks * 512.0) * N * (K - 1)
    top = float(np.median(tops[2:]))
    print(f"N={N} K={K} T={T}: forward {top*1e3:.1f} us ({dfma/top/1e6/18421.7:.3f} of FP64 issue), decode step {dt*1e3:.4f} ms = {T/dt/1e9:.2f} Gsamples/s, "
          f"chunks {info.n_chunks} repaired {info.fwd_repaired}/{info.bwd_repaired} ll {ll.value:.6e}", flush=True)
