"""Config-4 share, device-resident: one hmm_viterbi_dev_f64 call with C channels vs C single-channel calls."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import torch
hm = ge.load_package(); L = hm.lib()
Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 18_000_000
N, K = 4, 48
rng = np.random.default_rng(1000)
dev = torch.device("cuda", 0)
y = torch.empty((Cn, T), dtype=torch.float64, device=dev)
sts, trs, mus, sig = [], [], [], []
for c in range(Cn):
    prm = [(rng.uniform(2, 4), rng.uniform(0.3, 0.9), rng.uniform(0.1, 0.3)) for _ in range(N)]
    temps = np.stack([hm.create_spike_template(K, *q) for q in prm], axis=1)
    pp = rng.uniform(0.0005, 0.004, size=N)
    if c < 4:
        y[c].copy_(torch.from_numpy(hm.create_signal(T, 0.3, pp, temps, hm.make_rng(1000 + c))))
        keep = (temps, pp)
    else:
        y[c].copy_(y[c % 4]); 
    if c >= 4:
        sts.append(sts[c % 4]); trs.append(trs[c % 4]); mus.append(mus[c % 4]); sig.append(0.3); continue
    mu = np.asfortranarray(temps.copy()); mu[0, :] = 0
    lA = hm.StateMatrix(N, K, np.log(pp), False)
    sts.append(np.asfortranarray(lA.states).ravel(order="F")); trs.append(lA.transitions); mus.append(mu.ravel(order="F")); sig.append(0.3)
st = np.ascontiguousarray(np.concatenate(sts)); tr = np.ascontiguousarray(np.concatenate(trs)); mu = np.ascontiguousarray(np.concatenate(mus)); sg = np.asarray(sig)
x = torch.empty((Cn, T), dtype=torch.int16, device=dev); x2 = torch.empty_like(x)
ll = np.zeros(Cn); ll2 = np.zeros(Cn); info = hm.HmmInfo()
p = lambda a: a.ctypes.data_as(C.c_void_p)
ntr = trs[0].size; ns = lA.nstates
def batched():
    hm._lib.check(L.hmm_viterbi_dev_f64(C.c_void_p(y.data_ptr()), C.c_int64(T), C.c_int32(Cn), p(st), C.c_int32(0), C.c_int32(N), C.c_int32(K),
                                        C.c_int32(ns), p(tr), C.c_int64(ntr), p(mu), p(sg), C.c_void_p(x.data_ptr()), p(ll), C.c_int32(2), C.byref(info)))
def single():
    for c in range(Cn):
        hm._lib.check(L.hmm_viterbi_dev_f64(C.c_void_p(y[c].data_ptr()), C.c_int64(T), C.c_int32(1), p(sts[c]), C.c_int32(1), C.c_int32(N), C.c_int32(K),
                                            C.c_int32(ns), p(trs[c]), C.c_int64(ntr), p(mus[c]), p(sg[c:c+1]), C.c_void_p(x2[c].data_ptr()),
                                            p(ll2[c:c+1]), C.c_int32(2), C.byref(info)))
for f, name in ((batched, "one call, C channels"), (single, "C single-channel calls")):
    f(); f(); torch.cuda.synchronize()
    t0 = time.perf_counter(); f(); f(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    print(f"{name}: {dt*1e3:.2f} ms -> {Cn*T/dt/1e6:.0f} Msamples/s")
print("same:", bool(torch.equal(x, x2)), np.array_equal(ll, ll2))
