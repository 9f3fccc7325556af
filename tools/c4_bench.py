"""Config 4 shape on ONE GPU's share: C channels x T samples, independent per-channel HMMs
(N=4, K=48), through the host-pointer batch API (pinned host memory)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
hm = ge.load_package(); L = hm.lib()
Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 18_000_000
N, K = 4, 48
rng = np.random.default_rng(1000)
yh, xh = C.c_void_p(), C.c_void_p()
hm._lib.check(L.hmm_host_alloc(C.byref(yh), C.c_uint64(8 * T * Cn)))
hm._lib.check(L.hmm_host_alloc(C.byref(xh), C.c_uint64(2 * T * Cn)))
Y = np.ctypeslib.as_array(C.cast(yh, C.POINTER(C.c_double)), shape=(Cn, T))   # channel-major == [T x C] column-major
X = np.ctypeslib.as_array(C.cast(xh, C.POINTER(C.c_int16)), shape=(Cn, T))
sts, trs, mus, sig = [], [], [], []
for c in range(Cn):
    prm = [(rng.uniform(2, 4), rng.uniform(0.3, 0.9), rng.uniform(0.1, 0.3)) for _ in range(N)]
    temps = np.stack([hm.create_spike_template(K, *q) for q in prm], axis=1)
    pp = rng.uniform(0.0005, 0.004, size=N)
    Y[c] = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(1000 + c))
    mu = np.asfortranarray(temps.copy()); mu[0, :] = 0
    lA = hm.StateMatrix(N, K, np.log(pp), False)
    sts.append(np.asfortranarray(lA.states).ravel(order="F")); trs.append(lA.transitions); mus.append(mu.ravel(order="F")); sig.append(0.3)
st = np.ascontiguousarray(np.concatenate(sts)); tr = np.ascontiguousarray(np.concatenate(trs)); mu = np.ascontiguousarray(np.concatenate(mus)); sig = np.asarray(sig)
ll = np.empty(Cn); info = hm.HmmInfo()
p = lambda a: a.ctypes.data_as(C.c_void_p)
def run():
    hm._lib.check(L.hmm_viterbi_batch_f64(yh, C.c_int64(T), C.c_int32(Cn), p(st), C.c_int32(0), C.c_int32(N), C.c_int32(K),
                                          C.c_int32(lA.nstates), p(tr), C.c_int64(lA.transitions.size), p(mu), p(sig), xh, p(ll),
                                          C.c_int32(2), C.byref(info)))
run(); run()
t0 = time.perf_counter(); run(); dt = time.perf_counter() - t0
print(f"config-4 share: {Cn} channels x {T} samples: {dt*1e3:.1f} ms end to end -> {Cn*T/dt/1e6:.0f} Msamples/s (pinned host), noise frac {(X==1).mean():.3f}")
