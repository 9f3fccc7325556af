"""ONE process, all GPUs of the box through hmm_set_devices: end-to-end (host buffers in, x out) throughput of
(a) one long recording (config 2, 18 M samples: time-sharded inside the library) and (b) a 32-channel batch of config-4
channels, from pinned host memory -- what a Julia caller of viterbi / the batch call gets from one C call."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench
hm = ge.load_package(); L = hm.lib()
nd = hm.device_count()
p = lambda a: a.ctypes.data_as(C.c_void_p)
def pinned(n, dtype):
    h = C.c_void_p(); hm._lib.check(L.hmm_host_alloc(C.byref(h), C.c_uint64(n * np.dtype(dtype).itemsize)))
    ct = {np.float64: C.c_double, np.int16: C.c_int16}[dtype]
    return np.ctypeslib.as_array(C.cast(h, C.POINTER(ct)), shape=(n,))
T = 18_000_000
S, lA, mu, sig = bench.make_c2(hm, 2, T=T)
yh = pinned(T, np.float64); yh[:] = S; xh = pinned(T, np.int16)
st = np.asfortranarray(lA.states); tr = np.ascontiguousarray(lA.transitions)
def one():
    ll = C.c_double(0)
    hm._lib.check(L.hmm_viterbi_f64(p(yh), C.c_int64(T), p(st), C.c_int32(3), C.c_int32(60), C.c_int32(lA.nstates), p(tr), C.c_int64(tr.size), p(mu), C.c_double(sig), p(xh), C.byref(ll), None, None))
    return ll.value
for devs in ([0], list(range(nd))):
    hm.set_devices(devs if len(devs) > 1 else None)
    for _ in range(2): ll = one()
    x_ref = xh.copy() if len(devs) == 1 else x_ref
    t0 = time.perf_counter()
    for _ in range(5): ll = one()
    dt = (time.perf_counter() - t0) / 5
    print(f"one 18M-sample recording, {len(devs)} device(s): {dt*1e3:.2f} ms per decode = {T/dt/1e9:.2f} Gsamples/s end to end, same x: {np.array_equal(xh, x_ref)} ll {ll:.6e}", flush=True)
# batch of 32 channels (N=4, K=48)
Cn = 32
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(16) as ex: ch = list(ex.map(lambda c: bench.make_c4_channel(hm, c, T), range(Cn)))
Yh = pinned(T * Cn, np.float64).reshape(Cn, T); Xh = pinned(T * Cn, np.int16).reshape(Cn, T)
for c in range(Cn): Yh[c] = ch[c][0]
stb = np.ascontiguousarray(np.concatenate([np.asfortranarray(c[1].states).ravel(order="F") for c in ch]))
trb = np.ascontiguousarray(np.concatenate([c[1].transitions for c in ch])); mub = np.ascontiguousarray(np.concatenate([c[2].ravel(order="F") for c in ch]))
sgb = np.asarray([c[3] for c in ch]); llb = np.zeros(Cn); info = hm.HmmInfo()
def batch():
    hm._lib.check(L.hmm_viterbi_batch_f64(p(Yh), C.c_int64(T), C.c_int32(Cn), p(stb), C.c_int32(0), C.c_int32(4), C.c_int32(48), C.c_int32(ch[0][1].nstates), p(trb), C.c_int64(ch[0][1].transitions.size), p(mub), p(sgb), p(Xh), p(llb), C.c_int32(2), C.byref(info)))
for devs in ([0], list(range(nd))):
    hm.set_devices(devs if len(devs) > 1 else None)
    batch()
    X_ref = Xh.copy() if len(devs) == 1 else X_ref
    t0 = time.perf_counter()
    for _ in range(2): batch()
    dt = (time.perf_counter() - t0) / 2
    print(f"{Cn}-channel batch, {len(devs)} device(s): {dt*1e3:.1f} ms per call = {Cn*T/dt/1e9:.2f} Gsamples/s end to end, same x: {np.array_equal(Xh, X_ref)}", flush=True)
hm.set_devices(None)
