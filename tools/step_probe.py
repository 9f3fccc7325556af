"""Host-side timing of the device-resident config-2 decode step (hmm_viterbi_dev_f64), per call, with and without
torch.distributed initialised -- to see what a step costs beyond its kernels when several ranks share the host."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench, torch
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", os.environ.get("PROBE_DEV", "0")))
use_dist = world > 1 and not os.environ.get("PROBE_NO_DIST")
torch.cuda.set_device(local)
if use_dist:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hm = ge.load_package(); L = hm.lib(); L.hmm_set_device(C.c_int32(local))
T = 18_000_000
S, lA, mu, sig = bench.make_c2(hm, 2 + rank, T=T)
dev = torch.device("cuda", local)
y = torch.from_numpy(S).to(dev); x = torch.empty(T, dtype=torch.int16, device=dev)
st = np.asfortranarray(lA.states); tr = np.ascontiguousarray(lA.transitions); sg = np.array([sig])
p = lambda a: a.ctypes.data_as(C.c_void_p)
info = hm.HmmInfo(); ll = C.c_double(0)
if os.environ.get("PROBE_TORCH_STREAM"):
    work = torch.cuda.Stream(device=dev); torch.cuda.set_stream(work); L.hmm_set_stream(C.c_void_p(work.cuda_stream))
def step():
    hm._lib.check(L.hmm_viterbi_dev_f64(C.c_void_p(y.data_ptr()), C.c_int64(T), C.c_int32(1), p(st), C.c_int32(1), C.c_int32(3), C.c_int32(60),
          C.c_int32(lA.nstates), p(tr), C.c_int64(tr.size), p(mu), p(sg), C.c_void_p(x.data_ptr()), C.byref(ll), C.c_int32(2), C.byref(info)))
for _ in range(6): step()
if use_dist: dist.barrier(device_ids=[local])
torch.cuda.synchronize()
ts = []
for _ in range(200):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(f"rank {rank}/{world} dist={use_dist} graph={'no' if os.environ.get('HMMCUDA_NO_GRAPH') else 'yes'} torch_stream={bool(os.environ.get('PROBE_TORCH_STREAM'))}: "
      f"step wall ms median {np.median(ts):.4f} p10 {np.percentile(ts,10):.4f} p90 {np.percentile(ts,90):.4f} max {ts.max():.4f}", flush=True)
if use_dist: dist.destroy_process_group()
