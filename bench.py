#!/usr/bin/env python
"""bench.py -- headline benchmark of the HMM hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N=1: BASELINE config 2 -- single channel, 30 kHz x 10 min
(18 000 000 samples, Float64), N=3 neurons x K=60 states, Viterbi decode with
fixed lA/mu/sigma.  One "step" = one full decode of the recording.  At N>1 every
rank decodes its own 18 M-sample channel (independent electrode channels, no
data-path collective): weak scaling, value = all ranks' samples / max time.

Timed three ways (all printed in ONE JSON line by rank 0):
  value    device-resident decode through hmm_viterbi_dev_f64 (y already in HBM,
           x left in HBM), wall clock over K steps bracketed by barrier +
           device synchronise, max over ranks.
  e2e      the reference-facing call hmm_viterbi_f64 semantics with HOST buffers
           (pinned): H2D of y and D2H of x inside the timed region.
  roofline dominant kernel (ring forward: FIR + max-plus recursion) -- algorithmic
           10 B/sample (8 B y read + 2 B x write) over its CUDA-event
           time measured inside the library on its launching stream.
A second, smaller block reports the Baum-Welch half of the metric (config 3).
`--impl reference` times the CPU oracle port of the reference algorithm on the
host cores instead (Julia is not available).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

T_C2 = 18_000_000
T_C3 = 1_800_000
BYTES_PER_SAMPLE = 10  # 8 B read of S + 2 B write of x
# dram__bytes_read.sum + dram__bytes_write.sum per launch of ring_vit_forward_ws<3,8,59,8,double> from the ncu --set full
# capture profiles/r02_ring_vit_forward_ws_ncu.md (149.7 MB read + 49.2 MB written; ncu cannot run inside a benchmark)
NCU_TRAFFIC_BYTES = 198830592
# dram bytes of one E/M iteration at config 3, summed over its 8 kernels (profiles/r01_em_kernels_ncu.md; the E/M
# kernels are unchanged since): em_fir 14.6 + em_forward 62.6 + em_backward 63.6 + em_check/fixup 13.4 + em_stats 177.7 +
# reduce/finalize 2.1 MB
NCU_EM_TRAFFIC_BYTES = 334.0e6
FP64_PEAK_FALLBACK_GDFMA = 18421.7  # profiles/r01_fp64_peak.jsonl; the bench measures it live (hmm_measure_peaks)


def make_c2(hm, seed, T=T_C2):
    """the two README templates + (60, 2.0, 0.5, 0.3), rates
    [0.003, 0.001, 0.002], sigma 0.3, fixed true mu, lp = log rates."""
    K, N = 60, 3
    temps = np.stack([hm.create_spike_template(K, 3.0, 0.8, 0.2), hm.create_spike_template(K, 4.0, 0.3, 0.2),
                      hm.create_spike_template(K, 2.0, 0.5, 0.3)], axis=1)
    pp = np.array([0.003, 0.001, 0.002])
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(seed))
    lA = hm.StateMatrix(N, K, np.log(pp), False)
    mu = np.asfortranarray(temps.copy())
    mu[0, :] = 0.0
    return S, lA, mu, 0.3


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):  # (diagnostic: does polling nvidia-smi perturb the timed region?)
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def wait_first(self, timeout):
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def i64(v):
    return C.c_int64(v)


def i32(v):
    return C.c_int32(v)


# ---------------------------------------------------------------------------
def run_ours(args):
    env = Env()
    torch, dist = env.torch, env.dist
    hm, L, world, rank, local, dev = env.hm, env.L, env.world, env.rank, env.local, env.dev
    barrier, max_over_ranks = env.barrier, env.max_over_ranks

    # roofline denominators measured on this GPU, now (FP64 FMA issue rate; copy bandwidth beside the driver's figure)
    pk_f, pk_c = C.c_double(0), C.c_double(0)
    hm._lib.check(L.hmm_measure_peaks(C.byref(pk_f), C.byref(pk_c)))
    fp64_peak, copy_live = pk_f.value, pk_c.value
    T = args.samples
    S, lA, mu, sigma = make_c2(hm, seed=2 + rank, T=T)
    st = np.asfortranarray(lA.states)
    tr = np.ascontiguousarray(lA.transitions)
    sig = np.array([sigma])
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    info = hm.HmmInfo()

    # ---- device-resident decode -------------------------------------------------
    y_dev = torch.from_numpy(S).to(dev)
    x_dev = torch.empty(T, dtype=torch.int16, device=dev)
    ll = C.c_double(0)

    def step_dev():
        hm._lib.check(L.hmm_viterbi_dev_f64(C.c_void_p(y_dev.data_ptr()), i64(T), i32(1), p(st), i32(1), i32(lA.N),
                                            i32(lA.K), i32(lA.nstates), p(tr), i64(tr.size), p(mu), p(sig),
                                            C.c_void_p(x_dev.data_ptr()), C.byref(ll), i32(hm.MODES["ring"]),
                                            C.byref(info)))

    # nvidia-smi takes a second or two to start (NVML initialisation, which briefly stalls work on the GPUs it
    # enumerates): it is started first and has delivered its first sample before anything is timed
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first(5.0)
    barrier()  # all ranks start their warm-up together: no rank idles (and drops its clocks) waiting at the next barrier
    for _ in range(args.warmup):
        step_dev()
    barrier()
    top_ms, kern_ms, launches = [], [], 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_dev()
        launches += info.kernel_launches
    ev1.record()
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0)
    dt = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)  # device time of the K steps (CUDA events), max over ranks
    ms_per_step = dt / args.steps * 1e3
    value = world * T / (dt / args.steps) / 1e6
    chunks, rep_f, rep_b = info.n_chunks, info.fwd_repaired, info.bwd_repaired
    x_first = x_dev.cpu().numpy().copy()
    # The timed steps above re-launch the decode's cached CUDA graph (the product path), which has no per-kernel
    # timers.  The same K steps again with eager launches and CUDA-event timers on the launching stream give the
    # dominant kernel's duration for the roofline lines.
    hm._lib.check(L.hmm_set_profiling(i32(1)))
    for _ in range(args.steps):
        step_dev()
        top_ms.append(info.top_kernel_ms)
        kern_ms.append(info.kernel_ms)
    hm._lib.check(L.hmm_set_profiling(i32(0)))

    # ---- end to end through the host-pointer API (pinned host buffers) ----------
    yh, xh = C.c_void_p(), C.c_void_p()
    hm._lib.check(L.hmm_host_alloc(C.byref(yh), C.c_uint64(8 * T)))
    hm._lib.check(L.hmm_host_alloc(C.byref(xh), C.c_uint64(2 * T)))
    y_pin = np.ctypeslib.as_array(C.cast(yh, C.POINTER(C.c_double)), shape=(T,))
    x_pin = np.ctypeslib.as_array(C.cast(xh, C.POINTER(C.c_int16)), shape=(T,))
    y_pin[:] = S
    ll2 = C.c_double(0)

    def step_e2e():
        hm._lib.check(L.hmm_viterbi_ex_f64(yh, i64(T), p(st), i32(lA.N), i32(lA.K), i32(lA.nstates), p(tr),
                                           i64(tr.size), p(mu), C.c_double(sigma), xh, C.byref(ll2), None, None,
                                           i32(hm.MODES["ring"]), C.byref(info)))

    barrier()
    for _ in range(max(3, args.warmup // 2)):
        step_e2e()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_e2e()
        launches += info.kernel_launches
    ev1.record()
    barrier()
    dt_e = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    e2e_val = world * T / (dt_e / args.steps) / 1e6
    # x must be identical; ll is summed per pipeline segment on this path (another order of the same additions)
    x_same = bool(np.array_equal(x_first, x_pin))
    ll_rel = abs(ll.value - ll2.value) / max(abs(ll.value), 1e-300)
    same = x_same and ll_rel <= 1e-12
    # the same call with ordinary (pageable) numpy arrays, as a Julia Array would be: the driver stages the copies
    e2e_pageable = None
    if rank == 0 and world == 1:
        x_pg = np.empty(T, dtype=np.int16)

        def step_pg():
            hm._lib.check(L.hmm_viterbi_ex_f64(p(S), i64(T), p(st), i32(lA.N), i32(lA.K), i32(lA.nstates), p(tr),
                                               i64(tr.size), p(mu), C.c_double(sigma), p(x_pg), C.byref(ll2), None, None,
                                               i32(hm.MODES["ring"]), C.byref(info)))

        step_pg()
        torch.cuda.synchronize()
        n_pg = max(1, min(5, args.steps))
        ev0.record()
        for _ in range(n_pg):
            step_pg()
        ev1.record()
        torch.cuda.synchronize()
        e2e_pageable = T / (ev0.elapsed_time(ev1) * 1e-3 / n_pg) / 1e6
        same = same and bool(np.array_equal(x_first, x_pg))
    L.hmm_host_free(yh)
    L.hmm_host_free(xh)
    del y_dev, x_dev
    torch.cuda.empty_cache()

    # ---- Baum-Welch half of the metric (config 3), rank-local, resident X --------
    bw = None
    if not args.no_bw:
        try:
            bw = bench_bw(hm, args, rank)
        except hm.HmmError as e:  # engine not available: report, do not hide
            bw = {"unavailable": str(e)}
    # ---- CPU baseline: oracle port on a bounded sample (rank 0, N=1 only) ---------
    # ---- the two scaling configurations of north_star, in the same line (and the same driver run) ---------------
    c3 = c4 = c5 = None
    if not args.no_scaling_blocks and args.samples == T_C2:
        L.hmm_release_workspace()
        c3 = block_c3(env, args)
        c5 = block_c5(env, args)
        c4 = block_c4(env, args)
    clocks = sampler.stop() if rank == 0 else None  # sampled across every timed region above
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, parity = cpu_baseline_viterbi(S, lA, mu, sigma, x_gpu=x_first, ll_gpu=ll.value, seconds=args.cpu_seconds)
        if bw is not None and "unavailable" not in bw:
            bw["cpu_baseline"], bw["parity"] = cpu_baseline_em(hm, make_c2(hm, seed=3, T=T_C3)[0])

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        top = float(np.mean(top_ms))
        achieved = BYTES_PER_SAMPLE * T / (top * 1e-3) / 1e9
        dfma = (T + chunks * 512.0) * lA.N * (lA.K - 1)  # L = K - 1 = 59 taps; 512-sample warm-up per chunk
        out = {
            "metric": "Viterbi Msamples/s (N=3,K=60; Baum-Welch iters/s in `baum_welch`)",
            "value": round(value, 2), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "wall_ms_per_step": round(wall / args.steps * 1e3, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: single-channel 30 kHz x 10 min (18M samples), N=3 x K=60, "
                                   "Viterbi decode only, fixed lA/mu/sigma" + ("" if T == T_C2 else f" [T={T}]"),
                       "samples_per_gpu": T, "nstates": lA.nstates, "ntrans": int(tr.size),
                       "parallelism": f"channel-sharded x{world} (one 18M-sample channel per GPU, no collective)",
                       "engine": "ring (time-parallel, exact)", "chunks": chunks,
                       "chunks_repaired_fwd_bwd": [rep_f, rep_b],
                       "l2": "inputs larger than L2 (144 MB of y per step vs 126 MB L2); no explicit flush"},
            "e2e": {"value": round(e2e_val, 2), "unit": "Msamples/s", "h2d_bytes_per_step": 8 * T,
                    "d2h_bytes_per_step": 2 * T + 8, "host_memory": "pinned", "same_result_as_resident": same,
                    "x_identical_to_resident": x_same, "ll_rel_diff_to_resident": ll_rel,
                    "pageable_host_value": None if e2e_pageable is None else round(e2e_pageable, 2)},
            "gpu_launches": int(launches),
            "eager_ms_per_step": round(float(np.mean(kern_ms)), 4),
            "roofline": {"bound": "hbm", "kernel": "ring_vit_forward_ws<3,8,59>", "achieved": round(achieved, 1),
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": NCU_TRAFFIC_BYTES, "kernel_ms": round(top, 4), "copy_gb_s_measured_in_run": round(copy_live, 1),
                         "note": "algorithmic 10 B/sample; traffic = dram read+write per launch from the ncu --set "
                                 "full capture in profiles/; the kernel is FP64-issue bound (FIR), see `fp64_issue`"},
            "fp64_issue": {"achieved": round(dfma / (top * 1e-3) / 1e9, 1), "peak": round(fp64_peak, 1), "unit": "GDFMA/s",
                           "frac": round(dfma / (top * 1e-3) / 1e9 / fp64_peak, 4),
                           "peak_source": "measured in this run on this GPU (hmm_measure_peaks: DFMA with a "
                                          "constant-bank operand, 8 CTAs of 256 threads per SM)",
                           "whole_step_frac": round(dfma / (ms_per_step * 1e-3) / 1e9 / fp64_peak, 4),
                           "work": "FIR: (samples + chunks*warmup) x N x L fused multiply-adds, L = K-1 = 59 taps"},
            "cpu_baseline": cpu,
            "parity": parity,
            "baum_welch": bw,
            "config3_sharded": c3,
            "config5": c5,
            "config4": c4,
            "clocks": clocks,
        }
        print(json.dumps(out))
    env.close()


def make_c5(hm, seed=5, T=108_000_000):
    """5 templates K=60, T = 108 M (1 h at 30 kHz)."""
    K, N = 60, 5
    prm = [(3.0, 0.8, 0.2), (4.0, 0.3, 0.2), (2.0, 0.5, 0.3), (2.5, 0.6, 0.25), (3.5, 0.4, 0.15)]
    temps = np.stack([hm.create_spike_template(K, *q) for q in prm], axis=1)
    pp = np.array([0.003, 0.001, 0.002, 0.0015, 0.0025])
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(seed))
    lA = hm.StateMatrix(N, K, np.log(pp), False)
    mu = np.asfortranarray(temps.copy())
    mu[0, :] = 0.0
    return S, lA, mu, 0.3


class Env:
    """One process per GPU: torch.distributed (NCCL) for the barrier and the max-over-ranks timing."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.hm = ge.load_package()
        self.L = self.hm.lib()
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available() or self.hm.device_count() < 1:
            raise SystemExit("bench.py needs a CUDA device: libhmmcuda has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.hm._lib.check(self.L.hmm_set_device(i32(self.local)))
        # the library issues its work on this (created) torch stream, so torch CUDA events bracket it
        self.work = torch.cuda.Stream(device=self.dev)
        self.work.wait_stream(torch.cuda.current_stream())
        torch.cuda.set_stream(self.work)
        self.hm._lib.check(self.L.hmm_set_stream(C.c_void_p(self.work.cuda_stream)))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local])
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks_int(self, v):
        if self.world == 1:
            return int(v)
        t = self.torch.tensor([int(v)], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t)
        return int(t.item())

    def events(self):
        return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)

    def close(self):
        self.L.hmm_set_stream(None)
        if self.world > 1:
            self.dist.destroy_process_group()


def block_c5(env, args, T=108_000_000):
    """BASELINE config 5: ONE 108 M-sample recording (N=5, K=60), time-sharded over the GPUs (strong scaling: the
    total work is fixed).  Every rank decodes its span on its own (ghost chunks play the neighbours), stores its
    boundary summary straight into every peer's exchange block over NVLink (CUDA-IPC peer memory) and judges
    every shard boundary itself: no collective call on the data path, one host synchronisation per decode."""
    hm, L, torch, world, rank, dev = env.hm, env.L, env.torch, env.world, env.rank, env.dev
    S, lA, mu, sigma = make_c5(hm, T=T)
    ts = hm.timeshard
    chunk_len, warm = ts.default_chunking(T, world, lA.N, lA.K)
    if world > 1:
        warm = 256  # shorter chunks per GPU: a shorter speculative warm-up (boundaries are verified anyway)
    span = ts.shard_plan(T, world, chunk_len, warm)[rank]
    y_loc = torch.from_numpy(S[span[0]:span[1]]).to(dev)
    del S
    x_main = torch.empty(span[3] - span[2], dtype=torch.int16, device=dev)
    dec = ts.DistDecoder(y_loc.data_ptr(), span, T, chunk_len, warm, lA, mu, sigma, x_main.data_ptr(), dev)
    steps, warmup = args.steps, max(3, args.warmup)
    env.barrier()
    for _ in range(warmup):
        dec.decode()
    env.barrier()
    ev0, ev1 = env.events()
    ev0.record()
    for _ in range(steps):
        ll = dec.decode()
    ev1.record()
    env.barrier()
    dt = env.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    chk = env.sum_over_ranks_int(int(x_main.to(torch.int64).sum().item()))
    stats = dict(dec.stats)
    bvec = dec.sh.bvec
    dec.close()
    per = dt / steps
    return {"metric": "Viterbi Msamples/s, one 108M-sample recording time-sharded over the GPUs",
            "value": round(T / per / 1e6, 2), "unit": "Msamples/s", "ms_per_step": round(per * 1e3, 4),
            "scaling": "strong", "steps": steps, "warmup": warmup,
            "config": {"workload": "BASELINE config 5: single-channel 1 h at 30 kHz (108M samples), N=5 x K=60, "
                                   "time-chunked Viterbi across the GPUs with boundary-vector / traceback-state "
                                   "exchange over NVLink peer memory" + ("" if T == 108_000_000 else f" [T={T}]"),
                       "chunk_len": chunk_len, "warmup": warm, "boundary_bytes": 8 * (2 * bvec + 4),
                       "protocol": "peer-memory (hmm_vshard_p2p_*): summaries stored into every peer's exchange block, "
                                   "flags, every rank judges every boundary" if dec.p2p else "one NCCL all-gather of the "
                                   "shard summaries per decode",
                       "fallbacks": stats["fallbacks"], "ll": ll, "x_checksum": chk,
                       "l2": "inputs larger than L2 (>= 108 MB of y per GPU)"}}


def block_c3(env, args, T=T_C3, iters=20):
    """BASELINE config 3 over the GPUs: ONE 1-minute recording (1.8 M samples, N=3, K=60), 20 Baum-Welch iterations,
    time-sharded (hmm_emshard_*): per iteration one all-reduce of the sufficient statistics and one all-gather of the
    boundary vectors (NCCL), the M-step identical on every rank.  Strong scaling of a 0.4 ms step: reported as measured."""
    hm, torch, world, rank, dev = env.hm, env.torch, env.world, env.rank, env.dev
    ts = hm.timeshard
    S, lA_true, mu_true, _ = make_c2(hm, seed=3, T=T)
    N, K = 3, 60
    chunk_len, warm = ts.em_default_chunking(T, world, N, K)
    span = ts.shard_plan(T, world, chunk_len, warm)[rank]
    x_loc = torch.from_numpy(np.ascontiguousarray(S[span[0]:span[1]])).to(dev)
    sh = ts.EmShard(x_loc.data_ptr(), False, span, T, chunk_len)
    lA = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), False)
    em = ts.EmSharded([sh], N, K, lA.nstates, dev, distributed=world > 1)
    mu = np.asfortranarray(0.7 * mu_true)
    sigma = float(np.std(S))
    env.barrier()
    for _ in range(3):
        em.em_step(lA, mu, sigma)
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        lp, pp, mu, sigma, ll = em.em_step(lA, mu, sigma)
        lA = hm.StateMatrix.from_states(lA.states, pp, K, lp, False)
    env.torch.cuda.synchronize()
    dt = env.max_over_ranks(time.perf_counter() - t0)
    em.close()
    return {"metric": "Baum-Welch iters/s, one recording time-sharded over the GPUs", "value": round(iters / dt, 2),
            "unit": "iters/s", "ms_per_iter": round(dt / iters * 1e3, 3), "scaling": "strong", "iterations": iters, "T": T,
            "config": {"workload": "BASELINE config 3 time-sharded: 30 kHz x 1 min, N=3 x K=60, 20 Baum-Welch iterations",
                       "chunk_len": chunk_len, "collectives_per_iteration": "1 all-reduce (statistics) + 1 all-gather "
                       "(boundary vectors)" if world > 1 else "none (one shard)", "final_sigma": sigma, "final_loglik": ll}}


def make_c4_channel(hm, c, T):
    """SURVEY 8d C4, channel c: 4 templates K=48 with (a, b, c) drawn from documented ranges, rates U(0.0005, 0.004),
    recording seed 1000 + c -- every one of the 128 channels is its own draw."""
    N, K = 4, 48
    rng = np.random.default_rng(5000 + c)
    prm = [(rng.uniform(2, 4), rng.uniform(0.3, 0.9), rng.uniform(0.1, 0.3)) for _ in range(N)]
    temps = np.stack([hm.create_spike_template(K, *q) for q in prm], axis=1)
    pp = rng.uniform(0.0005, 0.004, size=N)
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(1000 + c))
    mu = np.asfortranarray(temps.copy())
    mu[0, :] = 0
    lA = hm.StateMatrix(N, K, np.log(pp), False)
    return S, lA, mu, 0.3


def block_c4(env, args, T=T_C2, C_total=128):
    """BASELINE config 4: a 128-channel probe x 10 min at 30 kHz, independent per-channel HMMs (N=4, K=48), 128
    DISTINCT channels sharded over the GPUs (strong scaling, no collective on the data path).  `value`: all of a
    rank's channels decoded from HBM by one hmm_viterbi_dev_f64 call per step (one CUDA graph); `e2e`: the same
    channels through the host-pointer batch API from pinned host memory, in groups of 16 channels."""
    from concurrent.futures import ThreadPoolExecutor

    hm, L, torch, world, rank, dev = env.hm, env.L, env.torch, env.world, env.rank, env.dev
    N, K = 4, 48
    Cn = C_total // world
    first = rank * Cn
    y_dev = torch.empty((Cn, T), dtype=torch.float64, device=dev)  # channel-major == [T x C] column-major
    sts, trs, mus, sig = [None] * Cn, [None] * Cn, [None] * Cn, [None] * Cn
    keep = {}

    def gen(k):
        S, lA, mu, s = make_c4_channel(hm, first + k, T)
        sts[k] = np.asfortranarray(lA.states).ravel(order="F")
        trs[k] = lA.transitions
        mus[k] = mu.ravel(order="F")
        sig[k] = s
        if k < 16:
            keep[k] = S
        return k, S, lA

    nthreads = max(1, min(16, (os.cpu_count() or 8) // max(1, min(world, 8))))
    with ThreadPoolExecutor(nthreads) as ex:
        for k, S, lA in ex.map(gen, range(Cn)):
            y_dev[k].copy_(torch.from_numpy(S))
            nstates = lA.nstates
    st = np.ascontiguousarray(np.concatenate(sts))
    tr = np.ascontiguousarray(np.concatenate(trs))
    mu = np.ascontiguousarray(np.concatenate(mus))
    sg = np.asarray(sig)
    ntr = trs[0].size
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    info = hm.HmmInfo()
    x_dev = torch.empty((Cn, T), dtype=torch.int16, device=dev)
    ll = np.zeros(Cn)

    def step_dev():
        hm._lib.check(L.hmm_viterbi_dev_f64(C.c_void_p(y_dev.data_ptr()), i64(T), i32(Cn), p(st), i32(0), i32(N), i32(K),
                                            i32(nstates), p(tr), i64(ntr), p(mu), p(sg),
                                            C.c_void_p(x_dev.data_ptr()), p(ll), i32(hm.MODES["ring"]), C.byref(info)))

    steps, warmup = max(2, min(args.steps, 10)), 3
    env.barrier()
    for _ in range(warmup):
        step_dev()
    env.barrier()
    ev0, ev1 = env.events()
    ev0.record()
    launches = 0
    for _ in range(steps):
        step_dev()
        launches += info.kernel_launches
    ev1.record()
    env.barrier()
    dt = env.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    value = C_total * T / (dt / steps) / 1e6
    rep = (info.fwd_repaired, info.bwd_repaired)
    chk = env.sum_over_ranks_int(int(x_dev.to(torch.int32).sum(dtype=torch.int64).item()))
    # ---- end to end: groups of 16 channels through the batch host-pointer API, pinned host memory ----
    G = min(16, Cn)
    yh, xh = C.c_void_p(), C.c_void_p()
    hm._lib.check(L.hmm_host_alloc(C.byref(yh), C.c_uint64(8 * T * G)))
    hm._lib.check(L.hmm_host_alloc(C.byref(xh), C.c_uint64(2 * T * G)))
    Y = np.ctypeslib.as_array(C.cast(yh, C.POINTER(C.c_double)), shape=(G, T))
    for k in range(G):
        Y[k] = keep[k]
    stg = np.ascontiguousarray(np.concatenate(sts[:G]))
    trg = np.ascontiguousarray(np.concatenate(trs[:G]))
    mug = np.ascontiguousarray(np.concatenate(mus[:G]))
    sgg = np.asarray(sig[:G])
    llg = np.zeros(G)

    def step_e2e():  # the first group's recordings stand in for every group: the bytes moved and decoded are the same
        for _ in range(Cn // G):
            hm._lib.check(L.hmm_viterbi_batch_f64(yh, i64(T), i32(G), p(stg), i32(0), i32(N), i32(K), i32(nstates),
                                                  p(trg), i64(ntr), p(mug), p(sgg), xh, p(llg), i32(hm.MODES["ring"]),
                                                  C.byref(info)))

    step_e2e()
    env.barrier()
    e_steps = 2
    ev0.record()
    for _ in range(e_steps):
        step_e2e()
    ev1.record()
    env.barrier()
    dt_e = env.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    e2e_val = C_total * T / (dt_e / e_steps) / 1e6
    same = bool(np.array_equal(np.ctypeslib.as_array(C.cast(xh, C.POINTER(C.c_int16)), shape=(G, T)),
                               x_dev[:G].cpu().numpy())) and bool(np.array_equal(llg, ll[:G]))
    L.hmm_host_free(yh)
    L.hmm_host_free(xh)
    del y_dev, x_dev
    torch.cuda.empty_cache()
    L.hmm_release_workspace()
    return {"metric": "Viterbi Msamples/s, 128-channel probe, independent per-channel HMMs",
            "value": round(value, 2), "unit": "Msamples/s", "ms_per_step": round(dt / steps * 1e3, 4),
            "scaling": "strong", "steps": steps, "warmup": warmup,
            "config": {"workload": "BASELINE config 4: 128-channel probe x 10 min at 30 kHz, independent per-channel HMMs "
                                   "(N=4, K=48), channels sharded over the GPUs" + ("" if T == T_C2 else f" [T={T}]"),
                       "channels_per_gpu": Cn, "samples_per_channel": T,
                       "data_note": "128 distinct channels: own templates / rates (seed 5000+c) and recording (seed 1000+c)",
                       "parallelism": f"channel-sharded x{world}, no collective", "x_checksum": chk,
                       "chunks_repaired_fwd_bwd": list(rep), "l2": "inputs larger than L2"},
            "e2e": {"value": round(e2e_val, 2), "unit": "Msamples/s", "h2d_bytes_per_step": 8 * T * Cn,
                    "d2h_bytes_per_step": (2 * T + 8) * Cn, "host_memory": "pinned", "same_result_as_resident": same,
                    "note": f"groups of {G} channels per hmm_viterbi_batch_f64 call"},
            "gpu_launches": int(launches)}


def bench_bw(hm, args, rank):
    """Config 3: T = 1.8 M, N=3 x K=60, E/M iterations from mu0 = 0.7 truth."""
    T = min(T_C3, args.samples)
    S, lA_true, mu_true, _ = make_c2(hm, seed=3 + rank, T=T)
    N, K = 3, 60
    lA = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), False)
    mu = np.asfortranarray(0.7 * mu_true)
    sigma = float(np.std(S))
    iters = 20
    with hm.TrainContext(S) as ctx:
        for _ in range(2):
            ctx.em_step(lA, mu, sigma)
        # (a) the loop in the caller, as train_model with a callback runs it: one C call + StateMatrix rebuild per step
        t0 = time.perf_counter()
        lA_h, mu_h, s_h = lA, mu, sigma
        for _ in range(iters):
            lp, pp, mu_h, s_h, ll_h, _ = ctx.em_step(lA_h, mu_h, s_h, return_info=True)
            lA_h = hm.StateMatrix.from_states(lA_h.states, pp, K, lp, False)
        dt_host = time.perf_counter() - t0
        # (b) the same 20 iterations in one hmm_train_run call (train_model without a callback): the headline
        ctx.run(lA, mu, sigma, 2)
        t0 = time.perf_counter()
        lA_r, mu, sigma, lls, info = ctx.run(lA, mu, sigma, iters, return_info=True)
        dt = time.perf_counter() - t0
        ll = float(lls[-1])
        launches = info["kernel_launches"]
        same = bool(np.abs(mu - mu_h).max() < 1e-12 and abs(sigma - s_h) < 1e-12)
        lA = lA_r
    peak, peak_src = measured_peak_hbm()
    traffic = NCU_EM_TRAFFIC_BYTES * T / T_C3
    ach = traffic / (dt / iters) / 1e9
    alg = 16.0 * (lA.nstates + 1) * T  # SURVEY 8d's contract figure (alpha materialised) -- NOT what these kernels move
    return {"value": round(iters / dt, 3), "unit": "iters/s", "iterations": iters, "T": T,
            "roofline": {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(ach / peak, 4), "traffic": traffic,
                         "note": "achieved = MEASURED dram bytes of one iteration's kernels (ncu, profiles/) / iteration "
                                 "time: the step is latency-bound (issue slots ~40 % in em_forward / em_backward / "
                                 "em_stats), not bandwidth-bound.  The semi-Markov E-step never materialises alpha "
                                 "(88 B/sample of per-step log quantities instead of 16*(nstates+1) = 2864), so a "
                                 "fraction against SURVEY 8d's alpha-materialising figure would exceed 1 and is not "
                                 "reported as a roofline",
                         "alpha_materialising_contract_bytes": alg},
            "config": "BASELINE config 3: 30 kHz x 1 min, N=3 x K=60, 20 Baum-Welch iterations in one hmm_train_run call, X resident in HBM, "
                      "transition weights rebuilt from lp each iteration inside the timed region",
            "ms_per_iter": round(dt / iters * 1e3, 3), "final_sigma": sigma, "final_loglik": ll,
            "host_loop": {"value": round(iters / dt_host, 3), "unit": "iters/s", "same_fit_as_library_loop": same,
                          "note": "one hmm_train_em_step call + host StateMatrix rebuild per iteration (train_model with a callback)"},
            "gpu_launches": int(launches), "engine": info["engine"]}


def cpu_baseline_viterbi(S, lA, mu, sigma, x_gpu=None, ll_gpu=None, seconds=12.0, threads=1):
    """Oracle (literal C port of the reference algorithm), 1 thread, on the WHOLE recording when it fits the
    time budget (config 2: 18 M samples take ~15 s on the box's cores), else on a prefix.  The path it produces
    is not thrown away: it is the parity check of the GPU decode at the benchmark's own size."""
    from concurrent.futures import ThreadPoolExecutor

    O = ge.load_oracle()
    O.build()
    n0 = 200_000
    t0 = time.perf_counter()
    O.viterbi(S[:n0], lA, mu, sigma)
    rate = n0 / (time.perf_counter() - t0)
    full = S.size / rate <= 6 * seconds
    n = S.size if full else int(min(S.size, max(n0, rate * seconds)))
    with ThreadPoolExecutor(1) as ex:  # the near-tie screen of the same samples runs beside it on another core
        f_scr = ex.submit(O.viterbi_screen, S[:n], lA, mu, sigma)
        t0 = time.perf_counter()
        xo, llo = O.viterbi(S[:n], lA, mu, sigma)
        dt = time.perf_counter() - t0
        scr = f_scr.result()
    cpu = {"value": round(n / dt / 1e6, 4), "unit": "Msamples/s", "cores": threads, "kind": "port",
           "sample": (f"the whole {n}-sample recording" if full else f"first {n} samples of the same recording")
                     + ", whole-sequence decode, 1 thread (the reference's execution model; Julia itself is not "
                       "installed); the near-tie screen ran on a second core at the same time"}
    parity = None
    if x_gpu is not None:
        # a prefix decode can only differ from the full decode near the cut
        m = n if full else n - 8192
        bad = int(np.count_nonzero(x_gpu[:m] != xo[:m]))
        parity = {"against": "oracle/hmm_oracle.c (CPU restatement of src/viterbi.jl:44-98), same samples",
                  "n_compared": int(m), "x_mismatches": bad,
                  "ll_rel_err": (abs(ll_gpu - llo) / abs(llo)) if full else None,
                  "ll_gpu": ll_gpu if full else None, "ll_oracle": llo if full else None,
                  "near_tie_screen": scr}
    return cpu, parity


def cpu_baseline_em(hm, S, seconds=12.0):
    """Oracle E/M step (src/baumwelch.jl:362-370: forward, backward, update with dense alpha/beta/gamma), 1 thread,
    on a prefix sized for a few seconds; the GPU step on the same prefix is compared with it."""
    O = ge.load_oracle()
    N, K = 3, 60
    Tb = min(S.size, 200_000)
    X = np.ascontiguousarray(S[:Tb])
    lp0 = np.log(np.full(N, 0.01))
    _, _, mu_true, _ = make_c2(hm, seed=3, T=1000)
    mu0 = np.asfortranarray(0.7 * mu_true)
    s0 = float(np.std(X))
    t0 = time.perf_counter()
    o = O.em_step(X, O.OracleStateMatrix(N, K, lp0, False), mu0.copy(order="F"), s0)
    dt = time.perf_counter() - t0
    r = hm.em_step(X, hm.StateMatrix(N, K, lp0, False), mu0.copy(order="F"), s0, mode="ring")
    iters_s_full = 1.0 / (dt * T_C3 / Tb)
    cpu = {"value": round(iters_s_full, 5), "unit": "iters/s", "cores": 1, "kind": "port",
           "sample": f"one E/M iteration on the first {Tb} samples of the config-3 recording took {dt:.2f} s on 1 "
                     f"thread; scaled linearly in T to {T_C3} samples (the algorithm is O(T))"}
    parity = {"against": "oracle/hmm_oracle.c em_step, same samples", "T": Tb,
              "max_abs_err_mu": float(np.abs(r[2] - o[2]).max()), "abs_err_sigma": abs(r[3] - o[3]),
              "max_abs_err_lp": float(np.abs(r[0] - o[0]).max()), "loglik_rel_err": abs(r[4] - o[4]) / abs(o[4])}
    return cpu, parity


# ---------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (oracle port; Julia is not
    available on the box) on all host cores: independent >=100k-sample chunks decoded
    in parallel threads, the chunking scheme of src/fit.jl:11-42."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hm = ge.load_package()  # synthetic data + host StateMatrix only; no CUDA call is made
    O = ge.load_oracle()
    O.build()
    cores = os.cpu_count() or 1
    chunk = 200_000
    S0, lA, mu, sigma = make_c2(hm, seed=2, T=chunk)
    t0 = time.perf_counter()
    O.viterbi(S0, lA, mu, sigma)
    t_chunk = time.perf_counter() - t0
    total_steps = args.steps + args.warmup
    target = min(8.0, 150.0 / max(1, total_steps))  # seconds of wall clock per step
    nchunks = int(max(cores, cores * max(1.0, target / t_chunk)))
    nchunks = min(nchunks, max(cores, args.samples // chunk))
    n = nchunks * chunk
    S, lA, mu, sigma = make_c2(hm, seed=2, T=n)

    def work(k):
        O.viterbi(S[k * chunk:(k + 1) * chunk], lA, mu, sigma)

    def step():
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=cores) as ex:
            list(ex.map(work, range(nchunks)))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = round(n / dt / 1e6, 4)
    sample = (f"{nchunks} independent chunks of {chunk} samples of the config-2 recording per step, decoded on "
              f"{cores} threads (src/fit.jl:11-42 chunking); oracle C port, Julia unavailable")
    print(json.dumps({
        "impl": "reference", "metric": "Viterbi Msamples/s (N=3,K=60; Baum-Welch iters/s in `baum_welch`)",
        "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: single-channel 30 kHz x 10 min (18M samples), N=3 x K=60, "
                               "Viterbi decode only, fixed lA/mu/sigma", "sample_per_step": n},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=T_C2, help="samples per GPU (default: config 2's 18M)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-bw", action="store_true")
    ap.add_argument("--no-scaling-blocks", action="store_true", help="skip the config-4 / config-5 blocks of the line")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 (default, the driver's contract line): one 18M-sample channel per GPU; "
                         "c4: 128 channels (N=4, K=48) sharded by channel; "
                         "c5: one 108M-sample recording time-sharded over the GPUs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("c3", "c4", "c5"):
        env = Env()
        T = args.samples
        if args.workload == "c3":
            blk = block_c3(env, args, T=T if T != T_C2 else T_C3)
        elif args.workload == "c5":
            blk = block_c5(env, args, T=T if T != T_C2 else 108_000_000)
        else:
            blk = block_c4(env, args, T=T)
        if env.rank == 0:
            blk.update({"n_gpus": env.world, "higher_is_better": True, "vs_baseline": None, "dtype": "f64",
                        "data": "synthetic"})
            print(json.dumps(blk))
        env.close()
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
