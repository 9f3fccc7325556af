"""Randomised parity sweep: random ring / overlap models, lengths, noise levels, firing rates and chunkings through
the ring, generic and sequential engines against the CPU oracle (x identical, ll within 1e-9), plus one E/M step
(1e-9 per step).  Usage: python tools/fuzz_parity.py [seconds] [seed]."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402


def run(budget=60.0, seed=1, em_only=False, max_cases=None):
    hm = ge.load_package()
    O = ge.load_oracle()
    O.build()
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    stats = {"cases": 0, "ring": 0, "generic": 0, "faithful": 0, "em": 0, "repaired_cases": 0, "failures": []}
    _sweep(hm, O, rng, t_end, stats, em_only, max_cases)
    return stats


def _sweep(hm, O, rng, t_end, stats, em_only, max_cases):
    def templates(N, K):
        return np.stack([hm.create_spike_template(K, rng.uniform(1.0, 4.5), rng.uniform(0.2, 0.9), rng.uniform(0.1, 0.4))
                         for _ in range(N)], axis=1)

    while time.time() < t_end and (max_cases is None or stats["cases"] < max_cases):
        overlap = rng.random() < 0.25
        if overlap:
            N, K = 2, int(rng.integers(4, 30))
        else:
            N, K = int(rng.integers(1, 8)), int(rng.integers(4, 98))
        T = int(rng.integers(2048, 120_000))
        sigma = float(rng.uniform(0.15, 0.8))
        rates = rng.uniform(0.0003, 0.02 if rng.random() < 0.3 else 0.004, size=N)
        temps = templates(N, K)
        S = hm.create_signal(T, sigma, rates, temps, hm.make_rng(int(rng.integers(1, 1 << 30))))
        mu = np.asfortranarray(temps * rng.uniform(0.7, 1.1))
        if rng.random() < 0.5:
            mu[0, :] = 0.0
        sig_m = sigma * float(rng.uniform(0.8, 1.3))
        lA = hm.StateMatrix(N, K, np.log(rates * rng.uniform(0.5, 2.0, size=N)), overlap)
        chunk = int(rng.choice([0, 0, 512, 1024, 2048, 4096]))
        warm = int(rng.choice([0, 0, 128, 256, 512]))
        case = dict(N=N, K=K, T=T, overlap=bool(overlap), sigma=sigma, chunk=chunk, warm=warm)
        try:
            xo, llo = (None, 0.0) if em_only else O.viterbi(S, lA, mu, sig_m)
            modes = [] if em_only else ["generic", "faithful"] if overlap or K < 4 else ["ring", "generic"]
            if T > 40_000 and "faithful" in modes:
                modes.remove("faithful")
            hm.set_ring_params(chunk, warm)
            # a third of the cases exercise the repair paths: forced flags, or a warm-up of zero (real mis-speculation)
            dbg = rng.random()
            os.environ.pop("HMMCUDA_DEBUG_FLAG_EVERY", None)
            os.environ.pop("HMMCUDA_DEBUG_WARMUP", None)
            if dbg < 0.2:
                os.environ["HMMCUDA_DEBUG_FLAG_EVERY"] = str(int(rng.integers(2, 6)))
            elif dbg < 0.33:
                os.environ["HMMCUDA_DEBUG_WARMUP"] = "0"
            # the generic engine's memory placements (tables / score columns in shared memory or in L2) and its two
            # traceback forms, as they occur for models of 10 000+ states
            os.environ.pop("HMMCUDA_DEBUG_GEN_SMEM_KB", None)
            os.environ.pop("HMMCUDA_DEBUG_GEN_DIRECT_TRACE", None)
            if rng.random() < 0.4:
                os.environ["HMMCUDA_DEBUG_GEN_SMEM_KB"] = str(int(rng.choice([3, 8, 16, 32, 64])))
                os.environ["HMMCUDA_DEBUG_GEN_DIRECT_TRACE"] = str(int(rng.integers(0, 2)))
            for mode in modes:
                x, ll, info = hm.viterbi(S, lA, mu, sig_m, mode=mode, return_info=True)
                ok = np.array_equal(x, xo) and abs(ll - llo) <= 1e-9 * abs(llo)
                stats[mode] += 1
                if info["fwd_repaired"] or info["bwd_repaired"]:
                    stats["repaired_cases"] += 1
                if not ok:
                    stats["failures"].append(dict(case, mode=mode, mismatches=int(np.sum(x != xo)), ll=ll, llo=llo, info=info))
            hm.set_ring_params(0, 0)
            for k in ("HMMCUDA_DEBUG_FLAG_EVERY", "HMMCUDA_DEBUG_WARMUP", "HMMCUDA_DEBUG_GEN_SMEM_KB", "HMMCUDA_DEBUG_GEN_DIRECT_TRACE"):
                os.environ.pop(k, None)
            # the same recording as time shards (hmm_vshard_*: ghost chunks, boundary exchange, verify rounds)
            if not em_only and "ring" in modes and T >= 30_000:
                n_sh = int(rng.integers(2, 6))
                try:
                    xs, lls = hm.viterbi_time_sharded(S, lA, mu, sig_m, n_sh)
                    stats["sharded"] = stats.get("sharded", 0) + 1
                    if not (np.array_equal(xs, xo) and abs(lls - llo) <= 1e-9 * abs(llo)):
                        stats["failures"].append(dict(case, mode=f"sharded x{n_sh}", mismatches=int(np.sum(xs != xo)), ll=lls, llo=llo))
                except hm.HmmArgumentError:
                    stats["sharded_refused"] = stats.get("sharded_refused", 0) + 1  # span too short for the plan
            if not overlap and T <= 60_000 and rng.random() < 0.5:
                mu0 = np.asfortranarray(mu.copy())
                mu0[0, :] = 0.0
                r = hm.em_step(S, lA, mu0.copy(order="F"), sig_m, mode="ring")
                o = O.em_step(S, lA, mu0.copy(order="F"), sig_m)
                rel = abs(r[4] - o[4]) / abs(o[4])
                if not np.isfinite(o[3]):  # the reference's own update broke down (a neuron's mass underflowed: 0/0)
                    stats["em_reference_nan"] = stats.get("em_reference_nan", 0) + 1
                    stats["cases"] += 1
                    continue
                # a neuron with n_i = T e^{lp_i} expected spikes: absolute posterior errors ~1e-12 per sample (the boundary
                # tolerance of the chunked E-step) weigh 1e-12 T / n_i on its statistics
                tol = 1e-9 + 1e-11 * np.exp(-np.minimum(o[0], 0.0))
                ok = bool(np.all(np.abs(r[0] - o[0]) <= tol) and np.all(np.nanmax(np.abs(r[2] - o[2]), axis=0) <= tol)
                          and abs(r[3] - o[3]) <= 1e-9 and rel < 1e-9)
                err = max(float(np.nanmax(np.abs(r[2] - o[2]))), abs(r[3] - o[3]), float(np.abs(r[0] - o[0]).max()))
                stats["em"] += 1
                if T >= 20_000:  # the same step over time shards (hmm_emshard_*), against the single-GPU step
                    import torch
                    ts = hm.timeshard
                    n_sh, cl = int(rng.integers(2, 5)), int(rng.choice([1024, 2048, 4096]))
                    dev = torch.device("cuda", 0)
                    em = None
                    try:
                        spans = ts.shard_plan(T, n_sh, cl, 256)
                        xs = [torch.from_numpy(np.ascontiguousarray(S[sp[0]:sp[1]])).to(dev) for sp in spans]
                        em = ts.EmSharded([ts.EmShard(xd.data_ptr(), False, sp, T, cl) for xd, sp in zip(xs, spans)], N, K,
                                          lA.nstates, dev)
                        q = em.em_step(lA, mu0.copy(order="F"), sig_m)
                        stats["em_sharded"] = stats.get("em_sharded", 0) + 1
                        oks = bool(np.all(np.abs(q[0] - r[0]) <= 10 * tol) and np.all(np.nanmax(np.abs(q[2] - r[2]), axis=0) <= 10 * tol)
                                   and abs(q[3] - r[3]) <= 1e-8 and abs(q[4] - r[4]) <= 1e-9 * abs(r[4]))
                        if not oks:
                            stats["failures"].append(dict(case, mode=f"em sharded x{n_sh} chunk {cl}",
                                                          lp_err=float(np.abs(q[0] - r[0]).max()),
                                                          mu_err=np.nanmax(np.abs(q[2] - r[2]), axis=0).tolist(),
                                                          sigma=[q[3], r[3]], ll=[q[4], r[4]], lp=r[0].tolist()))
                    except (hm.HmmError, RuntimeError, ValueError) as e:
                        stats.setdefault("em_sharded_refused", []).append(repr(e)[:120])
                    finally:
                        if em is not None:
                            em.close()
                if not ok:
                    dmu = np.abs(r[2] - o[2])
                    stats["failures"].append(dict(case, mode="em", err=err, ll_rel=rel, lp_gpu=r[0].tolist(), lp_oracle=o[0].tolist(),
                                                  sigma=[r[3], o[3]], mu_err_per_neuron=np.nanmax(dmu, axis=0).tolist(),
                                                  mu_nan_gpu=int(np.isnan(r[2]).sum()), mu_nan_oracle=int(np.isnan(o[2]).sum()),
                                                  mu_absmax_oracle=np.nanmax(np.abs(o[2]), axis=0).tolist(),
                                                  rates=rates.tolist()))
        except Exception as e:  # noqa: BLE001
            stats["failures"].append(dict(case, error=repr(e)))
            hm.set_ring_params(0, 0)
            for k in ("HMMCUDA_DEBUG_FLAG_EVERY", "HMMCUDA_DEBUG_WARMUP", "HMMCUDA_DEBUG_GEN_SMEM_KB", "HMMCUDA_DEBUG_GEN_DIRECT_TRACE"):
                os.environ.pop(k, None)
        stats["cases"] += 1


if __name__ == "__main__":
    print(json.dumps(run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 1,
                         len(sys.argv) > 3 and sys.argv[3] == "em")))
