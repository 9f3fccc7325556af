"""Throughput of the time-parallel per-state engine beside the sequential one, on the reference's own test model
(overlap, 3 600 states, test/runtests.jl:24) and on a ring model; device-resident timings from hmm_info."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402

hm = ge.load_package()
out = {}
for name, K, overlap, T in [("overlap_N2_K60_3600states", 60, True, 2_000_000), ("overlap_N2_K24", 24, True, 4_000_000),
                            ("ring_N3_K60", 60, False, 4_000_000)]:
    N = 2 if overlap else 3
    pars = [(3.0, 0.8, 0.2), (4.0, 0.3, 0.2), (2.0, 0.5, 0.3)]
    temps = np.stack([hm.create_spike_template(K, *pars[i]) for i in range(N)], 1)
    pp = np.array([0.003, 0.001, 0.002][:N])
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(7))
    lA = hm.StateMatrix(N, K, np.log(pp), overlap)
    mu = np.asfortranarray(temps)
    row = {"T": T, "nstates": int(lA.nstates)}
    ref = None
    for mode in ("generic", "faithful"):
        Tm = T if mode == "generic" else min(T, 500_000)
        best = None
        for _ in range(3 if mode == "generic" else 1):
            t0 = time.perf_counter()
            x, ll, info = hm.viterbi(S[:Tm], lA, mu, 0.3, mode=mode, return_info=True)
            wall = time.perf_counter() - t0
            k = info["kernel_ms"]
            best = k if best is None else min(best, k)
        row[mode] = {"T": Tm, "kernel_ms": best, "Msamples_per_s": Tm / best / 1e3, "wall_s": wall,
                     "n_chunks": info["n_chunks"], "fwd_repaired": info["fwd_repaired"], "bwd_repaired": info["bwd_repaired"]}
        if mode == "generic":
            ref = x
        else:
            row["x_equal_on_prefix"] = bool(np.array_equal(ref[:Tm - 2000], x[:Tm - 2000]))
    out[name] = row
# the CLI's model sizes (src/hmmsort.jl:54, up to four templates with overlaps): beyond the sequential engine
for name, N, K, T in [("overlap_N3_K60_10621states", 3, 60, 1_000_000), ("overlap_N4_K60_21123states", 4, 60, 500_000)]:
    pars = [(3.0, 0.8, 0.2), (4.0, 0.3, 0.2), (2.0, 0.5, 0.3), (2.5, 0.6, 0.25)]
    temps = np.stack([hm.create_spike_template(K, *pars[i]) for i in range(N)], 1)
    pp = np.array([0.003, 0.001, 0.002, 0.0015][:N])
    S = hm.create_signal(T, 0.3, pp, temps, hm.make_rng(7))
    lA = hm.StateMatrix(N, K, np.log(pp), True)
    mu = np.asfortranarray(temps)
    best = None
    for _ in range(2):
        x, ll, info = hm.viterbi(S, lA, mu, 0.3, mode="auto", return_info=True)
        best = info["kernel_ms"] if best is None else min(best, info["kernel_ms"])
    out[name] = {"T": T, "nstates": int(lA.nstates), "engine": info["engine"],
                 "generic": {"kernel_ms": best, "Msamples_per_s": T / best / 1e3, "n_chunks": info["n_chunks"],
                             "fwd_repaired": info["fwd_repaired"], "bwd_repaired": info["bwd_repaired"]},
                 "spike_fraction": float(np.mean(x != 1))}
print(json.dumps(out))
