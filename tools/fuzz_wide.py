"""Second randomised sweep, over the entry points tools/fuzz_parity.py does not reach: dense forward / backward (ring
and per-state engines), batches of channels, overlap models with three neurons, E/M steps of overlap models (dense
path), the library-side training loop against the oracle's loop, reconstruct_signal / unroll_mlseq, and -- with
`long` -- a few multi-million-sample recordings through the pipelined host-pointer decode.
Usage: python tools/fuzz_wide.py [seconds] [seed] [long]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402


def run(budget=60.0, seed=1, long=False, max_cases=None):
    hm = ge.load_package()
    O = ge.load_oracle()
    O.build()
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    st = {"cases": 0, "failures": []}

    def bump(k):
        st[k] = st.get(k, 0) + 1

    def fail(case, what, **kw):
        st["failures"].append(dict(case, what=what, **kw))

    def model(N, K, overlap, T, sigma=None):
        temps = np.stack([hm.create_spike_template(K, rng.uniform(1.0, 4.5), rng.uniform(0.2, 0.9), rng.uniform(0.1, 0.4))
                          for _ in range(N)], axis=1)
        sigma = float(rng.uniform(0.15, 0.8)) if sigma is None else sigma
        rates = rng.uniform(0.0005, 0.008, size=N)
        S = hm.create_signal(T, sigma, rates, temps, hm.make_rng(int(rng.integers(1, 1 << 30))))
        mu = np.asfortranarray(temps * rng.uniform(0.8, 1.1))
        mu[0, :] = 0.0
        lA = hm.StateMatrix(N, K, np.log(rates * rng.uniform(0.7, 1.5, size=N)), overlap)
        return S, lA, mu, sigma * float(rng.uniform(0.9, 1.2))

    def relerr(a, b):
        fin = np.isfinite(b)
        if not np.array_equal(np.isfinite(a), fin):
            return np.inf
        return float(np.max(np.abs(a[fin] - b[fin]) / (1.0 + np.abs(b[fin])))) if fin.any() else 0.0

    while time.time() < t_end and (max_cases is None or st["cases"] < max_cases):
        st["cases"] += 1
        try:
            if long:
                N, K = int(rng.integers(2, 6)), int(rng.choice([32, 48, 60, 80]))
                T = int(rng.integers(4_200_000, 9_000_000))
                S, lA, mu, sig = model(N, K, False, T, sigma=0.3)
                case = dict(kind="long", N=N, K=K, T=T)
                xo, llo = O.viterbi(S, lA, mu, sig)
                x, ll, info = hm.viterbi(S, lA, mu, sig, return_info=True)  # pageable host array: pipelined + staged
                bump("long")
                if not (np.array_equal(x, xo) and abs(ll - llo) <= 1e-9 * abs(llo)):
                    fail(case, "pipelined decode", mismatches=int(np.sum(x != xo)), ll=ll, llo=llo, info=info)
                xf, llf = hm.viterbi_f32(S.astype(np.float32), lA, mu, sig)  # FP32 recording, FP64 arithmetic
                xo32, llo32 = O.viterbi(S.astype(np.float32).astype(np.float64), lA, mu, sig)
                bump("long_f32_input")
                # (the Float32 entry points run the FIR in FP32: ll within 1e-4, x may differ at near-ties -- rate reported)
                st["f32_x_mismatch_rate_max"] = max(st.get("f32_x_mismatch_rate_max", 0.0), float(np.mean(xf != xo32)))
                if not (abs(llf - llo32) <= 1e-4 * abs(llo32) and np.mean(xf != xo32) < 1e-3):
                    fail(case, "f32 recording", mismatches=int(np.sum(xf != xo32)), ll=llf, llo=llo32)
                continue
            kind = rng.choice(["fb_ring", "fb_state", "batch", "overlap3", "em_overlap", "train_loop", "recon"])
            case = dict(kind=str(kind))
            if kind == "fb_ring":
                N, K, T = int(rng.integers(1, 6)), int(rng.integers(4, 70)), int(rng.integers(2048, 12000))
                S, lA, mu, sig = model(N, K, False, T)
                case.update(N=N, K=K, T=T)
                a, b = hm.forward(S, lA, mu, sig), hm.backward(S, lA, mu, sig)
                ao, bo = O.forward(S, lA, mu, sig), O.backward(S, lA, mu, sig)
                bump("fb_ring")
                ea, eb = relerr(a, ao), relerr(b, bo)
                if not (ea < 1e-9 and eb < 1e-9):
                    fail(case, "dense alpha/beta (ring)", alpha=ea, beta=eb)
            elif kind == "fb_state":
                N, K, T = 2, int(rng.integers(3, 12)), int(rng.integers(50, 3000))
                S, lA, mu, sig = model(N, K, True, T)
                case.update(N=N, K=K, T=T)
                a, b = hm.forward(S, lA, mu, sig), hm.backward(S, lA, mu, sig)
                ao, bo = O.forward(S, lA, mu, sig), O.backward(S, lA, mu, sig)
                bump("fb_state")
                ea, eb = relerr(a, ao), relerr(b, bo)
                if not (ea < 1e-11 and eb < 1e-11):
                    fail(case, "dense alpha/beta (per-state)", alpha=ea, beta=eb)
            elif kind == "batch":
                N, K, T, Cn = int(rng.integers(1, 6)), int(rng.integers(4, 70)), int(rng.integers(2048, 60000)), int(rng.integers(2, 6))
                chans = [model(N, K, False, T) for _ in range(Cn)]
                case.update(N=N, K=K, T=T, C=Cn)
                Y = np.asfortranarray(np.stack([c[0] for c in chans], axis=1))
                x, ll = hm.viterbi_batch(Y, [(c[1], c[2], c[3]) for c in chans])
                bump("batch")
                for c in range(Cn):
                    xo, llo = O.viterbi(chans[c][0], chans[c][1], chans[c][2], chans[c][3])
                    if not (np.array_equal(x[:, c], xo) and abs(ll[c] - llo) <= 1e-9 * abs(llo)):
                        fail(case, f"batch channel {c}", mismatches=int(np.sum(x[:, c] != xo)), ll=float(ll[c]), llo=llo)
            elif kind == "overlap3":
                N, K, T = 3, int(rng.integers(3, 11)), int(rng.integers(100, 40000))
                S, lA, mu, sig = model(N, K, True, T)
                case.update(N=N, K=K, T=T, nstates=int(lA.nstates))
                xo, llo = O.viterbi(S, lA, mu, sig)
                for mode in ("generic", "faithful" if T < 20000 else "auto"):
                    x, ll = hm.viterbi(S, lA, mu, sig, mode=mode)
                    if not (np.array_equal(x, xo) and abs(ll - llo) <= 1e-9 * abs(llo)):
                        fail(case, f"overlap3 {mode}", mismatches=int(np.sum(x != xo)), ll=ll, llo=llo)
                bump("overlap3")
            elif kind == "em_overlap":
                N, K, T = 2, int(rng.integers(3, 9)), int(rng.integers(200, 4000))
                S, lA, mu, sig = model(N, K, True, T)
                case.update(N=N, K=K, T=T)
                r = hm.em_step(S, lA, mu.copy(order="F"), sig)
                o = O.em_step(S, lA, mu.copy(order="F"), sig)
                bump("em_overlap")
                if np.isfinite(o[3]):
                    err = max(relerr(r[0], o[0]), relerr(r[2], o[2]), abs(r[3] - o[3]), abs(r[4] - o[4]) / abs(o[4]))
                    if not err < 1e-8:
                        fail(case, "E/M step, overlap model", err=err, lp=[r[0].tolist(), o[0].tolist()])
            elif kind == "train_loop":
                N, K, T, steps = int(rng.integers(1, 5)), int(rng.integers(8, 64)), int(rng.integers(3000, 25000)), 3
                S, lA, mu, sig = model(N, K, False, T)
                case.update(N=N, K=K, T=T)
                m_lib = mu.copy(order="F")
                lA_l, m_lib, s_lib = hm.train_model(S, lA, m_lib, sig, steps)
                lo, mo, so = lA, mu.copy(order="F"), sig
                for _ in range(steps):
                    lp, pp, mo, so, _ = O.em_step(S, lo, mo, so)
                    lo = hm.StateMatrix.from_states(lo.states, pp, K, lp, False)
                bump("train_loop")
                if np.isfinite(so):
                    err = max(relerr(m_lib, mo), abs(s_lib - so), relerr(lA_l.transitions["lp"], lo.transitions["lp"]))
                    if not err < 1e-6:
                        fail(case, "3 E/M steps in the library vs the oracle loop", err=err)
            else:
                N, K, T = int(rng.integers(1, 6)), int(rng.integers(3, 40)), int(rng.integers(1, 50000))
                overlap = bool(rng.random() < 0.4) and N <= 3 and K <= 12
                lA = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), overlap)
                mu = np.asfortranarray(rng.normal(size=(K, N)))
                x = rng.integers(1, lA.nstates + 1, size=T).astype(np.int16)
                case.update(N=N, K=K, T=T, overlap=overlap)
                bump("recon")
                if not np.array_equal(hm.reconstruct_signal(x, lA, mu, 0.3), O.reconstruct_signal(x, lA, mu)):
                    fail(case, "reconstruct_signal")
                if not np.array_equal(hm.unroll_mlseq(x, lA), O.unroll_mlseq(x, lA)):
                    fail(case, "unroll_mlseq")
        except Exception as e:  # noqa: BLE001
            fail(case if "case" in dir() else {}, "exception", error=repr(e)[:300])
    return st


if __name__ == "__main__":
    print(json.dumps(run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 1,
                         len(sys.argv) > 3 and sys.argv[3] == "long")))
