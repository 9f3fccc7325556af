"""Regenerates tests/golden/em_end_mass_case.npz: the E/M case of the randomised sweep (tools/fuzz_parity.py, seed 1)
in which a nearly silent neuron's posterior mass sits in a spike cut off by the end of the recording -- the case that
exposed the cancellation in the per-phase occupancies.  Replays the sweep's random stream on the CPU up to that case
(N=5, K=80, T=22477) and stores its inputs together with the CPU oracle's E/M step.  Run from the repository root."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as ge  # noqa: E402

hm = ge.load_package()
O = ge.load_oracle()
O.build()
# optional arguments: seed N K T output.npz (replays another case of the sweep, e.g. for debugging)
SEED, WANT, OUT = 1, (5, 80, 22477), None
if len(sys.argv) >= 6:
    SEED, WANT, OUT = int(sys.argv[1]), (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])), sys.argv[5]
rng = np.random.default_rng(SEED)
while True:
    overlap = rng.random() < 0.25
    if overlap:
        N, K = 2, int(rng.integers(4, 30))
    else:
        N, K = int(rng.integers(1, 8)), int(rng.integers(4, 98))
    T = int(rng.integers(2048, 120_000))
    sigma = float(rng.uniform(0.15, 0.8))
    rates = rng.uniform(0.0003, 0.02 if rng.random() < 0.3 else 0.004, size=N)
    pars = [(rng.uniform(1.0, 4.5), rng.uniform(0.2, 0.9), rng.uniform(0.1, 0.4)) for _ in range(N)]
    seed = int(rng.integers(1, 1 << 30))
    scale = rng.uniform(0.7, 1.1)
    zero_row = rng.random() < 0.5
    sig_m = sigma * float(rng.uniform(0.8, 1.3))
    lp = np.log(rates * rng.uniform(0.5, 2.0, size=N))
    rng.choice([0, 0, 512, 1024, 2048, 4096])
    rng.choice([0, 0, 128, 256, 512])
    if not overlap and T <= 60_000:
        rng.random()
    if (N, K, T) == WANT:
        break
temps = np.stack([hm.create_spike_template(K, *p) for p in pars], axis=1)
S = hm.create_signal(T, sigma, rates, temps, hm.make_rng(seed))
mu0 = np.asfortranarray(temps * scale)
mu0[0, :] = 0.0
lA = hm.StateMatrix(N, K, lp, False)
if OUT:  # debugging replay: inputs only
    np.savez_compressed(OUT, S=S, mu0=mu0, lp0=lp, sigma0=sig_m, N=N, K=K)
    print(OUT)
    sys.exit(0)
o = O.em_step(S, lA, mu0.copy(order="F"), sig_m)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "em_end_mass_case.npz")
np.savez_compressed(out, S=S, mu0=mu0, lp0=lp, sigma0=sig_m, N=N, K=K, lp=o[0], pp=o[1], mu=o[2], sigma=o[3], loglik=o[4])
print(out, os.path.getsize(out), "bytes; lp", o[0], "sigma", o[3])
