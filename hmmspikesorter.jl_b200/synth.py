"""Synthetic inputs: `create_spike_template` / `create_signal` restated
(src/utils.jl:49-86).  Julia's MersenneTwister stream cannot be reproduced
without Julia, so `create_signal` draws from a documented numpy
`Generator(MT19937(seed))` instead: the same process (Gaussian noise plus
non-overlapping template insertions with per-idle-sample onset probabilities
`pp`, first neuron that fires wins), a different random stream.
"""
from __future__ import annotations

import numpy as np


def create_spike_template(nstates: int, a: float = 1.0, b: float = 0.8, c: float = 0.2) -> np.ndarray:
    """src/utils.jl:51-55: a*sin(2*pi*x)*exp(-(b-x)^2/c), x = range(0, 1.5, nstates)."""
    x = np.linspace(0.0, 1.5, nstates)
    return a * np.sin(2 * np.pi * x) * np.exp(-((b - x) ** 2) / c)


def make_rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.MT19937(seed))


def create_signal(N: int, sigma: float, pp, templates: np.ndarray, rng: np.random.Generator, return_truth=False):
    """src/utils.jl:57-86.  While idle, every sample tries neurons j = 1.. in
    order and the first with `pp[j] > rand()` starts its template at that very
    sample; a template occupies `nstates` consecutive samples, then the process
    is idle again.  Here the idle gaps are drawn directly (geometric) instead
    of one uniform per idle sample per neuron -- identical in distribution."""
    templates = np.asarray(templates, dtype=np.float64)
    if templates.ndim == 1:
        templates = templates[:, None]
    K, ncells = templates.shape
    pp = np.asarray(pp, dtype=np.float64)
    S = sigma * rng.standard_normal(N)
    surv = np.concatenate(([1.0], np.cumprod(1.0 - pp)[:-1]))
    q = pp * surv  # P(first hit is neuron j) at an idle sample
    p_any = float(q.sum())
    starts = np.empty(0, dtype=np.int64)
    ids = np.empty(0, dtype=np.int64)
    if p_any > 0:
        n_max = int(N * p_any / (1.0 + K * p_any) * 1.3) + 64
        while True:
            gaps = rng.geometric(p_any, n_max).astype(np.int64) - 1
            st = np.cumsum(gaps) + K * np.arange(n_max, dtype=np.int64)
            if st[-1] >= N:
                break
            n_max *= 2
        starts = st[st < N]
        ids = rng.choice(ncells, size=starts.size, p=q / p_any)
        idx = starts[:, None] + np.arange(K, dtype=np.int64)[None, :]
        vals = templates[:, ids].T
        ok = idx < N
        S[idx[ok]] += vals[ok]
    if return_truth:
        return S, starts, ids
    return S
