"""TEST INFRASTRUCTURE: a literal Python restatement of the reference's host-side training driver -- the outer loop
of train_model (src/baumwelch.jl:324-354) with its merge / prune phase (condense_templates :446-540,
find_best_overlap :545-571, remove_sparse :573-592, remove_small :423-432, get_lp / prune_templates
src/types.jl:42-61,161-166) -- so that BASELINE config 1 can be replayed end to end with the E/M steps supplied
either by the CPU oracle or by libhmmcuda.  In the drop-in this code stays Julia ("OUT OF SCOPE");
it is restated here WITH the reference's quirks, which decide what model the second E/M phase runs on:
  * `μ_new[xi2, kk] .=+ 0.5*μ[xi2,i2]` assigns, it does not accumulate (:462);
  * `N -= 1` precedes `setdiff(1:N, [i1,i2])` (:457,464), so a merge leaves the last column of μ_new zero;
  * prune_templates indexes `lp[findall(in(tidx), idx)]`, i.e. POSITIONS in idx, not template numbers.
Indices below are 0-based where Python needs them; the restated arithmetic is order-for-order the reference's."""
from __future__ import annotations

import numpy as np
from scipy.stats import chi2


def get_lp(sm):
    """src/types.jl:42-61."""
    lp = np.zeros(sm.N)
    lidx = np.zeros(sm.N, dtype=np.int64)
    k = 0
    for rec in sm.transitions:
        if rec["src"] == 1 and rec["dst"] > 1:
            v = np.nonzero(np.asarray(sm.states)[:, rec["dst"] - 1] > 1)[0]
            if v.size == 1:
                lp[k] = rec["lp"]
                lidx[k] = v[0] + 1
                if k == sm.N - 1:
                    break
                k += 1
    return lp, lidx


def prune_templates(ctor, sm, idx, resolve_overlaps):
    """src/types.jl:161-166; idx: 1-based template numbers."""
    lp, tidx = get_lp(sm)
    pos = [k for k, v in enumerate(idx) if v in set(tidx.tolist())]  # findall(in(tidx), idx): positions in idx
    return ctor(len(idx), sm.K, lp[pos], resolve_overlaps)


def find_best_overlap(mu, i1, i2):
    """src/baumwelch.jl:545-571 (i1, i2 0-based columns); returns ((range1, range2), xm) with 0-based ranges."""
    K = mu.shape[0]
    xi = (range(0, K), range(0, K))
    xm = -np.inf
    shifts = [(range(0, s), range(K - s, K)) for s in range(1, K + 1)]
    shifts += [(range(s, K), range(0, K - s)) for s in range(1, K)]
    for sh in shifts:
        x = 0.0
        for k1, k2 in zip(*sh):
            x += mu[k1, i1] * mu[k2, i2]
        if x > xm:
            xm = x
            xi = sh
    return xi, xm


def condense_candidates(mu, sigma2, alpha=0.05):
    """src/baumwelch.jl:482-540: the pair to merge next (most similar first) or None."""
    K, N = mu.shape
    cands, stats, ovl = [], [], []
    for i1 in range(N - 1):
        for i2 in range(i1 + 1, N):
            xi, _ = find_best_overlap(mu, i1, i2)
            x = 0.0
            for k1, k2 in zip(*xi):
                x += abs(mu[k1, i1] - mu[k2, i2]) ** 2
            x /= sigma2
            n = len(xi[0])
            pval = 0.0 if n < 5 else 1 - chi2.cdf(x, n - 1)
            if pval > alpha:
                cands.append((i1, i2))
                stats.append(x)
                ovl.append(xi)
    if cands:
        m = int(np.argmax(stats))
        return cands[m], stats[m], ovl[m]
    return None


def condense_templates(ctor, sm, mu, sigma, alpha=0.05):
    """src/baumwelch.jl:446-480, quirks included."""
    sigma2 = sigma ** 2
    lp, _ = get_lp(sm)
    K, N = mu.shape
    c = condense_candidates(mu, sigma2, alpha)
    while c is not None:
        (i1, i2), _, (xi1, xi2) = c
        N -= 1
        mu_new = np.zeros((K, N))
        lp_new = np.zeros(N)
        mu_new[list(xi1), 0] = 0.5 * mu[list(xi1), i1]
        mu_new[list(xi2), 0] = +0.5 * mu[list(xi2), i2]          # `.=+` : assignment
        lp_new[0] = np.log(0.5 * np.exp(lp[i1]) + 0.5 * np.exp(lp[i2]))
        idx = [j for j in range(N) if j not in (i1, i2)]         # setdiff(1:N, [i1,i2]) with N already decremented
        for ii, jj in enumerate(idx, start=1):
            mu_new[:, ii] = mu[:, jj]
            lp_new[ii] = lp[jj]
        lp, mu = lp_new, mu_new
        c = condense_candidates(mu, sigma2, alpha)
    if N < sm.N:
        return ctor(N, K, lp, sm.resolve_overlaps), np.asfortranarray(mu)
    return sm, mu


def remove_sparse(ctor, sm, lp0=-70.0):
    """src/baumwelch.jl:573-592; returns (state matrix or None when empty, 1-based template numbers)."""
    tt = [r for r in sm.transitions if r["src"] == 1 and r["dst"] != 1 and r["lp"] > lp0]
    if not tt:
        return None, []
    tidx = []
    st = np.asarray(sm.states)
    for r in tt:
        for j in range(st.shape[0]):
            if st[j, r["dst"] - 1] == 2:
                tidx.append(j + 1)
                break
    return prune_templates(ctor, sm, tidx, sm.resolve_overlaps), tidx


def remove_small(ctor, sm, mu, sigma, alpha=0.05):
    """src/baumwelch.jl:423-432: chi-square test of the template energy against noise."""
    K = mu.shape[0]
    Z = (mu ** 2).sum(axis=0) / (sigma * sigma)
    pvals = 1 - chi2.cdf(Z, K - 1)
    tidx = [int(i) + 1 for i in np.nonzero(pvals < alpha)[0]]
    return prune_templates(ctor, sm, tidx, sm.resolve_overlaps), tidx


def train_model(em_step, ctor, X, sm, mu, sigma, nsteps, callback=None):
    """src/baumwelch.jl:324-354.  em_step(X, sm, mu, sigma) -> (lp_new, pp, mu_new, sigma_new, ...); ctor(N, K, lp,
    resolve_overlaps) builds a StateMatrix.  Returns (sm, mu, sigma, log) with the intermediate models in `log`."""
    log = {}
    mu = np.asfortranarray(mu.copy())
    for _ in range(nsteps):
        if callback is not None:
            callback(mu)
        r = em_step(X, sm, mu, sigma)
        mu, sigma = np.asfortranarray(r[2]), r[3]
        sm = ctor(sm.N, sm.K, r[0], sm.resolve_overlaps)
        if np.asarray(sm.states).size == 0:
            break
    log["after_phase1"] = (sm, mu.copy(), sigma)
    sm, mu = condense_templates(ctor, sm, mu, sigma, 0.05)
    log["after_condense"] = (sm.N, mu.copy())
    sm, idx = remove_sparse(ctor, sm)
    log["after_sparse"] = list(idx)
    if sm is None:
        return None, mu, sigma, log
    sm, idx2 = remove_small(ctor, sm, mu[:, [i - 1 for i in idx]], sigma, 0.05)
    log["after_small"] = list(idx2)
    mu = np.asfortranarray(mu[:, [idx[i - 1] - 1 for i in idx2]])
    for _ in range(nsteps // 2):
        r = em_step(X, sm, mu, sigma)
        mu, sigma = np.asfortranarray(r[2]), r[3]
        sm = ctor(sm.N, sm.K, r[0], sm.resolve_overlaps)
    return sm, mu, sigma, log
