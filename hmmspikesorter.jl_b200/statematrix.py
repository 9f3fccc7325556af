"""Host-side mirror of the reference's `StateMatrix` (src/types.jl:1-9).

In the drop-in deployment the *unchanged Julia constructor* builds this object
and passes its fields through `ccall`.  This module is the Python equivalent a
ctypes caller needs: same fields, same 1-based indices, same column-major
layouts, same 24-byte transition records, and the same accumulation order for
the transition log-weights (src/types.jl:94-127), so the arrays handed to
`libhmmcuda.so` are what Julia would hand it.
"""
from __future__ import annotations

import numpy as np

# == Julia Tuple{Int64,Int64,Float64} (isbits, 24 bytes): src/types.jl:3
TRANS_DTYPE = np.dtype([("src", "<i8"), ("dst", "<i8"), ("lp", "<f8")], align=False)
assert TRANS_DTYPE.itemsize == 24


def generate_states(N: int, K: int, allow_overlaps: bool = True) -> np.ndarray:
    """src/types.jl:65-92 -- 0-based ring phases, Int16 [N x nstates], F-order."""
    n = 1 + N * (K - 1)
    if allow_overlaps:
        n += (N * (N - 1) * (K - 1) * (K - 1)) // 2
    states = np.zeros((N, n), dtype=np.int16, order="F")
    k = 1
    ph = np.arange(1, K, dtype=np.int16)
    for i in range(N):
        states[i, k:k + K - 1] = ph
        k += K - 1
    if allow_overlaps:
        k1 = np.repeat(ph, K - 1)
        k2 = np.tile(ph, K - 1)
        for i in range(N - 1):
            for j in range(i + 1, N):
                m = (K - 1) * (K - 1)
                states[i, k:k + m] = k1
                states[j, k:k + m] = k2
                k += m
    return states


def lpz_of(lp: np.ndarray) -> float:
    """log1p(-exp(sum(lp))), src/types.jl:96 (left-to-right sum)."""
    s = 0.0
    for i, v in enumerate(np.asarray(lp, dtype=np.float64)):
        s = float(v) if i == 0 else s + float(v)
    return float(np.log1p(-np.exp(s)))


def get_valid_transitions(states0: np.ndarray, K: int, lp: np.ndarray, row_block: int = 256) -> np.ndarray:
    """src/types.jl:94-127.  Vectorised over (src, dst) pairs; the per-neuron
    terms are added in neuron order exactly as `lpt += ...` does, and an
    impossible neuron transition makes the whole sum -Inf (the `break`)."""
    N, n = states0.shape
    lp = np.asarray(lp, dtype=np.float64)
    lpz = lpz_of(lp)  # the WHOLE vector (src/types.jl:96): overlap models pass N + N(N-1)/2 entries, only lp[1:N] are indexed
    out = []
    s = states0.astype(np.int32)
    for r0 in range(0, n, row_block):
        r1 = min(n, r0 + row_block)
        lpt = np.zeros((r1 - r0, n), dtype=np.float64)
        for i in range(N):
            s1 = s[i, r0:r1][:, None]
            s2 = s[i, :][None, :]
            c1 = (s1 == 0) & (s2 == 0)
            c2 = (s1 == 0) & (s2 == 1)
            c3 = ((s2 - s1) == 1) | ((s1 == K - 1) & (s2 == 0))
            term = np.where(c1, lpz, np.where(c2, lp[i], np.where(c3, 0.0, -np.inf)))
            lpt = lpt + term
        src, dst = np.nonzero(np.isfinite(lpt))
        rec = np.empty(src.size, dtype=TRANS_DTYPE)
        rec["src"] = src + r0 + 1
        rec["dst"] = dst + 1
        rec["lp"] = lpt[src, dst]
        out.append(rec)
    return np.concatenate(out) if out else np.empty(0, dtype=TRANS_DTYPE)


_TOPOLOGY_CACHE: dict = {}


def _topology(states0: np.ndarray, K: int, allow_overlaps: bool):
    """(src, dst, codes) of the finite transitions of a state layout.  Which transitions are
    finite does not depend on lp (src/types.jl:94-113), only their weights do, so the
    O(nstates^2 N) scan is done once per layout; codes[e, i] says which term neuron i adds to
    edge e: 0 -> lpz, 1 -> lp[i], 2 -> 0.0."""
    N, n = states0.shape
    key = (N, int(K), bool(allow_overlaps), n, hash(states0.tobytes()))
    hit = _TOPOLOGY_CACHE.get(key)
    if hit is not None:
        return hit
    out = _enumerate_topology(states0, K)
    if len(_TOPOLOGY_CACHE) > 32:
        _TOPOLOGY_CACHE.clear()
    _TOPOLOGY_CACHE[key] = out
    return out


def _enumerate_topology(states0: np.ndarray, K: int):
    """The finite transitions of a state layout by direct enumeration instead of the reference's O(nstates^2 N) scan of
    all pairs (src/types.jl:114-127; 45 s in numpy for the CLI's 21 123-state model, 13 M checks for the 3 600-state
    one): from a source state every neuron has one or two admissible next phases -- silent: stay silent or start;
    active: advance, or return to silence from the last phase (src/types.jl:94-113) -- and a combination is a
    transition iff it is a state of the layout.  Same records in the same (src, dst) order; `tests/test_oracle.py`
    compares it with the literal scan."""
    N, n = states0.shape
    cols = [tuple(int(v) for v in states0[:, j]) for j in range(n)]
    index = {c: j for j, c in enumerate(cols)}
    src, dst, codes = [], [], []
    for j, c in enumerate(cols):
        combos = [((), ())]
        for s1 in c:
            if s1 == 0:
                opts = ((0, 0), (1, 1))
            elif s1 == K - 1:
                opts = ((0, 2),)
            else:
                opts = ((s1 + 1, 2),)
            combos = [(st + (s2,), cd + (code,)) for st, cd in combos for s2, code in opts]
        found = []
        for st, cd in combos:
            k = index.get(st)
            if k is not None:
                found.append((k, cd))
        found.sort(key=lambda q: q[0])
        for k, cd in found:
            src.append(j)
            dst.append(k)
            codes.append(cd)
    return (np.asarray(src, dtype=np.int64) + 1, np.asarray(dst, dtype=np.int64) + 1,
            np.asarray(codes, dtype=np.int8).reshape(len(src), N))


def _scan_topology(states0: np.ndarray, K: int):
    """The same by the literal all-pairs scan (kept for the cross-check)."""
    N = states0.shape[0]
    tr = get_valid_transitions(states0, K, np.full(N, np.log(0.5 / max(N, 1))))
    src = tr["src"].astype(np.int64) - 1
    dst = tr["dst"].astype(np.int64) - 1
    s1 = states0[:, src].astype(np.int32).T  # [ntrans, N]
    s2 = states0[:, dst].astype(np.int32).T
    codes = np.where((s1 == 0) & (s2 == 0), 0, np.where((s1 == 0) & (s2 == 1), 1, 2)).astype(np.int8)
    return tr["src"].copy(), tr["dst"].copy(), codes


def transitions_fast(states0: np.ndarray, K: int, lp, allow_overlaps: bool) -> np.ndarray:
    """Same records as get_valid_transitions, bit for bit (the per-neuron terms are added in
    the same order), using the cached topology."""
    src, dst, codes = _topology(states0, K, allow_overlaps)
    lp = np.asarray(lp, dtype=np.float64)
    N = states0.shape[0]
    lpz = lpz_of(lp)  # the WHOLE vector (src/types.jl:96): overlap models pass N + N(N-1)/2 entries, only lp[1:N] are indexed
    lpt = np.zeros(src.size, dtype=np.float64)
    for i in range(N):
        c = codes[:, i]
        lpt = lpt + np.where(c == 0, lpz, np.where(c == 1, lp[i], 0.0))
    if not np.all(np.isfinite(lpt)):  # degenerate lp (e.g. -Inf): the set of finite edges changes
        return get_valid_transitions(states0, K, lp)
    rec = np.empty(src.size, dtype=TRANS_DTYPE)
    rec["src"], rec["dst"], rec["lp"] = src, dst, lpt
    return rec


class StateMatrix:
    """Fields as src/types.jl:1-9.  NOTE the reference's field comments are
    swapped; as in the constructor call (src/types.jl:150) `K` is states per
    ring (incl. the silent one) and `N` the number of neurons."""

    __slots__ = ("states", "transitions", "pi", "K", "N", "nstates", "resolve_overlaps")

    def __init__(self, N: int, K: int, lp, allow_overlaps: bool = True, pp=None, _states0=None):
        states0 = generate_states(N, K, allow_overlaps) if _states0 is None else _states0
        nstates = states0.shape[1]
        if nstates > 32767:
            raise ValueError("nstates exceeds Int16 range")
        lp = np.ascontiguousarray(lp, dtype=np.float64)
        if pp is None:  # src/types.jl:138 log.(ones(nstates)./nstates)
            pp = np.log(np.ones(nstates) / nstates)
        self.transitions = transitions_fast(states0, K, lp, allow_overlaps)
        self.states = np.asfortranarray(states0 + np.int16(1))  # src/types.jl:150
        self.pi = np.ascontiguousarray(pp, dtype=np.float64)
        self.K = int(K)
        self.N = int(states0.shape[0])
        self.nstates = int(nstates)
        self.resolve_overlaps = bool(allow_overlaps)

    @classmethod
    def from_states(cls, states1: np.ndarray, pp, K: int, lp, allow_overlaps: bool = True) -> "StateMatrix":
        """StateMatrix(states, pp, K, lp; allow_overlaps) of src/types.jl:148-151,
        called with `lA.states .- 1` by `update` (src/baumwelch.jl:265)."""
        s0 = np.asfortranarray(np.asarray(states1, dtype=np.int16) - np.int16(1))
        return cls(s0.shape[0], K, lp, allow_overlaps, pp=pp, _states0=s0)

    def isempty(self) -> bool:
        return self.states.size == 0

    def get_lp(self):
        """src/types.jl:42-61: noise->head log-probs and their neuron index."""
        lp = np.zeros(self.N)
        lidx = np.zeros(self.N, dtype=np.int64)
        k = 0
        for rec in self.transitions:
            if rec["src"] == 1 and rec["dst"] > 1:
                v = np.nonzero(self.states[:, rec["dst"] - 1] > 1)[0]
                if v.size == 1:
                    lp[k] = rec["lp"]
                    lidx[k] = v[0] + 1
                    if k == self.N - 1:
                        break
                    k += 1
        return lp, lidx
