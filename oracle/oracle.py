"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see
hmm_oracle.c).  Imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

TRANS_DTYPE = np.dtype([("src", "<i8"), ("dst", "<i8"), ("lp", "<f8")])


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hmm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_lpz.restype = C.c_double
        _lib.orc_funcl3.restype = C.c_double
        _lib.orc_funcl4.restype = C.c_double
        _lib.orc_logsumexpl.restype = C.c_double
        _lib.orc_loglik_from_alpha.restype = C.c_double
        _lib.orc_nstates.restype = C.c_int64
        _lib.orc_get_valid_transitions.restype = C.c_int64
        for f in ("orc_funcl3",):
            getattr(_lib, f).argtypes = [C.c_double] * 3
        _lib.orc_funcl4.argtypes = [C.c_double] * 4
        _lib.orc_logsumexpl.argtypes = [C.c_double] * 2
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


i64 = C.c_int64
f64 = C.c_double


class OracleStateMatrix:
    """StateMatrix built by the oracle's own restatement of src/types.jl."""

    def __init__(self, N, K, lp, allow_overlaps=True, states0=None):
        L = lib()
        lp = np.ascontiguousarray(lp, dtype=np.float64)
        if states0 is None:
            ns = L.orc_nstates(i64(N), i64(K), C.c_int(int(allow_overlaps)))
            states0 = np.zeros((N, ns), dtype=np.int16, order="F")
            L.orc_generate_states(i64(N), i64(K), C.c_int(int(allow_overlaps)), _p(states0))
        ns = states0.shape[1]
        N = states0.shape[0]
        n = L.orc_get_valid_transitions(_p(states0), i64(N), i64(ns), i64(K), _p(lp), i64(lp.size), None, i64(0))
        tr = np.empty(n, dtype=TRANS_DTYPE)
        n2 = L.orc_get_valid_transitions(_p(states0), i64(N), i64(ns), i64(K), _p(lp), i64(lp.size), _p(tr), i64(n))
        assert n2 == n
        self.states = np.asfortranarray(states0 + np.int16(1))
        self.transitions = tr
        self.K, self.N, self.nstates = int(K), int(N), int(ns)
        self.resolve_overlaps = bool(allow_overlaps)


def _margs(lA, mu):
    mu = np.asfortranarray(mu, dtype=np.float64)
    assert mu.shape == (lA.K, lA.N), (mu.shape, lA.K, lA.N)
    st = np.asfortranarray(lA.states, dtype=np.int16)
    tr = np.ascontiguousarray(lA.transitions)
    return st, tr, mu


def viterbi(y, lA, mu, sigma, trellis=False):
    """(x, ll) as src/viterbi.jl:44-98; with trellis=True also (T2, T1)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    st, tr, mu = _margs(lA, mu)
    T = y.size
    x = np.empty(T, dtype=np.int16)
    ll = f64(0)
    T1 = T2 = None
    if trellis:
        T1 = np.empty((lA.nstates, T), dtype=np.float64, order="F")
        T2 = np.empty((lA.nstates, T), dtype=np.int16, order="F")
    rc = lib().orc_viterbi(_p(y), i64(T), _p(st), i64(lA.N), i64(lA.K), i64(lA.nstates), _p(tr), i64(tr.size), _p(mu),
                           f64(sigma), _p(x), C.byref(ll), _p(T2), _p(T1))
    if rc:
        raise RuntimeError(f"orc_viterbi rc={rc}")
    return (x, ll.value, T2, T1) if trellis else (x, ll.value)


def viterbi_screen(y, lA, mu, sigma, from_col=2, rel_a=1e-9, rel_b=64 * 2.220446049250313e-16):
    """Near-tie screen in the reference's arithmetic: dict with the smallest absolute / relative decision margin
    at columns >= from_col (0-based) and the number of decisions with relative margin below rel_a / rel_b."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    st, tr, mu = _margs(lA, mu)
    out = np.zeros(6)
    rc = lib().orc_viterbi_screen(_p(y), i64(y.size), _p(st), i64(lA.N), i64(lA.K), i64(lA.nstates), _p(tr),
                                  i64(tr.size), _p(mu), f64(sigma), i64(from_col), f64(rel_a), f64(rel_b), _p(out))
    if rc:
        raise RuntimeError(f"orc_viterbi_screen rc={rc}")
    return {"min_abs_margin": float(out[0]), "min_rel_margin": float(out[1]), "n_below_rel_a": int(out[2]),
            "n_below_rel_b": int(out[3]), "rel_a": rel_a, "rel_b": rel_b, "argmin_col": int(out[4]),
            "decisions": int(out[5]), "from_col": from_col}


def forward(V, lA, mu, sigma):
    V = np.ascontiguousarray(V, dtype=np.float64)
    st, tr, mu = _margs(lA, mu)
    a = np.empty((lA.nstates, V.size), dtype=np.float64, order="F")
    rc = lib().orc_forward(_p(V), i64(V.size), _p(st), i64(lA.N), i64(lA.K), i64(lA.nstates), _p(tr), i64(tr.size),
                           _p(mu), f64(sigma), _p(a))
    assert rc == 0
    return a


def backward(V, lA, mu, sigma):
    V = np.ascontiguousarray(V, dtype=np.float64)
    st, tr, mu = _margs(lA, mu)
    a = np.empty((lA.nstates, V.size), dtype=np.float64, order="F")
    rc = lib().orc_backward(_p(V), i64(V.size), _p(st), i64(lA.N), i64(lA.K), i64(lA.nstates), _p(tr), i64(tr.size),
                            _p(mu), f64(sigma), _p(a))
    assert rc == 0
    return a


def loglik(alpha):
    return lib().orc_loglik_from_alpha(_p(alpha), i64(alpha.shape[1]), i64(alpha.shape[0]))


def update(alpha, beta, lA, mu, sigma, x, want_gamma=False):
    """Returns (lp_new, pp, mu_new, sigma_new[, gamma]); `mu` is NOT mutated
    here (a copy is), unlike the reference -- wrappers that mirror the Julia
    in-place behaviour copy it back."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    st, tr, mu = _margs(lA, mu)
    mu = mu.copy(order="F")
    T = x.size
    nxi = int((tr["src"] == 1).sum())
    lp = np.empty(max(nxi - 1, 1), dtype=np.float64)
    pp = np.empty(lA.nstates, dtype=np.float64)
    s = f64(0)
    g = np.empty((lA.nstates, T), dtype=np.float64, order="F") if want_gamma else None
    rc = lib().orc_update(_p(alpha), _p(beta), i64(T), _p(st), i64(lA.N), i64(lA.K), i64(lA.nstates), _p(tr),
                          i64(tr.size), _p(mu), f64(sigma), _p(x), _p(lp), i64(lp.size), _p(pp), C.byref(s), _p(g))
    assert rc == 0
    out = (lp[:nxi - 1], pp, mu, s.value)
    return out + (g,) if want_gamma else out


def em_step(X, lA, mu, sigma):
    """One E/M step (src/baumwelch.jl:362-370) -> (lp_new, pp, mu_new, sigma_new, loglik)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    st, tr, mu = _margs(lA, mu)
    mu = mu.copy(order="F")
    nxi = int((tr["src"] == 1).sum())
    lp = np.empty(max(nxi - 1, 1), dtype=np.float64)
    pp = np.empty(lA.nstates, dtype=np.float64)
    s = f64(sigma)
    ll = f64(0)
    rc = lib().orc_em_step(_p(X), i64(X.size), _p(st), i64(lA.N), i64(lA.K), i64(lA.nstates), _p(tr), i64(tr.size),
                           _p(mu), C.byref(s), _p(lp), i64(lp.size), _p(pp), C.byref(ll))
    if rc:
        raise RuntimeError(f"orc_em_step rc={rc}")
    return lp[:nxi - 1], pp, mu, s.value, ll.value


def reconstruct_signal(x, lA, mu, sigma=None):
    x = np.ascontiguousarray(x, dtype=np.int16)
    st, tr, mu = _margs(lA, mu)
    Y = np.empty(x.size, dtype=np.float64)
    rc = lib().orc_reconstruct(_p(x), i64(x.size), _p(st), i64(lA.N), i64(lA.nstates), _p(mu), i64(lA.K), _p(Y))
    if rc:
        raise ValueError("state index out of range")
    return Y


def unroll_mlseq(x, lA):
    x = np.ascontiguousarray(x, dtype=np.int16)
    st = np.asfortranarray(lA.states, dtype=np.int16)
    out = np.empty((lA.N, x.size), dtype=np.int16, order="F")
    rc = lib().orc_unroll_mlseq(_p(x), i64(x.size), _p(st), i64(lA.N), i64(lA.nstates), _p(out))
    if rc:
        raise ValueError("state index out of range")
    return out


def fit_chunked(X, lA, mu, sigma, chunksize):
    X = np.ascontiguousarray(X, dtype=np.float64)
    st, tr, mu = _margs(lA, mu)
    ml = np.empty(X.size, dtype=np.int16)
    ll = f64(0)
    rc = lib().orc_fit_chunked(_p(X), i64(X.size), i64(chunksize), _p(st), i64(lA.N), i64(lA.K), i64(lA.nstates),
                               _p(tr), i64(tr.size), _p(mu), f64(sigma), _p(ml), C.byref(ll))
    assert rc == 0
    return ml, ll.value
