"""hmmcuda -- host-side mirror of the HMMSpikeSorter.jl hot-path API over
libhmmcuda.so (B200 / sm_100a CUDA kernels behind a C ABI).

Same function names, argument meaning and error behaviour as the Julia
methods they stand in for (no Julia is available in the build environment;
this ctypes layer binds the identical symbols with the identical memory
layouts the Julia `ccall` shim in `julia/HMMCuda.jl` uses):

    viterbi(y, lA, mu, sigma)            src/viterbi.jl:44-98      -> (x, ll)
    viterbi(..., trellis=True)           README.md:34 form         -> (x, T2, T1)
    forward / backward(V, lA, mu, sigma) src/baumwelch.jl:25-51,73-98
    update(alpha, beta, lA, mu, sigma, x) src/baumwelch.jl:205-309 -> (lA_new, mu, sigma)
    train_model(X, lA, mu0, sigma0)      src/baumwelch.jl:362-370  -> (lA_new, mu, sigma)
    train_model(X, lA, mu, sigma, nsteps, callback)  the E/M loop of src/baumwelch.jl:324-335
    reconstruct_signal(x, lA, mu, sigma) src/reconstruction.jl:1-9
    unroll_mlseq(mlseq, lA)              src/extraction.jl:4-13
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MODES, HmmArgumentError, HmmError, HmmInfo, check, lib
from .statematrix import TRANS_DTYPE, StateMatrix, generate_states, get_valid_transitions
from .synth import create_signal, create_spike_template, make_rng
from . import sharding
from . import timeshard
from .timeshard import viterbi_time_sharded

__all__ = [
    "StateMatrix", "viterbi", "viterbi_batch", "forward", "backward", "update", "train_model", "em_step",
    "reconstruct_signal", "unroll_mlseq", "TrainContext", "create_signal", "create_spike_template", "make_rng",
    "HmmError", "HmmArgumentError", "device_count", "set_ring_params", "set_devices", "set_precision", "viterbi_f32", "viterbi_rawfile", "save_sort_result", "viterbi_time_sharded", "transition_weights",
]

i64, i32, f64 = C.c_int64, C.c_int32, C.c_double


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count() -> int:
    return int(lib().hmm_device_count())


def set_devices(devices=None) -> None:
    """Devices the host-pointer decodes spread their work over inside the library (hmm_set_devices): the channels
    of viterbi_batch in blocks, one long recording of viterbi as time shards.  None / one device: single-device."""
    d = [] if devices is None else [int(v) for v in devices]
    arr = (C.c_int * max(1, len(d)))(*d)
    check(lib().hmm_set_devices(arr, C.c_int(len(d))))


def set_precision(precision: str = "f64") -> None:
    """"f64" (default) or "f32": FP32 mode of the ring decode (hmm_set_precision): the FIR in FP32, everything
    else FP64; ll within 1e-4 relative, x not promised bit-exact."""
    check(lib().hmm_set_precision(i32({"f64": 0, "f32": 1}[precision])))


def viterbi_f32(y, lA, mu, sigma, *, mode: str = "auto", return_info: bool = False):
    """FP32 mode with a Float32 recording (hmm_viterbi_ex_f32): returns (x, ll) like viterbi."""
    y = np.ascontiguousarray(np.asarray(y), dtype=np.float32)
    if y.ndim != 1:
        raise HmmArgumentError(_lib.HMM_EINVAL, "y must be a vector")
    st, tr, mu = _model_args(lA, mu)
    x = np.empty(y.size, dtype=np.int16)
    ll = f64(0)
    info = HmmInfo()
    check(lib().hmm_viterbi_ex_f32(_p(y), i64(y.size), _p(st), i32(lA.N), i32(lA.K), i32(lA.nstates), _p(tr),
                                   i64(tr.size), _p(mu), f64(sigma), _p(x), C.byref(ll), i32(MODES[mode]), C.byref(info)))
    return (x, ll.value, info.asdict()) if return_info else (x, ll.value)


RAW_DTYPES = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int16): 2}


def viterbi_rawfile(path, dtype, n_file_channels: int, T: int, channels, models, *, interleaved: bool = True,
                    scale: float = 1.0, byte_offset: int = 0, mode: str = "auto", return_info: bool = False):
    """I/O front-end (hmm_viterbi_rawfile; src/hmmsort.jl:66-90 reads the data file, converts it to Float64 and
    decodes): decode `channels` (0-based indices) of a raw recording file -- int16 / float32 / float64 samples,
    interleaved [T x n_file_channels] or channel-major, starting at `byte_offset` -- straight from the file through
    pinned staging; models = [(lA, mu, sigma)] per requested channel.  Returns (x [T x C] Int16, ll [C])."""
    ch = np.ascontiguousarray(channels, dtype=np.int32)
    if len(models) != ch.size:
        raise HmmArgumentError(_lib.HMM_EINVAL, "one model per requested channel")
    lA0 = models[0][0]
    sts, trs, mus, sig = [], [], [], []
    for lA, mu, s in models:
        st, tr, mu = _model_args(lA, mu)
        sts.append(st.ravel(order="F")); trs.append(tr); mus.append(mu.ravel(order="F")); sig.append(float(s))
    st = np.ascontiguousarray(np.concatenate(sts)); tr = np.ascontiguousarray(np.concatenate(trs))
    mu = np.ascontiguousarray(np.concatenate(mus)); sig = np.asarray(sig, dtype=np.float64)
    x = np.empty((T, ch.size), dtype=np.int16, order="F")
    ll = np.empty(ch.size, dtype=np.float64)
    info = HmmInfo()
    check(lib().hmm_viterbi_rawfile(str(path).encode(), i64(byte_offset), i32(RAW_DTYPES[np.dtype(dtype)]),
                                    i32(n_file_channels), i32(1 if interleaved else 0), f64(scale), i64(T), i32(ch.size),
                                    _p(ch), _p(st), i32(0), i32(lA0.N), i32(lA0.K), i32(lA0.nstates), _p(tr),
                                    i64(lA0.transitions.size), _p(mu), _p(sig), _p(x), _p(ll), i32(MODES[mode]),
                                    C.byref(info)))
    return (x, ll, info.asdict()) if return_info else (x, ll)


def save_sort_result(outputfile, x, lA, mu, sigma, ll):
    """The .mat file sort_data writes for one channel (src/hmmsort.jl:92-101): the unrolled most likely sequence
    (Int16 [N x T], src/extraction.jl:4-13), ll, the templates, the noise -> active log-probabilities and sigma."""
    from scipy.io import savemat

    lp, _ = lA.get_lp()
    out = {"mlseq": unroll_mlseq(x, lA), "ll": float(ll), "waveforms": np.asarray(mu), "lp": lp, "sigma": float(sigma)}
    savemat(str(outputfile), out)
    return out


def set_ring_params(chunk_len: int = 0, warmup: int = 0) -> None:
    """Tunables of the time-parallel ring engine (0 = default): chunk length and
    speculative warm-up / look-ahead in samples."""
    check(lib().hmm_set_ring_params(i64(chunk_len), i64(warmup)))


def _model_args(lA, mu):
    mu = np.asarray(mu)
    if mu.dtype != np.float64 or mu.ndim != 2:
        raise HmmArgumentError(_lib.HMM_EINVAL, "mu must be a Float64 matrix [K x N]")
    if mu.shape != (lA.K, lA.N):
        raise HmmArgumentError(_lib.HMM_EINVAL, f"mu has shape {mu.shape}, expected (K, N) = ({lA.K}, {lA.N})")
    st = np.asfortranarray(lA.states, dtype=np.int16)
    tr = np.ascontiguousarray(lA.transitions, dtype=TRANS_DTYPE)
    return st, tr, np.asfortranarray(mu)


def _vec(y, name="y"):
    y = np.asarray(y)
    if y.ndim != 1:
        raise HmmArgumentError(_lib.HMM_EINVAL, f"{name} must be a vector")
    return np.ascontiguousarray(y, dtype=np.float64)  # non-unit strides are collected, like the Julia shim


def viterbi(y, lA, mu, sigma, *, trellis: bool = False, mode: str = "auto", return_info: bool = False):
    """Most likely state sequence.  Returns (x::Int16[T] 1-based, ll) as the live
    reference method (src/viterbi.jl:97); with trellis=True returns (x, T2, T1)
    with T2 Int16 / T1 Float64 [nstates x T] (src/viterbi.jl:52-53, README.md:34)."""
    y = _vec(y)
    st, tr, mu = _model_args(lA, mu)
    T = y.size
    x = np.empty(T, dtype=np.int16)
    ll = f64(0)
    T1 = np.empty((lA.nstates, T), dtype=np.float64, order="F") if trellis else None
    T2 = np.empty((lA.nstates, T), dtype=np.int16, order="F") if trellis else None
    info = HmmInfo()
    check(lib().hmm_viterbi_ex_f64(_p(y), i64(T), _p(st), i32(lA.N), i32(lA.K), i32(lA.nstates), _p(tr), i64(tr.size),
                                   _p(mu), f64(sigma), _p(x), C.byref(ll), _p(T2), _p(T1), i32(MODES[mode]),
                                   C.byref(info)))
    out = (x, T2, T1) if trellis else (x, ll.value)
    return out + (info.asdict(),) if return_info else out


def viterbi_batch(Y, models, *, mode: str = "auto", return_info: bool = False):
    """Decode C independent channels: Y [T x C] (column-major), models = list of
    (lA, mu, sigma) with a common topology.  Returns (x [T x C] Int16, ll [C])."""
    Y = np.asfortranarray(Y, dtype=np.float64)
    if Y.ndim != 2 or Y.shape[1] != len(models):
        raise HmmArgumentError(_lib.HMM_EINVAL, "Y must be [T x C] with one model per column")
    T, Cn = Y.shape
    lA0 = models[0][0]
    sts, trs, mus, sig = [], [], [], []
    for lA, mu, s in models:
        if (lA.N, lA.K, lA.nstates, lA.transitions.size) != (lA0.N, lA0.K, lA0.nstates, lA0.transitions.size):
            raise HmmArgumentError(_lib.HMM_EINVAL, "all channels must share N, K, nstates, ntrans")
        st, tr, mu = _model_args(lA, mu)
        sts.append(st.ravel(order="F")); trs.append(tr); mus.append(mu.ravel(order="F")); sig.append(float(s))
    st = np.ascontiguousarray(np.concatenate(sts))
    tr = np.ascontiguousarray(np.concatenate(trs))
    mu = np.ascontiguousarray(np.concatenate(mus))
    sig = np.asarray(sig, dtype=np.float64)
    x = np.empty((T, Cn), dtype=np.int16, order="F")
    ll = np.empty(Cn, dtype=np.float64)
    info = HmmInfo()
    check(lib().hmm_viterbi_batch_f64(_p(Y), i64(T), i32(Cn), _p(st), i32(0), i32(lA0.N), i32(lA0.K),
                                      i32(lA0.nstates), _p(tr), i64(lA0.transitions.size), _p(mu), _p(sig), _p(x),
                                      _p(ll), i32(MODES[mode]), C.byref(info)))
    return (x, ll, info.asdict()) if return_info else (x, ll)


def _fb(fn, V, lA, mu, sigma):
    V = _vec(V, "V")
    st, tr, mu = _model_args(lA, mu)
    out = np.empty((lA.nstates, V.size), dtype=np.float64, order="F")
    check(fn(_p(V), i64(V.size), _p(st), i32(lA.N), i32(lA.K), i32(lA.nstates), _p(tr), i64(tr.size), _p(mu),
             f64(sigma), _p(out)))
    return out


def forward(V, lA, mu, sigma):
    """alpha [nstates x T], src/baumwelch.jl:25-51."""
    return _fb(lib().hmm_forward_f64, V, lA, mu, sigma)


def backward(V, lA, mu, sigma):
    """beta [nstates x T], src/baumwelch.jl:73-98."""
    return _fb(lib().hmm_backward_f64, V, lA, mu, sigma)


def _rebuild(lA, lp, pp):
    # src/baumwelch.jl:265: StateMatrix(lA.states .- 1, pp, K, xb[2:end]; allow_overlaps=lA.resolve_overlaps)
    return StateMatrix.from_states(lA.states, pp, lA.K, lp, lA.resolve_overlaps)


def transition_weights(lA, lp):
    """Transition records of `lA` with the weights a new `lp` gives them (hmm_transition_weights: the weight part of the
    StateMatrix rebuild, src/types.jl:94-127; host arithmetic, no device needed).  Returns None when a weight would
    not be finite (the set of transitions changes: run the constructor)."""
    st = np.asfortranarray(lA.states, dtype=np.int16)
    tr = np.ascontiguousarray(lA.transitions).copy()
    lp = np.ascontiguousarray(lp, dtype=np.float64)
    ok = i32(0)
    check(lib().hmm_transition_weights(_p(st), i32(lA.N), i32(lA.nstates), _p(tr), i64(tr.size), _p(lp), i32(lp.size),
                                       C.byref(ok)))
    return tr if ok.value else None


def _nxi(lA):
    return int((lA.transitions["src"] == 1).sum())


def update(alpha, beta, lA, mu, sigma, x):
    """(lA_new, mu, sigma), src/baumwelch.jl:205-309.  `mu` is overwritten in
    place like the reference's fill!(mu, 0.0) and also returned."""
    x = _vec(x, "x")
    if not (isinstance(mu, np.ndarray) and mu.dtype == np.float64 and mu.flags.f_contiguous and mu.flags.writeable):
        raise HmmArgumentError(_lib.HMM_EINVAL, "mu must be a writeable column-major Float64 matrix (it is updated in place)")
    st, tr, _ = _model_args(lA, mu)
    alpha = np.asfortranarray(alpha, dtype=np.float64)
    beta = np.asfortranarray(beta, dtype=np.float64)
    if alpha.shape != (lA.nstates, x.size) or beta.shape != alpha.shape:
        raise HmmArgumentError(_lib.HMM_EINVAL, "alpha/beta must be [nstates x T]")
    lp = np.empty(max(_nxi(lA) - 1, 1), dtype=np.float64)
    pp = np.empty(lA.nstates, dtype=np.float64)
    s = f64(sigma)
    check(lib().hmm_update_f64(_p(alpha), _p(beta), i64(x.size), _p(st), i32(lA.N), i32(lA.K), i32(lA.nstates),
                               _p(tr), i64(tr.size), _p(mu), C.byref(s), _p(x), _p(lp), _p(pp)))
    return _rebuild(lA, lp[:_nxi(lA) - 1], pp), mu, s.value


def em_step(X, lA, mu, sigma, *, mode: str = "auto", return_info: bool = False):
    """One fused E/M step; returns (lp_new, pp, mu_new, sigma_new, loglik).  `mu`
    is not modified (the raw form used by tests and benches)."""
    X = _vec(X, "X")
    st, tr, mu = _model_args(lA, mu)
    mu = mu.copy(order="F")
    lp = np.empty(max(_nxi(lA) - 1, 1), dtype=np.float64)
    pp = np.empty(lA.nstates, dtype=np.float64)
    s, ll = f64(sigma), f64(0)
    info = HmmInfo()
    check(lib().hmm_em_step_ex_f64(_p(X), i64(X.size), _p(st), i32(lA.N), i32(lA.K), i32(lA.nstates), _p(tr),
                                   i64(tr.size), _p(mu), C.byref(s), _p(lp), _p(pp), C.byref(ll), i32(MODES[mode]),
                                   C.byref(info)))
    out = (lp[:_nxi(lA) - 1].copy(), pp, mu, s.value, ll.value)
    return out + (info.asdict(),) if return_info else out


class TrainContext:
    """Keeps X resident in HBM across E/M iterations (hmm_train_* of the C ABI)."""

    def __init__(self, X):
        X = _vec(X, "X")
        self._h = C.c_void_p()
        self.T = X.size
        check(lib().hmm_train_create(_p(X), i64(X.size), C.byref(self._h)))

    def em_step(self, lA, mu, sigma, return_info=False):
        st, tr, mu = _model_args(lA, mu)
        mu = mu.copy(order="F")
        lp = np.empty(max(_nxi(lA) - 1, 1), dtype=np.float64)
        pp = np.empty(lA.nstates, dtype=np.float64)
        s, ll = f64(sigma), f64(0)
        info = HmmInfo()
        check(lib().hmm_train_em_step(self._h, _p(st), i32(lA.N), i32(lA.K), i32(lA.nstates), _p(tr), i64(tr.size),
                                      _p(mu), C.byref(s), _p(lp), _p(pp), C.byref(ll), C.byref(info)))
        out = (lp[:_nxi(lA) - 1].copy(), pp, mu, s.value, ll.value)
        return out + (info.asdict(),) if return_info else out

    def run(self, lA, mu, sigma, nsteps, return_info=False):
        """`nsteps` E/M iterations in one library call (hmm_train_run): the loop of src/baumwelch.jl:325-335 without a
        callback.  Returns (lA_new, mu, sigma, loglik[steps_done]); steps_done < nsteps only if lp degenerated."""
        st, tr, mu = _model_args(lA, mu)
        mu = mu.copy(order="F")
        tr = tr.copy()
        nlp = max(_nxi(lA) - 1, 1)
        lp = np.empty(nlp, dtype=np.float64)
        pp = np.empty(lA.nstates, dtype=np.float64)
        ll = np.zeros(max(int(nsteps), 1), dtype=np.float64)
        s, done = f64(sigma), i32(0)
        info = HmmInfo()
        check(lib().hmm_train_run(self._h, _p(st), i32(lA.N), i32(lA.K), i32(lA.nstates), _p(tr), i64(tr.size), _p(mu),
                                  C.byref(s), _p(lp), i32(nlp), _p(pp), _p(ll), i32(int(nsteps)), C.byref(done),
                                  C.byref(info)))
        lA_new = _rebuild(lA, lp[:_nxi(lA) - 1], pp) if done.value > 0 else lA
        out = (lA_new, mu, s.value, ll[:done.value].copy())
        return out + (info.asdict(),) if return_info else out

    def close(self):
        if self._h:
            lib().hmm_train_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def train_model(X, lA, mu, sigma, nsteps=None, callback=None, *, verbose: int = 0):
    """train_model(X, lA, mu0, sigma0) -> (lA_new, mu, sigma): one E/M step
    (src/baumwelch.jl:362-370).  With `nsteps`, the E/M loop of
    src/baumwelch.jl:325-335 -- `callback(mu)` before every step, stop on an empty
    model -- with X kept device-resident; `mu` is updated in place each step
    exactly as the reference's `update` does.  The merge/prune phase
    of src/baumwelch.jl:340-352 is host model management and stays with the
    caller."""
    X = _vec(X, "X")
    if not (isinstance(mu, np.ndarray) and mu.dtype == np.float64 and mu.flags.f_contiguous and mu.flags.writeable):
        raise HmmArgumentError(_lib.HMM_EINVAL, "mu must be a writeable column-major Float64 matrix (it is updated in place)")
    if nsteps is None:
        lp, pp, mu_new, s, _ = em_step(X, lA, mu, sigma)
        mu[...] = mu_new
        return _rebuild(lA, lp, pp), mu, s
    if callback is None and verbose <= 0:  # the whole loop inside the library
        with TrainContext(X) as ctx:
            left = int(nsteps)
            while left > 0 and not lA.isempty():
                lA, mu_new, sigma, ll = ctx.run(lA, mu, sigma, left)  # stops early only if lp degenerated:
                mu[...] = mu_new                                      # lA is then rebuilt here (new topology)
                left -= max(ll.size, 1)
        return lA, mu, sigma
    with TrainContext(X) as ctx:
        for i in range(nsteps):
            if verbose > 0:
                print(f"{i + 1} ", end="", flush=True)
            if callback is not None:
                callback(mu)
            lp, pp, mu_new, sigma, _ = ctx.em_step(lA, mu, sigma)
            mu[...] = mu_new
            lA = _rebuild(lA, lp, pp)
            if lA.isempty():
                break
    if verbose > 0:
        print()
    return lA, mu, sigma


def reconstruct_signal(x, lA, mu, sigma=None):
    """Y[i] = sum_j mu[states[j, x[i]], j], src/reconstruction.jl:1-9."""
    x = np.asarray(x)
    if not np.issubdtype(x.dtype, np.integer) or x.ndim != 1:
        raise HmmArgumentError(_lib.HMM_EINVAL, "x must be an integer vector")
    if x.size and (x.min() < 1 or x.max() > lA.nstates):
        raise HmmArgumentError(_lib.HMM_EINVAL, "state index outside 1..nstates")  # Julia: BoundsError
    x = np.ascontiguousarray(x, dtype=np.int16)
    st, _, mu = _model_args(lA, mu)
    Y = np.empty(x.size, dtype=np.float64)
    check(lib().hmm_reconstruct_f64(_p(x), i64(x.size), _p(st), i32(lA.N), i32(lA.nstates), _p(mu), i32(lA.K), _p(Y)))
    return Y


def unroll_mlseq(mlseq, lA):
    """Int16 [N x T] per-neuron ring phase, src/extraction.jl:4-13."""
    x = np.ascontiguousarray(mlseq, dtype=np.int16)
    st = np.asfortranarray(lA.states, dtype=np.int16)
    out = np.empty((lA.N, x.size), dtype=np.int16, order="F")
    check(lib().hmm_unroll_mlseq_i16(_p(x), i64(x.size), _p(st), i32(lA.N), i32(lA.nstates), _p(out)))
    return out
