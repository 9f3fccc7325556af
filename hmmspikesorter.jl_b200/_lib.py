"""Loader for libhmmcuda.so.  There is no Python or CPU fallback: if the
shared library has not been built, importing the compute entry points fails
loudly (build it with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C hmmspikesorter.jl_b200/csrc`)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LIBHMMCUDA", os.path.join(_HERE, "libhmmcuda.so"))  # same override as the Julia shim

HMM_OK, HMM_EINVAL, HMM_ECUDA, HMM_ENOMEM, HMM_ENODEV, HMM_EUNSUPPORTED = range(6)
MODE_AUTO, MODE_FAITHFUL, MODE_RING, MODE_GENERIC = 0, 1, 2, 3
MODES = {"auto": MODE_AUTO, "faithful": MODE_FAITHFUL, "ring": MODE_RING, "generic": MODE_GENERIC}


class HmmInfo(C.Structure):
    _fields_ = [
        ("engine", C.c_int32),
        ("n_chunks", C.c_int32),
        ("fwd_repaired", C.c_int32),
        ("bwd_repaired", C.c_int32),
        ("kernel_launches", C.c_int64),
        ("device_ms", C.c_double),
        ("kernel_ms", C.c_double),
        ("top_kernel_ms", C.c_double),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class HmmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libhmmcuda error {code}: {msg}")
        self.code = code


class HmmArgumentError(HmmError, ValueError):
    """HMM_EINVAL -- what the Julia shim raises as ArgumentError."""


_lib = None

# every symbol include/hmmcuda.h declares
EXPORTS = [
    "hmm_version", "hmm_last_error", "hmm_device_count", "hmm_set_device", "hmm_get_device", "hmm_set_ring_params",
    "hmm_viterbi_f64", "hmm_viterbi_ex_f64", "hmm_viterbi_batch_f64", "hmm_viterbi_dev_f64",
    "hmm_forward_f64", "hmm_backward_f64", "hmm_update_f64", "hmm_em_step_f64", "hmm_em_step_ex_f64",
    "hmm_train_create", "hmm_train_create_dev", "hmm_train_em_step", "hmm_train_run", "hmm_transition_weights", "hmm_train_destroy",
    "hmm_reconstruct_f64", "hmm_reconstruct_dev_f64", "hmm_unroll_mlseq_i16", "hmm_host_alloc", "hmm_host_free",
    "hmm_vshard_chunking", "hmm_vshard_create", "hmm_vshard_bvec", "hmm_vshard_forward", "hmm_vshard_fwd_boundary_get",
    "hmm_vshard_fwd_boundary_set", "hmm_vshard_fwd_verify", "hmm_vshard_trace", "hmm_vshard_trace_boundary_get",
    "hmm_vshard_trace_boundary_set", "hmm_vshard_trace_verify", "hmm_vshard_finish", "hmm_vshard_destroy",
    "hmm_vshard_repairs", "hmm_set_stream", "hmm_vshard_finish_ex", "hmm_vshard_summary_len", "hmm_vshard_summary_dev",
    "hmm_vshard_judge_dev", "hmm_release_workspace", "hmm_set_profiling", "hmm_vshard_p2p_init", "hmm_vshard_p2p_attach", "hmm_vshard_p2p_launch",
    "hmm_vshard_p2p_finish", "hmm_set_devices", "hmm_get_devices", "hmm_set_precision", "hmm_get_precision", "hmm_viterbi_f32",
    "hmm_viterbi_ex_f32", "hmm_viterbi_batch_f32", "hmm_viterbi_dev_f32", "hmm_measure_peaks", "hmm_vshard_set_y", "hmm_viterbi_rawfile", "hmm_emshard_create", "hmm_emshard_chunking", "hmm_emshard_stats_len",
    "hmm_emshard_boundary_len", "hmm_emshard_estep", "hmm_emshard_mstep", "hmm_emshard_destroy",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: libhmmcuda.so must be built (nvcc, sm_100a) -- there is no fallback path")
        L = C.CDLL(SO_PATH)
        L.hmm_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(L, name)  # raises AttributeError if a declared symbol is not exported
        _lib = L
    return _lib


def check(rc: int):
    if rc != HMM_OK:
        msg = lib().hmm_last_error().decode("utf-8", "replace")
        raise (HmmArgumentError if rc == HMM_EINVAL else HmmError)(rc, msg)
