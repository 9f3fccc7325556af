"""The C-ABI library loads, exports every symbol include/hmmcuda.h declares, and
fails loudly (no CPU fallback) when there is no device.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hmmcuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hmm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(hm):
    L = hm.lib()
    decl = _declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/hmmcuda.h but not exported"
    assert sorted(hm._lib.EXPORTS) == decl


def test_version_and_error_string(hm):
    L = hm.lib()
    assert L.hmm_version() >= 100
    assert isinstance(L.hmm_last_error(), bytes)


def test_trans_record_layout(hm):
    assert hm.TRANS_DTYPE.itemsize == 24
    assert [hm.TRANS_DTYPE.fields[k][1] for k in ("src", "dst", "lp")] == [0, 8, 16]


def test_no_device_fails_loudly(hm):
    """On a machine without a GPU every compute entry point must return HMM_ENODEV
    -- never silently compute on the CPU."""
    if hm.device_count() > 0:
        pytest.skip("a CUDA device is present")
    lA = hm.StateMatrix(2, 5, np.log([0.01, 0.02]), False)
    mu = np.asfortranarray(np.zeros((5, 2)))
    y = np.zeros(64)
    for call in (lambda: hm.viterbi(y, lA, mu, 0.3), lambda: hm.forward(y, lA, mu, 0.3),
                 lambda: hm.em_step(y, lA, mu, 0.3),
                 lambda: hm.reconstruct_signal(np.ones(4, dtype=np.int16), lA, mu)):
        with pytest.raises(hm.HmmError) as ei:
            call()
        assert ei.value.code == hm._lib.HMM_ENODEV
        assert "no CPU fallback" in str(ei.value)


def test_python_side_argument_errors(hm):
    lA = hm.StateMatrix(2, 5, np.log([0.01, 0.02]), False)
    with pytest.raises(hm.HmmArgumentError):
        hm.viterbi(np.zeros(10), lA, np.zeros((4, 2)), 0.3)  # wrong K
    with pytest.raises(hm.HmmArgumentError):
        hm.reconstruct_signal(np.array([0, 1]), lA, np.asfortranarray(np.zeros((5, 2))))
    with pytest.raises(hm.HmmArgumentError):
        hm.viterbi(np.zeros((3, 3)), lA, np.asfortranarray(np.zeros((5, 2))), 0.3)


def test_product_does_not_reach_the_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "hmmspikesorter.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "hmm_oracle" not in txt and "import oracle" not in txt, f


def test_transition_weights_host_helper_matches_the_constructor(hm, O):
    """hmm_transition_weights (host only, what hmm_train_run applies between E/M steps): for ring and overlap layouts
    and random lp it must give the records the StateMatrix constructor builds (src/types.jl:94-127) -- bit for bit
    those of the oracle's constructor (same libm), within an ulp of log1p those of the numpy mirror -- including lpz
    over the WHOLE lp vector for overlap models, and refuse a degenerate lp."""
    import numpy as np
    rng = np.random.default_rng(5)
    for N, K, overlap in [(3, 12, False), (1, 5, False), (5, 9, False), (2, 7, True), (3, 5, True)]:
        lA = hm.StateMatrix(N, K, np.log(np.full(N, 0.01)), overlap)
        nlp = int((lA.transitions["src"] == 1).sum()) - 1
        for _ in range(5):
            lp = np.log(rng.uniform(1e-6, 0.05, size=nlp))
            tr = hm.transition_weights(lA, lp)
            ref = hm.StateMatrix(N, K, lp, overlap).transitions
            assert tr is not None and np.array_equal(tr["src"], ref["src"]) and np.array_equal(tr["dst"], ref["dst"])
            assert np.allclose(tr["lp"], ref["lp"], rtol=1e-14, atol=0)
            assert np.array_equal(tr["lp"], O.OracleStateMatrix(N, K, lp, overlap).transitions["lp"])
        lp = np.log(np.full(nlp, 0.01))
        lp[0] = -np.inf
        assert hm.transition_weights(lA, lp) is None
