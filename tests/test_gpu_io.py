"""I/O front-end (hmm_viterbi_rawfile): raw recording file -> pinned staging -> HBM -> decode, and the .mat result the
reference's CLI writes (src/hmmsort.jl:36-104).  The decode of the file must equal the decode of the same values handed
over as arrays."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,interleaved", [(np.int16, True), (np.float32, True), (np.float64, False), (np.int16, False)])
def test_rawfile_decode_equals_array_decode(hm, O, case_factory, tmp_path, dtype, interleaved):
    nfile, T, N, K = 6, 300_000, 3, 60
    cases = [case_factory(N, K, T, 700 + c) for c in range(nfile)]
    scale = 1.0 / 2048 if dtype == np.int16 else 1.0
    raw = np.stack([np.round(c[0] / scale).astype(dtype) if dtype == np.int16 else c[0].astype(dtype) for c in cases], axis=1)
    path = tmp_path / "rec.bin"
    header = b"\x00" * 1000  # (a contiguous HDF5 dataset is a raw block at a byte offset)
    with open(path, "wb") as f:
        f.write(header)
        f.write(np.ascontiguousarray(raw if interleaved else raw.T).tobytes())
    pick = [4, 1, 5]
    models = [(cases[c][1], cases[c][2], cases[c][3]) for c in pick]
    x, ll, info = hm.viterbi_rawfile(path, dtype, nfile, T, pick, models, interleaved=interleaved, scale=scale,
                                     byte_offset=len(header), mode="ring", return_info=True)
    assert info["engine"] == 2
    Y = np.asfortranarray(raw[:, pick].astype(np.float64) * scale)
    x2, ll2 = hm.viterbi_batch(Y, models, mode="ring")
    assert np.array_equal(x, x2) and np.array_equal(ll, ll2)
    xo, llo = O.viterbi(Y[:, 0], *models[0])
    assert np.array_equal(x[:, 0], xo) and abs(ll[0] - llo) <= 1e-9 * abs(llo)


def test_sort_result_mat_file(hm, O, case_factory, tmp_path):
    from scipy.io import loadmat

    S, lA, mu, sig = case_factory(2, 30, 50_000, 710)
    x, ll = hm.viterbi(S, lA, mu, sig)
    out = hm.save_sort_result(tmp_path / "sorted.mat", x, lA, mu, sig, ll)
    m = loadmat(tmp_path / "sorted.mat")
    assert set(["mlseq", "ll", "waveforms", "lp", "sigma"]) <= set(m.keys())
    assert np.array_equal(m["mlseq"], O.unroll_mlseq(x, lA)) and m["mlseq"].shape == (2, 50_000)
    assert np.allclose(m["waveforms"], mu) and np.isclose(float(m["sigma"].ravel()[0]), sig)


def test_rawfile_errors(hm, case_factory, tmp_path):
    S, lA, mu, sig = case_factory(2, 30, 40_000, 711)
    path = tmp_path / "short.bin"
    S[:1000].astype(np.float32).tofile(path)
    with pytest.raises(hm.HmmArgumentError):
        hm.viterbi_rawfile(path, np.float32, 1, 40_000, [0], [(lA, mu, sig)])
    with pytest.raises(hm.HmmArgumentError):
        hm.viterbi_rawfile(tmp_path / "missing.bin", np.float32, 1, 40_000, [0], [(lA, mu, sig)])
