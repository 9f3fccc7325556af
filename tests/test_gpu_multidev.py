"""hmm_set_devices: multi-GPU dispatch INSIDE the library, behind the unchanged host-pointer entry points.  On a box
with one GPU the dispatcher is driven with several workers on that one device (HMMCUDA_DEBUG_ALLOW_DUP_DEVICES);
with more GPUs the same tests use distinct devices.  Results must be those of the single-device decode."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def devices(hm, monkeypatch):
    n = hm.device_count()
    monkeypatch.setenv("HMMCUDA_DEBUG_ALLOW_DUP_DEVICES", "1")
    devs = list(range(min(n, 8))) if n >= 2 else [0, 0, 0]
    yield devs
    hm.set_devices(None)


def test_batch_is_split_over_the_devices(hm, O, case_factory, devices):
    cases = [case_factory(4, 48, 300_000, 300 + c) for c in range(7)]
    Y = np.asfortranarray(np.stack([c[0] for c in cases], axis=1))
    models = [(c[1], c[2], c[3]) for c in cases]
    x1, ll1 = hm.viterbi_batch(Y, models, mode="ring")
    hm.set_devices(devices)
    x, ll, info = hm.viterbi_batch(Y, models, mode="ring", return_info=True)
    assert np.array_equal(x, x1) and np.array_equal(ll, ll1)
    xo, llo = O.viterbi(Y[:, 3], *models[3])
    assert np.array_equal(x[:, 3], xo) and abs(ll[3] - llo) <= 1e-9 * abs(llo)


def test_one_long_recording_is_time_sharded_over_the_devices(hm, O, case_factory, devices):
    """T >= 2^23 samples and a ring model: hmm_viterbi_f64 itself cuts the recording into one shard per device
    (ghost chunks, peer-memory summary exchange) -- one C call drives all of them."""
    T = (1 << 23) + 54_321
    S, lA, mu, sig = case_factory(3, 60, T, 77)
    x1, ll1 = hm.viterbi(S, lA, mu, sig)
    hm.set_devices(devices)
    x, ll, info = hm.viterbi(S, lA, mu, sig, return_info=True)
    assert info["engine"] == 2
    assert np.array_equal(x, x1)
    assert abs(ll - ll1) <= 1e-12 * abs(ll1)
    xo, llo = O.viterbi(S, lA, mu, sig)
    assert np.array_equal(x, xo) and abs(ll - llo) <= 1e-9 * abs(llo)


def test_set_devices_validation(hm):
    with pytest.raises(hm.HmmArgumentError):
        hm.set_devices([0, 99])
    hm.set_devices(None)
