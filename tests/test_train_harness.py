"""Pins the restated host-side merge / prune code (tests/train_harness.py) to the reference's own known answers
(test/runtests.jl:44-61 `overlap and combine`), on CPU."""
import numpy as np

import train_harness as th


def test_find_best_overlap_known_answers(hm):
    mu = np.array([[1.0, 1.0], [2.0, 2.0], [3.0, 3.0]])
    xi, xm = th.find_best_overlap(mu, 0, 1)
    assert (list(xi[0]), list(xi[1])) == ([0, 1, 2], [0, 1, 2]) and np.isclose(xm, 14.0)
    t1 = hm.create_spike_template(60, 3.0, 0.8, 0.2)
    t2 = np.zeros_like(t1)
    t2[4:] = t1[:56]
    xi, xm = th.find_best_overlap(np.stack([t1, t2], axis=1), 0, 1)
    assert xi[0] == range(0, 56) and xi[1] == range(4, 60)          # 1:56 and 5:60 in the reference's indices
    assert np.isclose(xm, 100.66411692920131, rtol=1e-13)
    c = th.condense_candidates(np.stack([t1, t2], axis=1), 0.1)     # sigma^2 = 0.1 as in the reference's test
    assert c is not None and c[0] == (0, 1) and c[2][0] == range(0, 56) and c[2][1] == range(4, 60)


def test_merge_quirks_are_reproduced(hm):
    """A merge leaves the LAST column of the new template matrix zero and drops the last template (N -= 1 before
    setdiff), and `.=+` assigns: the merged template is half of template i2 on the overlap rows of i2."""
    K = 20
    t = hm.create_spike_template(K, 3.0, 0.8, 0.2)
    u = hm.create_spike_template(K, 2.0, 0.3, 0.2)
    mu = np.stack([t, t * 1.0001, u], axis=1)
    mu[0, :] = 0
    sm = hm.StateMatrix(3, K, np.log([0.01, 0.02, 0.005]), False)
    sm2, mu2 = th.condense_templates(hm.StateMatrix, sm, mu, 0.3, 0.05)
    assert sm2.N == 2 and mu2.shape == (K, 2)
    assert np.allclose(mu2[:, 0], 0.5 * mu[:, 1]) and np.all(mu2[:, 1] == 0.0)
