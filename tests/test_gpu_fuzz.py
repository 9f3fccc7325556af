"""Randomised parity sweep (tools/fuzz_parity.py) as a test: random ring / overlap models, lengths, noise levels, rates
and chunkings through every engine that takes them, against the CPU oracle -- x identical, ll within 1e-9, one E/M
step within 1e-9 (weighted by a neuron's expected spike count).  A 240 s run of the same sweep (1 538 decodes per
engine, 262 E/M steps) is what found the end-of-recording cancellation fixed in ring_em.cu."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))


@pytest.mark.parametrize("seed", [11, 12])
def test_randomised_sweep_has_no_failures(hm, O, seed):
    import fuzz_parity

    stats = fuzz_parity.run(budget=25.0, seed=seed, max_cases=120)
    assert stats["cases"] >= 10 and not stats["failures"], stats["failures"][:3]
    assert stats["generic"] >= 10 and stats["ring"] >= 5


def test_wide_sweep_has_no_failures(hm, O):
    """tools/fuzz_wide.py: dense forward / backward of both engines, batches, three-neuron overlap models, E/M steps
    of overlap models, the library-side training loop against the oracle's loop, reconstruct / unroll."""
    import fuzz_wide

    stats = fuzz_wide.run(budget=25.0, seed=21, max_cases=400)
    assert stats["cases"] >= 20 and not stats["failures"], stats["failures"][:3]
