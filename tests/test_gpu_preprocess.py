"""-m gpu: input-side kernels (seist_normalize / seist_dpk_labels through the C-ABI) against the numpy oracle that is
bit-exactly pinned to the reference's own sources (tests/test_cpu_preprocess.py).  Floating point: 2e-6."""
import numpy as np
import pytest
import torch

from oracle import preprocess_ref as PR
from seist_b200 import preprocess as PP

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["std", "max", ""])
def test_normalize_matches_oracle(mode):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((6, 3, 8192)) * rng.uniform(0.1, 30, (6, 3, 1)) + rng.uniform(-5, 5, (6, 3, 1))).astype(np.float32)
    x[1, 2] = 4.25                                       # constant trace: zero scale -> 1
    want = np.stack([PR.normalize(t, mode) for t in x])
    got = PP.normalize_(torch.from_numpy(x.copy()).cuda(), mode).cpu().numpy()
    assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("shape,width", [("gaussian", 25), ("triangle", 50), ("box", 11)])
def test_dpk_labels_match_oracle(shape, width):
    L = 8192
    A = PP.ABSENT
    cases = [([300], [700]), ([5], [60]), ([8100], [8185]), ([], []), ([100, 4000], [400, 5200]), ([800], []), ([], [500]),
             ([8191], [8191 + 30]), ([0], [3]), ([1000, 3000, 5000], [1500, 3600])]
    K = 3
    pp = torch.tensor([c[0] + [A] * (K - len(c[0])) for c in cases])
    ss = torch.tensor([c[1] + [A] * (K - len(c[1])) for c in cases])
    got = PP.dpk_soft_labels(pp.cuda(), ss.cuda(), L, width, shape, 1.4).cpu().numpy()
    for i, (p, s) in enumerate(cases):
        want = PR.dpk_labels(p, s, L, width, shape, 1.4)
        assert np.abs(got[i] - want).max() <= 2e-6, (i, p, s, np.abs(got[i] - want).max())
    assert got.min() >= 0.0 and got.max() <= 2.0 + 1e-6


def test_cpu_tensor_fails_loudly():
    with pytest.raises(RuntimeError):
        PP.normalize_(torch.zeros(1, 3, 64))
