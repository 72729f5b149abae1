"""K1: cdist(all_diffs(a, b)) on the GPU is bit-identical to the reference's NumPy twin (golden + oracle)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import retrieval_np as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "dist_d*.npz"))))
def test_golden_bit_exact(path):
    import multimodal_similarity_b200 as mm
    g = np.load(path)
    diff = mm.all_diffs(g["a"], g["b"])
    assert diff.shape == (g["a"].shape[0], g["b"].shape[0], g["a"].shape[1])
    for metric, key in (("squaredeuclidean", "sq"), ("euclidean", "eu"), ("l1", "l1")):
        got = mm.cdist(diff, metric)
        assert isinstance(got, np.ndarray) and got.dtype == np.float32
        assert np.array_equal(got, g[key]), f"{metric}: {np.abs(got - g[key]).max()}"


@pytest.mark.parametrize("m,n,d", [(1, 1, 1), (3, 5, 7), (130, 70, 96), (257, 300, 128), (64, 64, 300), (33, 31, 1500), (16, 64, 4000)])
def test_random_vs_oracle(m, n, d, rs):
    import multimodal_similarity_b200 as mm
    a = rs.randn(m, d).astype(np.float32)
    b = rs.randn(n, d).astype(np.float32)
    for metric in ("squaredeuclidean", "euclidean", "l1"):
        assert np.array_equal(mm.pairwise_distance(a, b, metric), O.cdist(O.all_diffs(a, b), metric)), metric


def test_torch_in_torch_out_and_self_distance(rs):
    import multimodal_similarity_b200 as mm
    a = torch.from_numpy(rs.randn(100, 128).astype(np.float32)).cuda()
    d = mm.cdist_tf(mm.all_diffs_tf(a, a)).materialize()
    assert d.is_cuda and d.shape == (100, 100)
    assert torch.all(torch.diagonal(d) == 0)                   # D_ii == 0 exactly in difference form (App. A.1)
    assert torch.equal(d, d.T)
    with pytest.raises(NotImplementedError):
        mm.pairwise_distance(a, a, "cosine")
