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


@pytest.mark.parametrize("n,k,e,bias,normalized", [(5924, 1024, 128, True, True), (37, 100, 7, False, True), (300, 1024, 256, True, False),
                                                   (65, 33, 160, True, True), (1, 1, 1, True, True)])
def test_project_normalize(n, k, e, bias, normalized, rs):
    """Embedding head (xw_plus_b + l2_normalize, src/networks.py:376-380, src/base_model_CUB.py:197-201) against a float64
    restatement; the TF graph computes in fp32, so the bar is fp32 round-off: 1e-5 of the row scale."""
    import multimodal_similarity_b200 as mm
    x = rs.randn(n, k).astype(np.float32)
    W = (rs.randn(k, e) / np.sqrt(k)).astype(np.float32)
    b = rs.randn(e).astype(np.float32) if bias else None
    y = x.astype(np.float64) @ W.astype(np.float64) + (0 if b is None else b.astype(np.float64))
    ref = y / np.sqrt(np.maximum((y * y).sum(1, keepdims=True), 1e-10)) if normalized else y
    got = mm.project_normalize(x, W, b, normalized=normalized)
    assert got.shape == (n, e) and got.dtype == np.float32
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    if normalized and n > 1:
        assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)


def test_project_normalize_zero_row_uses_epsilon():
    import multimodal_similarity_b200 as mm
    x = np.zeros((3, 8), np.float32)
    W = np.ones((8, 4), np.float32)
    got = mm.project_normalize(x, W)                      # y = 0 -> 0 * rsqrt(1e-10) = 0, no NaN (tf.nn.l2_normalize)
    assert np.array_equal(got, np.zeros((3, 4), np.float32))
