"""K2/K3: fused batch-hard / lifted loss forward + backward against the torch-CPU restatement of the TF graph.
Tolerances are the north-star ones: loss 1e-4 relative, gradients 1e-3 (relative to the gradient scale);
mined indices exact outside exact-distance ties."""
import numpy as np
import pytest
import torch

from conftest import clustered
from oracle import losses_torch as L

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_RTOL = 1e-3


def run_ours(kind, emb, pids, margin, weighted=True):
    import multimodal_similarity_b200 as mm
    e = torch.from_numpy(emb).cuda().requires_grad_(True)
    fn = mm.batch_hard if kind == "batch_hard" else mm.lifted_loss
    out = fn(e, torch.from_numpy(pids).cuda(), margin, weighted)
    out[0].backward()
    return out, e.grad.detach().cpu().numpy()


def compare(kind, emb, pids, margin, weighted=True):
    out, grad = run_ours(kind, emb, pids, margin, weighted)
    ref32, g32, d32 = L.loss_and_grad(kind, torch.from_numpy(emb), torch.from_numpy(pids), margin, weighted)
    ref64, g64, _ = L.loss_and_grad(kind, torch.from_numpy(emb).double(), torch.from_numpy(pids).double(), margin, weighted)
    loss = float(out[0])
    assert loss == pytest.approx(float(ref64[0]), rel=LOSS_RTOL), (loss, float(ref32[0]), float(ref64[0]))
    names = ["diff", "weights", "furthest_positive", "closest_negative"]
    for n, a, b in zip(names, out[2:], ref64[2:]):
        a, b = a.detach().cpu().numpy().astype(np.float64), b.detach().numpy()
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), n
        assert np.allclose(a[fin], b[fin], rtol=LOSS_RTOL, atol=1e-6), (n, np.abs(a[fin] - b[fin]).max())
    na = float(out[1])
    assert na == pytest.approx(float(ref64[1]), abs=1.5 / len(pids))       # a row within 1e-5 of the threshold may flip
    scale = max(np.abs(g64.numpy()).max(), 1e-12)
    err = np.abs(grad - g64.numpy()).max() / scale
    assert err < GRAD_RTOL, f"gradient error {err:.2e} of scale {scale:.3e}"
    return out, d32


def batch(rs, n_classes, per_class, d=128, background=False, noise=0.5):
    cent = rs.randn(n_classes, d).astype(np.float32)
    lab = np.repeat(np.arange(1, n_classes + 1), per_class)
    x = cent[lab - 1] + noise * rs.randn(lab.size, d).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if background:
        lab = lab.copy()
        lab[lab == 1] = 0
    p = rs.permutation(lab.size)
    return x[p].astype(np.float32), lab[p].astype(np.float32)


def test_kats():
    """SURVEY.md Appendix B (hand-checked)."""
    E = np.array([[0, 0], [1, 0], [0, 2], [3, 0]], np.float32)
    out, g = run_ours("batch_hard", E, np.array([1, 1, 2, 2], np.float32), 0.2)
    assert float(out[0]) == pytest.approx(4.6, rel=1e-6)
    assert np.allclose(g, [[0, 1], [1, 0], [-3, 1], [2, -2]], atol=1e-5)
    assert out[4].tolist() == [1, 1, 13, 13] and out[5].tolist() == [4, 4, 4, 4]
    assert out.pos_idx.tolist() == [1, 0, 3, 2] and out.neg_idx.tolist() == [2, 3, 0, 1]
    out, g = run_ours("batch_hard", E, np.array([1, 1, 2, 2], np.float32), "soft")
    assert float(out[0]) == pytest.approx(4.524355377, rel=1e-6)
    out, g = run_ours("lifted", E, np.array([1, 1, 2, 2], np.float32), 1.0)
    assert float(out[0]) == pytest.approx(5.079997649, rel=1e-6)
    assert np.allclose(g, [[0.0100392764, 0.7310585786], [0.8588364384, 0.2689414214],
                           [-2.8655089465, 0.9999864381], [1.9966332317, -1.9999864381]], atol=1e-5)
    assert float(out[1]) == 1.0
    out, g = run_ours("batch_hard", E, np.array([0, 1, 1, 2], np.float32), "soft")
    assert float(out[0]) == pytest.approx(1.531039002, rel=1e-6)
    assert out[3].tolist() == pytest.approx([0, 2 / 7, 2 / 7, 3 / 7])
    assert np.allclose(g, [[0.5611507372, 0.8354955184], [0.4485812620, -1.9577969928],
                           [-0.9788984964, 1.1223014743], [-0.0308335028, 0]], atol=1e-5)


@pytest.mark.parametrize("margin", ["soft", 0.2])
@pytest.mark.parametrize("background", [False, True])
def test_batch_hard_cfg1(margin, background, rs):
    """BASELINE config 1: batch 256 = 32 classes x 8, 128-d."""
    emb, pids = batch(rs, 32, 8, background=background)
    out, d32 = compare("batch_hard", emb, pids, margin)
    p_ref, n_ref = L.mined_indices(d32, torch.from_numpy(pids))
    dd = d32.numpy()
    for name, got, ref in (("pos", out.pos_idx, p_ref), ("neg", out.neg_idx, n_ref)):
        got, ref = got.cpu().numpy(), ref.numpy()
        bad = np.nonzero(got != ref)[0]
        for i in bad:     # only acceptable when the two candidates are an (almost) exact distance tie in fp32
            assert abs(dd[i, got[i]] - dd[i, ref[i]]) <= 4 * np.finfo(np.float32).eps * max(dd[i, ref[i]], 1e-6), (name, i)


@pytest.mark.parametrize("n,counts", [(512, {0: 200, 1: 160, 2: 50, 3: 50, 4: 25, 5: 20, 6: 7})])
def test_lifted_cfg2(n, counts, rs):
    """BASELINE config 2: batch 512, HDD-style class histogram incl. background, margin 1.0."""
    lab = np.concatenate([np.full(c, l) for l, c in counts.items()])
    cent = rs.randn(7, 128).astype(np.float32)
    x = cent[lab] + 0.5 * rs.randn(n, 128).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    p = rs.permutation(n)
    compare("lifted", x[p].astype(np.float32), lab[p].astype(np.float32), 1.0)


@pytest.mark.parametrize("kind,margin", [("batch_hard", "soft"), ("batch_hard", 0.5), ("lifted", 1.0), ("lifted", 0.2)])
@pytest.mark.parametrize("n,d", [(5, 3), (33, 17), (100, 64), (257, 128), (384, 130), (700, 32), (1024, 128)])
def test_shapes(kind, margin, n, d, rs):
    emb = rs.randn(n, d).astype(np.float32) * 0.3
    pids = rs.randint(0, 5, size=n).astype(np.float32)
    pids[:5] = np.arange(5)                                   # every class present
    compare(kind, emb, pids, margin)
    compare(kind, emb, pids, margin, weighted=False)


def test_edge_rows_and_ties():
    # a class with a single member (no positive), a row with no negatives is impossible unless single class
    E = np.array([[0, 0], [1, 0], [-1, 0], [0, 5], [0, 1]], np.float32)
    pids = np.array([1, 1, 1, 2, 3], np.float32)
    out, _ = compare("batch_hard", E, pids, 0.2, weighted=False)          # row 0: two furthest positives at distance 1
    assert out.pos_idx.tolist()[3] == -1 and float(out[4][3]) == 0.0
    compare("lifted", E, pids, 1.0, weighted=False)
    # single class: no negatives anywhere -> loss 0, closest_negative = +inf / -inf, zero gradient
    ones = np.ones(5, np.float32)
    out, g = run_ours("batch_hard", E, ones, 0.2, weighted=False)
    assert float(out[0]) == 0.0 and torch.isinf(out[5]).all() and not g.any()
    out, g = run_ours("lifted", E, ones, 1.0, weighted=False)
    assert float(out[0]) == 0.0 and not g.any()
    # duplicated embeddings: exact distance ties between negatives
    E2 = np.array([[0, 0], [2, 0], [1, 1], [1, 1], [1, -1]], np.float32)
    compare("batch_hard", E2, np.array([1, 1, 2, 2, 3], np.float32), 0.2, weighted=False)


def test_reference_call_site_form_and_forward_only(rs):
    """The three lines of src/base_model_batchhard.py:115-124 work unchanged."""
    import multimodal_similarity_b200 as mm
    emb, pids = batch(rs, 8, 4)
    e = torch.from_numpy(emb).cuda().requires_grad_(True)
    diffs = mm.all_diffs_tf(e, e)
    all_dist = mm.cdist_tf(diffs)
    loss, num_active, diff, weights, fp, cn = mm.batch_hard(all_dist, torch.from_numpy(pids).cuda(), "soft")
    loss.backward()
    direct = mm.batch_hard(torch.from_numpy(emb).cuda(), pids, "soft")       # no grad requested: forward only
    assert float(direct[0]) == float(loss) and e.grad is not None
    assert torch.equal(direct[2], diff)


def test_deterministic_loss(rs):
    emb, pids = batch(rs, 32, 8)
    a, _ = run_ours("batch_hard", emb, pids, "soft")
    b, _ = run_ours("batch_hard", emb, pids, "soft")
    assert float(a[0]) == float(b[0]) and torch.equal(a[2], b[2])


# ---------------------------------------------------------------------------------------------- tf.contrib triplet_semihard_loss
@pytest.mark.parametrize("n,d,classes,margin", [(4, 2, 2, 0.5), (64, 16, 5, 1.0), (256, 128, 32, 0.2), (300, 64, 100, 0.2), (513, 128, 7, 1.0)])
def test_contrib_triplet_semihard_vs_oracle(n, d, classes, margin):
    """Device forward + backward against the torch restatement of the TF 1.x source (autograd gives the TF gradient).
    Bars as for the other losses: 1e-4 relative on the loss, 1e-3 of the gradient scale."""
    import multimodal_similarity_b200 as mm
    from oracle import losses_torch as L
    rs = np.random.RandomState(n)
    if n == 4:
        x = np.array([[0., 0.], [1., 0.], [0., 2.], [3., 0.]], np.float32)
        lab = np.array([1, 1, 2, 2], np.int32)
    else:
        cent = rs.randn(classes, d).astype(np.float32)
        lab = rs.randint(0, classes, n).astype(np.int32)
        x = cent[lab] + 0.7 * rs.randn(n, d).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    e = torch.from_numpy(x).cuda().requires_grad_(True)
    loss = mm.triplet_semihard_loss(torch.from_numpy(lab).cuda(), e, margin)
    loss.backward()
    e64 = torch.from_numpy(x).double().requires_grad_(True)
    ref = L.contrib_triplet_semihard_loss(torch.from_numpy(lab).long(), e64, margin)
    ref.backward()
    assert float(loss) == pytest.approx(float(ref), rel=1e-4, abs=1e-6)
    g, gr = e.grad.cpu().numpy(), e64.grad.numpy()
    assert np.abs(g - gr).max() <= 1e-3 * max(np.abs(gr).max(), 1e-6)
    if n == 4:
        assert float(loss) == pytest.approx(3.25, rel=1e-6)


def test_contrib_triplet_semihard_numpy_inputs_and_errors():
    import multimodal_similarity_b200 as mm
    x = np.array([[0., 0.], [1., 0.]], np.float32)
    assert float(mm.triplet_semihard_loss(np.array([7, 7]), x, 0.5)) == pytest.approx(1.5, rel=1e-6)   # no negatives at all
    with pytest.raises(ValueError):
        mm.triplet_semihard_loss(np.array([1, 2, 3]), x)


@pytest.mark.parametrize("n,d,classes,margin", [(4, 2, 2, 1.0), (64, 16, 5, 1.0), (256, 128, 32, 0.2), (300, 64, 100, 1.0), (513, 128, 7, 1.0)])
def test_contrib_lifted_struct_vs_oracle(n, d, classes, margin):
    import multimodal_similarity_b200 as mm
    from oracle import losses_torch as L
    rs = np.random.RandomState(n + 1)
    if n == 4:
        x = np.array([[0., 0.], [1., 0.], [0., 2.], [3., 0.]], np.float32)
        lab = np.array([1, 1, 2, 2], np.int32)
    else:
        cent = rs.randn(classes, d).astype(np.float32)
        lab = rs.randint(0, classes, n).astype(np.int32)
        x = cent[lab] + 0.7 * rs.randn(n, d).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    e = torch.from_numpy(x).cuda().requires_grad_(True)
    loss = mm.lifted_struct_loss(torch.from_numpy(lab).cuda(), e, margin)
    loss.backward()
    e64 = torch.from_numpy(x).double().requires_grad_(True)
    ref = L.contrib_lifted_struct_loss(torch.from_numpy(lab).long(), e64, margin)
    ref.backward()
    assert float(loss) == pytest.approx(float(ref), rel=1e-4, abs=1e-6)
    g, gr = e.grad.cpu().numpy(), e64.grad.numpy()
    assert np.abs(g - gr).max() <= 1e-3 * max(np.abs(gr).max(), 1e-6)


def test_layouts_of_different_sizes_interleave():
    """The shared-memory limit is a per-kernel attribute while the need depends on (N, D): a small problem between two
    large ones must not lower it (regression: the cached 256 x 128 launch failed with 'invalid argument')."""
    import multimodal_similarity_b200 as mm
    g = torch.Generator().manual_seed(0)
    big = torch.randn(256, 128, generator=g).cuda()
    small = torch.randn(64, 16, generator=g).cuda()
    pb = (torch.arange(256) % 32 + 1).float().cuda()
    ps = (torch.arange(64) % 4 + 1).float().cuda()
    a = float(mm.batch_hard(big, pb, "soft")[0])
    mm.batch_hard(small, ps, "soft"); mm.lifted_loss(small, ps, 1.0)
    assert float(mm.batch_hard(big, pb, "soft")[0]) == a
    b = float(mm.lifted_loss(big, pb, 1.0)[0])
    mm.lifted_loss(small, ps, 1.0)
    assert float(mm.lifted_loss(big, pb, 1.0)[0]) == b
