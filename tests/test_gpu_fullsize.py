"""The BASELINE.json configurations at FULL size, checked through size-independent properties and oracle spot checks
(the oracle cannot finish these sizes in seconds): sortedness, self-retrieval, query-subset consistency, sharded ==
unsharded, and exact agreement with the reference arithmetic on a handful of queries."""
import numpy as np
import pytest
import torch

from oracle import retrieval_np as O

pytestmark = pytest.mark.gpu


def synth(n, d, clusters, seed, device="cuda"):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cent = torch.randn(clusters, d, generator=g, device=device)
    lab = torch.randint(0, clusters, (n,), generator=g, device=device)
    x = cent[lab] + 0.5 * torch.randn(n, d, generator=g, device=device)
    return (x / x.norm(dim=1, keepdim=True)).contiguous(), lab


def spot_check(q, g, dist, idx, k, rows):
    """The oracle (NumPy restatement of retrieve_one) on the sampled queries against the FULL gallery, on a host thread pool
    (NumPy releases the GIL inside the distance pass and the sort)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    gn = g.cpu().numpy()
    qn = q[torch.as_tensor(rows, device=q.device)].cpu().numpy()
    dn, jn = dist[torch.as_tensor(rows, device=dist.device)].cpu().numpy(), idx[torch.as_tensor(rows, device=idx.device)].cpu().numpy()

    def one(n):
        ref_d, ref_i = O.knn(qn[n:n + 1], gn, k)
        return np.array_equal(dn[n], ref_d[0]), np.array_equal(jn[n], ref_i[0])

    with ThreadPoolExecutor(max(1, min(32, os.cpu_count() or 1))) as ex:
        res = list(ex.map(one, range(len(rows))))
    for r, (ok_d, ok_i) in zip(rows, res):
        assert ok_d, f"query {r}: distances differ from the reference arithmetic"
        assert ok_i, f"query {r}: indices differ"


def sampled_rows(n, count, fixed, seed=7):
    rs = np.random.RandomState(seed)
    return sorted(set(fixed) | set(rs.choice(n, count, replace=False).tolist()))


def test_cfg5_sharded_knn_1m_gallery():
    """config 5: 100k queries x 1M gallery, 128-d (the metric's width), top-100."""
    import multimodal_similarity_b200 as mm
    from multimodal_similarity_b200.retrieval import check_status, knn_raw
    g, _ = synth(1_000_000, 128, 1000, 12345)
    q, _ = synth(100_000, 128, 1000, 12346)
    q[:1000] = g[5000:6000]                                       # known answers: these queries ARE gallery rows
    dist, idx, status = knn_raw(q, g, 100)
    fell_back = check_status(status)
    print("exact-fallback queries:", fell_back)
    assert fell_back < 100
    assert torch.all(dist[:, 1:] >= dist[:, :-1]), "distances must be sorted"
    assert torch.all(idx >= 0) and torch.all(idx < 1_000_000)
    assert torch.all(idx[:1000, 0] == torch.arange(5000, 6000, device="cuda")) and torch.all(dist[:1000, 0] == 0)
    assert torch.all(idx.sort(dim=1).values[:, 1:] != idx.sort(dim=1).values[:, :-1]), "no gallery row twice"
    spot_check(q, g, dist, idx, 100, sampled_rows(100_000, 256, [0, 999, 1000, 31337, 99_999]))   # >= 256 of 100,000 queries
    # a subset of the queries gives the same rows (no cross-query interference in the tiles)
    d2, i2, st2 = knn_raw(q[40_000:40_300].contiguous(), g, 100)
    assert torch.equal(d2, dist[40_000:40_300]) and torch.equal(i2, idx[40_000:40_300])
    # gallery split in two shards + merge == unsharded
    from multimodal_similarity_b200.sharded import merge_parts
    qs = q[:5000].contiguous()
    packed = torch.empty((2, 2, 5000, 100), dtype=torch.int32, device="cuda")
    for r, (lo, hi) in enumerate(((0, 500_000), (500_000, 1_000_000))):
        d_, i_, _ = knn_raw(qs, g[lo:hi], 100)
        packed[r, 0] = d_.view(torch.int32)
        packed[r, 1] = i_
    md, mi = merge_parts(packed[:, 0].view(torch.float32), packed[:, 1], torch.tensor([0, 500_000], device="cuda"), 100)
    assert torch.equal(md, dist[:5000]) and torch.equal(mi, idx[:5000].long())


def test_cfg4_late_fusion_20k_x_200k():
    """config 4: camera + sensor, 2 x 128-d, 20k queries x 200k gallery, top-50, AP@50."""
    import multimodal_similarity_b200 as mm
    from multimodal_similarity_b200.retrieval import average_precision_at_k
    cam, lab = synth(220_000, 128, 7, 1)
    sen, _ = synth(220_000, 128, 7, 2)
    lab = lab.cpu().numpy()
    dist, idx = mm.retrieve(cam[:20_000], cam[20_000:], 50, queries2=sen[:20_000], gallery2=sen[20_000:])
    assert torch.all(dist[:, 1:] >= dist[:, :-1])
    fused = mm.late_fusion(cam, sen)
    assert fused.shape == (220_000, 256)
    spot_check(fused[:20_000], fused[20_000:], dist, idx, 50, sampled_rows(20_000, 256, [0, 7, 19_999]))
    # d^2_fused = d^2_cam + d^2_sens (App. A.4 corollary): the fused distance of the top hit decomposes exactly in float32
    r = 123
    j = int(idx[r, 0]) + 20_000
    dc = O.pairwise_sum_f32(((cam[r] - cam[j]).cpu().numpy() ** 2).astype(np.float32))
    ds = O.pairwise_sum_f32(((sen[r] - sen[j]).cpu().numpy() ** 2).astype(np.float32))
    assert np.float32(np.sqrt(np.float32(dc + ds))) == dist[r, 0].item()
    rel = lab[20_000:][idx.cpu().numpy()] == lab[:20_000, None]
    ap = average_precision_at_k(rel, np.array([(lab[20_000:] == l).sum() for l in lab[:20_000]]))
    assert 0.0 <= np.nanmean(ap) <= 1.0


def test_cfg3_cub_recall_5924():
    """config 3: 5,924 x 128-d embeddings projected from 1024-d GoogleNet-shaped features, labels 101..200."""
    import multimodal_similarity_b200 as mm
    rs = np.random.RandomState(12345)
    lab = (np.arange(5924) % 100 + 101).astype(np.int32)
    rs.shuffle(lab)
    feats = rs.randn(100, 1024)[lab - 101] + rs.randn(5924, 1024)
    emb = (feats @ (rs.randn(1024, 128) / 32)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    mAP, mAP_event, mPrec, confusion, count, recall = mm.evaluate(emb, lab)
    assert len(recall) == 6 and all(recall[i] <= recall[i + 1] for i in range(5))      # R@1 <= R@2 <= ... <= R@32
    assert len(mAP_event) == 100 and 0 < mAP <= 1 and confusion["confusion_matrix"].shape == (100, 100)
    assert np.array_equal(count[1:, 0], np.bincount(lab - 101)[1:])
    s = mm.evaluate_simple(emb, lab)
    assert s[0] == pytest.approx(mAP, abs=1e-12) and s[2] == pytest.approx(recall[0], abs=1e-12)
    # the oracle on a 600-row prefix must agree exactly with the GPU on the same prefix (full 5,924 takes ~30 s on a CPU)
    ref = O.evaluate(emb[:600], lab[:600])
    got = mm.evaluate(emb[:600], lab[:600])
    assert got[0] == pytest.approx(ref[0], abs=1e-12) and got[5] == ref[5]
    # aligned=True (the intended label lookup) can only help R@K
    fixed = mm.evaluate(emb, lab, aligned=True)
    assert fixed[0] == pytest.approx(mAP, abs=1e-12)


def test_cfg1_cfg2_losses_at_config_size_are_consistent():
    """configs 1/2: gradient of the fused kernels == finite differences of their own loss (size-independent property)."""
    import multimodal_similarity_b200 as mm
    for kind, n, margin in (("bh", 256, "soft"), ("lifted", 512, 1.0)):
        e, lab = synth(n, 128, 32 if kind == "bh" else 7, 3)
        pids = (lab % (32 if kind == "bh" else 7)).float() + (1.0 if kind == "bh" else 0.0)
        fn = mm.batch_hard if kind == "bh" else mm.lifted_loss
        x = e.clone().requires_grad_(True)
        out = fn(x, pids, margin)
        out[0].backward()
        g = x.grad
        v = g / g.norm()                      # steepest direction: the directional derivative is ||g||, well above fp32 noise
        eps = 1e-3
        lp = float(fn(e + eps * v, pids, margin)[0])
        lm = float(fn(e - eps * v, pids, margin)[0])
        fd = (lp - lm) / (2 * eps)
        an = float((g * v).sum())
        assert an > 0 and an == pytest.approx(fd, rel=2e-2), (kind, an, fd)
