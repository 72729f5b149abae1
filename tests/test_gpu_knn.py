"""K4: fused tcgen05 kNN + exact re-rank against the oracle (the reference's per-query distance + argsort)."""
import os

import numpy as np
import pytest
import torch

from conftest import clustered, golden
from oracle import retrieval_np as O

pytestmark = pytest.mark.gpu


def assert_knn_equal(dist, idx, ref_d, ref_i):
    """Distances bit-exact; indices exact outside exact-distance ties (argsort is unstable in the reference)."""
    assert dist.dtype == np.float32 and idx.dtype == np.int64
    assert np.array_equal(dist, ref_d), f"max |d - ref| = {np.nanmax(np.abs(dist - ref_d))}"
    neq = idx != ref_i
    if neq.any():
        # a differing index must sit inside a run of equal distances (possibly cut by the k boundary)
        q, r = np.nonzero(neq)
        tied = np.zeros_like(neq)
        tied[:, 1:] |= ref_d[:, 1:] == ref_d[:, :-1]
        tied[:, :-1] |= ref_d[:, :-1] == ref_d[:, 1:]
        tied[:, -1] = True
        assert tied[q, r].all(), f"{neq.sum()} index mismatches outside ties"


CASES = [
    # nq, ng, d, k, classes
    (300, 1000, 128, 10, 7),
    (129, 4000, 256, 100, 20),
    (64, 777, 64, 5, 3),
    (200, 3000, 160, 50, 10),      # 128+32 late fusion width (evaluate_hallucination.py) -> padded to 192
    (1, 300, 128, 112, 2),
    (500, 257, 96, 16, 5),
    (100, 20000, 128, 100, 50),
]


@pytest.mark.parametrize("nq,ng,d,k,c", CASES)
def test_knn_vs_oracle(nq, ng, d, k, c, rs):
    import multimodal_similarity_b200 as mm
    g, _ = clustered(rs, ng, d, c)
    q, _ = clustered(rs, nq, d, c)
    q[: min(nq, 5)] = g[: min(nq, 5)]                      # exact duplicates: zero distances
    dist, idx = mm.retrieve(q, g, k)
    ref_d, ref_i = O.knn(q, g, k)
    assert_knn_equal(dist, idx, ref_d, ref_i)


def test_unnormalised_and_exclude_self(rs):
    import multimodal_similarity_b200 as mm
    x = (rs.randn(1500, 128) * 3.0 + 1.0).astype(np.float32)     # --no_normalized (App. A.1): not unit norm
    dist, idx = mm.retrieve(x, x, 20, exclude_self=True)
    ref_d, ref_i = O.knn(x, x, 20, exclude_self=True)
    assert_knn_equal(dist, idx, ref_d, ref_i)
    assert not (idx == np.arange(1500)[:, None]).any()


def test_gallery_smaller_than_k(rs):
    import multimodal_similarity_b200 as mm
    g = rs.randn(7, 128).astype(np.float32)
    q = rs.randn(40, 128).astype(np.float32)
    dist, idx = mm.retrieve(q, g, 10)
    ref_d, ref_i = O.knn(q, g, 7)
    assert_knn_equal(dist[:, :7], idx[:, :7], ref_d, ref_i)
    assert np.isinf(dist[:, 7:]).all() and (idx[:, 7:] == -1).all()


def test_many_exact_ties(rs):
    import multimodal_similarity_b200 as mm
    base = rs.randn(40, 128).astype(np.float32)
    g = np.repeat(base, 30, axis=0)                               # every row 30 times
    q = base[:10] + 0.01 * rs.randn(10, 128).astype(np.float32)
    dist, idx = mm.retrieve(q, g, 50)
    ref_d, ref_i = O.knn(q, g, 50)
    assert_knn_equal(dist, idx, ref_d, ref_i)
    assert np.array_equal(idx, ref_i)                             # ties broken by index on both sides


def test_golden_retrieve_one():
    """First k of the reference's retrieve_one (golden vectors from the unmodified reference)."""
    import multimodal_similarity_b200 as mm
    for name in ("small", "fused", "odd"):
        g = golden(f"retrieve_{name}.npz")
        x = g["x"]
        for n, qi in enumerate(g["queries"]):
            db = np.delete(x, qi, 0)
            dist, idx = mm.retrieve(x[qi:qi + 1], db, 50)
            ref_sorted = g["dist"][n][g["order"][n]][:50]
            assert np.array_equal(dist[0], ref_sorted)
            untied = np.r_[True, ref_sorted[1:] != ref_sorted[:-1]] & np.r_[ref_sorted[:-1] != ref_sorted[1:], False]
            assert np.array_equal(idx[0][untied], g["order"][n][:50][untied])


def test_late_fusion_equals_concat(rs):
    import multimodal_similarity_b200 as mm
    cam, _ = clustered(rs, 3000, 128, 7)
    sen, _ = clustered(rs, 3000, 128, 7)
    d1, i1 = mm.retrieve(cam[:100], cam[100:], 50, queries2=sen[:100], gallery2=sen[100:])
    fused = O.late_fusion(cam, sen)
    ref_d, ref_i = O.knn(fused[:100], fused[100:], 50)
    assert_knn_equal(d1, i1, ref_d, ref_i)


def test_cuda_tensors_and_status(rs):
    import multimodal_similarity_b200 as mm
    from multimodal_similarity_b200.retrieval import knn_raw
    g = torch.from_numpy(clustered(rs, 5000, 128, 9)[0]).cuda()
    q = g[:256].clone()
    d, i, status = knn_raw(q, g, 100)
    s = status.cpu().numpy()
    print("uncertified queries:", s[0])
    assert s[1] == 0 and s[2] == 0
    d2, i2 = mm.retrieve(q, g, 100)
    assert d2.is_cuda and torch.equal(d2, d) and torch.equal(i2, i.long())
    assert torch.all(i[:, 0] == torch.arange(256, device="cuda", dtype=torch.int32)) and torch.all(d[:, 0] == 0)


def test_certificate_is_not_vacuous(rs):
    """On ordinary data nearly every query is certified by the tcgen05 filter; the fallback stays (almost) idle."""
    from multimodal_similarity_b200.retrieval import knn_raw
    g = torch.from_numpy(clustered(rs, 30000, 128, 100)[0]).cuda()
    q = torch.from_numpy(clustered(rs, 1024, 128, 100)[0]).cuda()
    _, _, status = knn_raw(q, g, 100)
    assert int(status[0]) <= 10, f"{int(status[0])} of 1024 queries fell back to the exact path"


def test_rejects_bad_arguments(rs):
    import multimodal_similarity_b200 as mm
    x = rs.randn(10, 128).astype(np.float32)
    with pytest.raises(ValueError):
        mm.retrieve(x, x, 113)
    with pytest.raises(ValueError):
        mm.retrieve(x, rs.randn(10, 64).astype(np.float32), 3)
    with pytest.raises(ValueError):
        mm.retrieve(rs.randn(4, 300).astype(np.float32), rs.randn(9, 200).astype(np.float32), 3)   # wide, and mismatched
    from multimodal_similarity_b200.retrieval import knn_raw
    with pytest.raises(mm.MmsimError):                                                           # the tensor-core call itself: D <= 256
        knn_raw(torch.from_numpy(rs.randn(4, 300).astype(np.float32)).cuda(), torch.from_numpy(rs.randn(9, 300).astype(np.float32)).cuda(), 3)


def test_random_shape_sweep(rs):
    import multimodal_similarity_b200 as mm
    for trial in range(24):
        d = int(rs.choice([32, 64, 100, 128, 160, 192, 200, 256]))
        nq = int(rs.randint(1, 700))
        ng = int(rs.randint(1, 9000))
        k = int(rs.randint(1, 113))
        g, _ = clustered(rs, ng, d, int(rs.randint(1, 30)))
        q, _ = clustered(rs, nq, d, 5)
        dist, idx = mm.retrieve(q, g, k)
        kk = min(k, ng)
        ref_d, ref_i = O.knn(q, g, kk)
        assert_knn_equal(dist[:, :kk], idx[:, :kk], ref_d, ref_i)


@pytest.mark.parametrize("normalize", [True, False])
def test_unclustered_data_certifies(normalize):
    """Isotropic Gaussian data: distances concentrate, so the tcgen05 filter has the least slack.  The certificate must
    still pass for (nearly) every query, and the results must be exact either way."""
    from multimodal_similarity_b200.retrieval import knn_raw, check_status
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    g = torch.randn(200_000, 128, generator=gen, device="cuda")
    q = torch.randn(4096, 128, generator=gen, device="cuda")
    if normalize:
        g = g / g.norm(dim=1, keepdim=True)
        q = q / q.norm(dim=1, keepdim=True)
    g, q = g.contiguous(), q.contiguous()
    d, i, status = knn_raw(q, g, 100)
    fb = check_status(status)
    print("normalize", normalize, "exact-fallback queries:", fb)
    assert fb <= 40
    rows = [0, 1, 77, 4095]
    ref_d, ref_i = O.knn(q[rows].cpu().numpy(), g.cpu().numpy(), 100)
    assert_knn_equal(d[rows].cpu().numpy(), i[rows].cpu().numpy().astype(np.int64), ref_d, ref_i)


def test_cta_pair_sweep_matches(rs, monkeypatch):
    """The sweep as CTA pairs (MMSIM_KNN_PAIR=1: clusters of two CTAs, tcgen05 cta_group::2, each CTA's TMA loading half of
    every gallery tile) returns exactly what the single-CTA sweep returns -- odd and even numbers of query blocks."""
    import multimodal_similarity_b200 as mm
    x, _ = clustered(rs, 70000, 128, 40)
    g = torch.from_numpy(x).cuda()
    for nq in (300, 1000):                                   # 3 and 8 query blocks
        q = torch.from_numpy(clustered(rs, nq, 128, 40)[0]).cuda()
        monkeypatch.delenv("MMSIM_KNN_PAIR", raising=False)
        d0, i0 = mm.retrieve(q, g, 50)
        monkeypatch.setenv("MMSIM_KNN_PAIR", "1")
        d1, i1 = mm.retrieve(q, g, 50)
        assert torch.equal(d0, d1) and torch.equal(i0, i1)
    monkeypatch.delenv("MMSIM_KNN_PAIR", raising=False)


def test_query_streaming_sweep_matches(rs, monkeypatch):
    """The opt-in query-streaming sweep (MMSIM_KNN_SWEEP=q, csrc/knn_sweepq.cuh: resident gallery tile, streamed query
    blocks, global thresholds, candidates drained through a shared-memory ring) returns exactly what the default sweep
    returns -- several gallery tiles per CTA, query chunks, 64- / 128- / 256-d."""
    import multimodal_similarity_b200 as mm
    for ng, nq, d, k in ((70000, 1000, 128, 50), (30000, 4500, 64, 30), (20000, 300, 256, 100), (1500, 40, 96, 10)):
        g = torch.from_numpy(clustered(rs, ng, d, 40)[0]).cuda()
        q = torch.from_numpy(clustered(rs, nq, d, 40)[0]).cuda()
        monkeypatch.delenv("MMSIM_KNN_SWEEP", raising=False)
        d0, i0 = mm.retrieve(q, g, k)
        monkeypatch.setenv("MMSIM_KNN_SWEEP", "q")
        d1, i1 = mm.retrieve(q, g, k)
        assert torch.equal(d0, d1) and torch.equal(i0, i1)
    monkeypatch.delenv("MMSIM_KNN_SWEEP", raising=False)


def test_grouped_queries_exclude_self_vs_oracle(rs):
    """Enough queries and gallery rows for query grouping (sweep order != caller's order): leave-one-out retrieval with
    the queries being gallery rows, spot-checked against the oracle; and grouping on/off give identical results."""
    import multimodal_similarity_b200 as mm
    x, _ = clustered(rs, 70000, 64, 300)
    g = torch.from_numpy(x).cuda()
    nq, off, k = 4500, 1234, 30
    q = g[off:off + nq].clone()
    d, i = mm.retrieve(q, g, k, exclude_self=True, self_offset=off)
    d, i = d.cpu().numpy(), i.cpu().numpy()
    for qi in (0, 1, 777, 2048, nq - 1):
        rd, ri = O.knn(x[off + qi:off + qi + 1], x, k, exclude_self=True, self_offset=off + qi)
        assert np.array_equal(d[qi], rd[0]) and np.array_equal(i[qi], ri[0]), qi
    assert not (i == (off + np.arange(nq))[:, None]).any()
    os.environ["MMSIM_KNN_GROUP"] = "0"
    try:
        d0, i0 = mm.retrieve(q, g, k, exclude_self=True, self_offset=off)
    finally:
        del os.environ["MMSIM_KNN_GROUP"]
    assert np.array_equal(d0.cpu().numpy(), d) and np.array_equal(i0.cpu().numpy(), i)


@pytest.mark.parametrize("n,d,k,excl", [(3000, 384, 10, False), (2500, 64, 500, True), (900, 1024, 899, True), (700, 32, 700, False)])
def test_shapes_beyond_the_tensor_core_pipeline(n, d, k, excl, rs):
    """k > 112 (up to the reference's full ranking) and D > 256 take the exact per-query path: same distances, same order."""
    import multimodal_similarity_b200 as mm
    x, _ = clustered(rs, n, d, 9)
    q = x[:40] if excl else clustered(rs, 40, d, 9)[0]
    dist, idx = mm.retrieve(q, x, k, exclude_self=excl)
    ref_d, ref_i = O.knn(q, x, k, exclude_self=excl)
    assert dist.dtype == np.float32 and idx.dtype == np.int64
    assert np.array_equal(dist, ref_d) and np.array_equal(idx, ref_i)


def test_dual_query_block_sweep_matches(rs):
    """MMSIM_KNN_DUAL=1 (opt-in): two query blocks resident per CTA, every gallery tile multiplied with both -- same logs
    per query row, so bit-identical results (odd number of query blocks: the last CTA's second block is empty)."""
    from multimodal_similarity_b200.retrieval import knn_raw, check_status
    x, _ = clustered(rs, 70000, 128, 60)
    g = torch.from_numpy(x).cuda()
    q = torch.from_numpy(clustered(rs, 5 * 128 + 37, 128, 60)[0]).cuda()
    d0, i0, st0 = knn_raw(q, g, 100)
    check_status(st0)
    d0, i0 = d0.clone(), i0.clone()
    os.environ["MMSIM_KNN_DUAL"] = "1"
    try:
        d1, i1, st1 = knn_raw(q, g, 100)
        check_status(st1)
    finally:
        del os.environ["MMSIM_KNN_DUAL"]
    assert torch.equal(d0, d1) and torch.equal(i0, i1)
    ref_d, ref_i = O.knn(q[:64].cpu().numpy(), x, 100)
    assert np.array_equal(d1[:64].cpu().numpy(), ref_d)
