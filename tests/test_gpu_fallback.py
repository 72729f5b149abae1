"""The exact fallback of the kNN retrieval (csrc/knn_fallback.cuh): whatever the row order of the gallery and however
many exact ties it holds, ``retrieve`` returns the reference's ``argsort(norm(q - G))[:k]`` (src/utils.py:73-74).
Covers the round-1 advisor finding (class-sorted galleries), both tiers of the fallback through test hooks, and the
continuation protocol (status[1] > status[2] -> mmsim_knn_finish_f32)."""
import numpy as np
import pytest
import torch

from conftest import clustered
from oracle import retrieval_np as O
from test_gpu_knn import assert_knn_equal

pytestmark = pytest.mark.gpu


def _knn(q, g, k, **kw):
    from multimodal_similarity_b200.retrieval import check_status, knn_raw
    d, i, st = knn_raw(q, g, k, **kw)
    n_fb = check_status(st)
    s = st.tolist()
    assert s[1] == s[2], s
    return d, i, n_fb, s


def test_class_sorted_gallery_leave_one_out(rs):
    """CUB / SOP style gallery: rows sorted by class, classes of 120 rows, leave-one-out, k = 100.  With a contiguous pivot
    sample a query whose class was sampled got a threshold near its 12th neighbour and fewer than k candidates (round-1
    advisor finding); the sample is now strided, and whatever still fails the certificate is recomputed exactly."""
    n_cls, per, d, k = 500, 120, 128, 100
    cent = rs.randn(n_cls, d).astype(np.float32)
    lab = np.repeat(np.arange(n_cls), per)
    x = cent[lab] + 0.35 * rs.randn(lab.size, d).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x = x.astype(np.float32)
    g = torch.from_numpy(x).cuda()
    nq = 6000
    dd, ii, n_fb, s = _knn(g[:nq].contiguous(), g, k, exclude_self=True)
    print("class-sorted gallery: exact-fallback queries", n_fb, "of", nq, "status", s)
    assert n_fb <= nq // 20                         # the strided sample keeps the fallback rare
    dd, ii = dd.cpu().numpy(), ii.cpu().numpy().astype(np.int64)
    rows = np.r_[0:8, 119, 120, 121, 2999:3003, nq - 3:nq]
    # O.knn excludes gallery row (self_offset + position); the rows are not contiguous, so query one by one
    for n_, r in enumerate(rows):
        rd, ri = O.knn(x[r:r + 1], x, k, exclude_self=True, self_offset=int(r))
        assert_knn_equal(dd[r:r + 1], ii[r:r + 1], rd, ri)
    # every query found its own class first: the 100 nearest rows of a tight class of 120 are all of that class
    same = (lab[ii] == lab[:nq, None]).mean()
    assert same > 0.99, same


def test_periodic_class_order(rs):
    """Rows interleaved with period 32 (class = row % 32): a fixed stride of 64 would sample one class only."""
    n, d, k = 40000, 64, 50
    lab = np.arange(n) % 32
    cent = rs.randn(32, d).astype(np.float32)
    x = cent[lab] + 0.4 * rs.randn(n, d).astype(np.float32)
    x = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    g = torch.from_numpy(x).cuda()
    dd, ii, n_fb, s = _knn(g[:2000].contiguous(), g, k, exclude_self=True)
    print("periodic order: exact-fallback queries", n_fb, "status", s)
    rows = [0, 1, 31, 32, 1999]
    for r in rows:
        rd, ri = O.knn(x[r:r + 1], x, k, exclude_self=True, self_offset=r)
        assert_knn_equal(dd[r:r + 1].cpu().numpy(), ii[r:r + 1].cpu().numpy().astype(np.int64), rd, ri)


@pytest.mark.parametrize("nq,ng,d,k,mod", [(1000, 30000, 128, 100, 3), (300, 5000, 64, 7, 1), (4500, 70000, 96, 30, 7)])
def test_forced_fallback_tier1_equals_certified_result(rs, monkeypatch, nq, ng, d, k, mod):
    """MMSIM_KNN_FORCE_FALLBACK=m: every m-th query is treated as uncertified and recomputed by the tensor-core re-sweep +
    exact selection; the result must not change."""
    g = torch.from_numpy(clustered(rs, ng, d, 40)[0]).cuda()
    q = torch.from_numpy(clustered(rs, nq, d, 40)[0]).cuda()
    d0, i0, fb0, _ = _knn(q, g, k)
    d0, i0 = d0.clone(), i0.clone()
    monkeypatch.setenv("MMSIM_KNN_FORCE_FALLBACK", str(mod))
    d1, i1, fb1, s = _knn(q, g, k)
    assert fb1 >= -(-nq // mod) and s[1] == 0, (fb1, s)          # all through tier 1: nothing queued for the scan
    assert torch.equal(d0, d1) and torch.equal(i0, i1)
    rd, ri = O.knn(q[:3].cpu().numpy(), g.cpu().numpy(), k)
    assert_knn_equal(d1[:3].cpu().numpy(), i1[:3].cpu().numpy().astype(np.int64), rd, ri)


def test_forced_tier2_runs_in_waves(rs, monkeypatch):
    """MMSIM_KNN_FORCE_TIER2=1 sends every fallback query to the streaming exact scan; 140 of them need a second wave
    (mmsim_knn_finish_f32), which check_status launches."""
    from multimodal_similarity_b200.retrieval import check_status, knn_raw
    g = torch.from_numpy(clustered(rs, 20000, 128, 20)[0]).cuda()
    q = torch.from_numpy(clustered(rs, 700, 128, 20)[0]).cuda()
    d0, i0, _, _ = _knn(q, g, 33, exclude_self=True, self_offset=100)
    d0, i0 = d0.clone(), i0.clone()
    monkeypatch.setenv("MMSIM_KNN_FORCE_FALLBACK", "5")
    monkeypatch.setenv("MMSIM_KNN_FORCE_TIER2", "1")
    d1, i1, st = knn_raw(q, g, 33, True, 100)
    s = st.tolist()
    assert s[0] >= 140 and s[1] >= 140 and s[2] == 128, s        # one wave done by the call itself
    check_status(st)
    s = st.tolist()
    assert s[1] == s[2], s
    assert torch.equal(d0, d1) and torch.equal(i0, i1)


def test_massive_duplicates_overflow_to_the_scan(rs):
    """6,000 copies of one row next to the queries: more rows within the bound than a tier-1 log holds -> the streaming scan
    resolves the ties by index, like the oracle."""
    base, _ = clustered(rs, 20000, 128, 10)
    dup = np.repeat(base[:1], 6000, axis=0)
    x = np.concatenate([base[:7000], dup, base[7000:]]).astype(np.float32)
    q = (base[:1] + 0.001 * rs.randn(5, 128)).astype(np.float32)
    import multimodal_similarity_b200 as mm
    dist, idx = mm.retrieve(q, x, 100)
    ref_d, ref_i = O.knn(q, x, 100)
    assert_knn_equal(dist, idx, ref_d, ref_i)
    assert np.array_equal(idx, ref_i)


def test_host_call_with_forced_fallback(rs, monkeypatch):
    """The host-buffer call shares the fallback (device staging copies of Q and G are what it re-reads)."""
    from multimodal_similarity_b200.retrieval import check_status, knn_host
    g, _ = clustered(rs, 30000, 128, 11)
    q, _ = clustered(rs, 600, 128, 11)
    monkeypatch.setenv("MMSIM_KNN_FORCE_FALLBACK", "4")
    dist, idx, st = knn_host(torch.from_numpy(q).pin_memory(), torch.from_numpy(g).pin_memory(), 20)
    assert check_status(st) >= 150
    ref_d, ref_i = O.knn(q[:40], g, 20)
    assert_knn_equal(dist[:40].cpu().numpy(), idx[:40].cpu().numpy().astype(np.int64), ref_d, ref_i)
    monkeypatch.delenv("MMSIM_KNN_FORCE_FALLBACK")
    d2, i2, st2 = knn_host(torch.from_numpy(q).pin_memory(), torch.from_numpy(g).pin_memory(), 20)
    check_status(st2)
    assert torch.equal(dist.cpu(), d2.cpu()) and torch.equal(idx.cpu(), i2.cpu())


def test_retrieve_one_has_the_reference_return_value():
    """retrieve_one returns the reference's triple (full unsorted dist, full argsort, sklearn AP) -- golden vectors from the
    unmodified reference (round-1 advisor finding: it used to return a truncated, sorted top-k)."""
    import multimodal_similarity_b200 as mm
    from conftest import golden
    for name in ("small", "fused", "odd"):
        gv = golden(f"retrieve_{name}.npz")
        x, lab = gv["x"], gv["labels"]
        for n, qi in enumerate(gv["queries"]):
            db = np.delete(x, qi, 0)
            gl = np.delete(lab, qi, 0)
            ql = lab[qi] if lab[qi] > 0 else 1                                # as oracle/make_golden.py called the reference
            dist, idx, ap = mm.retrieve_one(x[qi], db, ql, gl)
            assert dist.dtype == np.float32 and dist.shape == (db.shape[0],) and np.array_equal(dist, gv["dist"][n])
            ref_order = gv["order"][n]
            assert np.array_equal(dist[idx], dist[ref_order])                 # same ranking up to exact ties
            assert abs(ap - gv["ap"][n]) < 1e-12
