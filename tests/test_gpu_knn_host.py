"""Host-buffer retrieval (mmsim_knn_host_f32 through knn_host / retrieve_host): the gallery arrives split by split on a
copy stream while the previous split is swept.  Results must equal the device-resident call bit for bit, and the oracle."""
import numpy as np
import pytest
import torch

from conftest import clustered
from oracle import retrieval_np as O
from test_gpu_knn import assert_knn_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm():
    import multimodal_similarity_b200 as mm
    return mm


def device_result(q, g, k, **kw):
    from multimodal_similarity_b200.retrieval import check_status, knn_raw
    d, i, st = knn_raw(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), k, **kw)
    check_status(st)
    return d.cpu(), i.cpu()


@pytest.mark.parametrize("splits", [None, 3, 8])
@pytest.mark.parametrize("nq,ng,d,k,pinned", [
    (300, 20000, 128, 100, True),
    (1000, 33333, 64, 10, True),       # ragged last tile, last split shorter than the others
    (129, 20000, 256, 50, False),      # pageable memory: the copies serialise, the result may not change
    (200, 9000, 30, 7, True),          # width that is not a multiple of 4: scalar operand copies
    (64, 700, 128, 16, True),          # small gallery: logged whole, no pivot pre-pass, no sample
])
def test_host_call_equals_device_call(mm, rs, monkeypatch, splits, nq, ng, d, k, pinned):
    from multimodal_similarity_b200.retrieval import check_status, knn_host
    if splits is not None:
        monkeypatch.setenv("MMSIM_KNN_SPLITS", str(splits))
    g, _ = clustered(rs, ng, d, 11)
    q, _ = clustered(rs, nq, d, 11)
    q[:3] = g[:3]
    qh, gh = torch.from_numpy(q), torch.from_numpy(g)
    if pinned:
        qh, gh = qh.pin_memory(), gh.pin_memory()
    dist, idx, st = knn_host(qh, gh, k)
    check_status(st)
    ref_d, ref_i = device_result(q, g, k)
    assert torch.equal(dist.cpu(), ref_d) and torch.equal(idx.cpu(), ref_i)
    if splits is None:
        o_d, o_i = O.knn(q, g, k)
        assert_knn_equal(dist.cpu().numpy(), idx.cpu().numpy().astype(np.int64), o_d, o_i)


def test_back_to_back_calls_share_staging(mm, rs, monkeypatch):
    """The second call's copies must wait for the first call's kernels (same staging buffers, same workspace), and
    results written into page-locked host memory are complete once the stream is."""
    from multimodal_similarity_b200.retrieval import knn_host
    monkeypatch.setenv("MMSIM_KNN_SPLITS", "4")
    nq, ng, d, k = 500, 30000, 128, 20
    dev = torch.device("cuda", 0)
    stage = (torch.empty((nq, d), device=dev), torch.empty((ng, d), device=dev))
    data, outs = [], []
    for rep in range(3):
        g, _ = clustered(rs, ng, d, 5 + rep)
        q, _ = clustered(rs, nq, d, 5 + rep)
        data.append((q, g, torch.from_numpy(q).pin_memory(), torch.from_numpy(g).pin_memory()))
        outs.append((torch.empty((nq, k), dtype=torch.float32).pin_memory(), torch.empty((nq, k), dtype=torch.int32).pin_memory(),
                     torch.empty(8, dtype=torch.int32, device=dev)))
    for (q, g, qh, gh), out in zip(data, outs):          # no synchronisation in between
        knn_host(qh, gh, k, stage=stage, out=out)
    torch.cuda.synchronize()
    for (q, g, _, _), (d_, i_, st) in zip(data, outs):
        assert st.tolist()[1] == 0 and st.tolist()[2] == 0
        ref_d, ref_i = device_result(q, g, k)
        assert torch.equal(d_, ref_d) and torch.equal(i_, ref_i)


def test_retrieve_host_numpy_in_numpy_out(mm, rs):
    x, _ = clustered(rs, 3000, 128, 9)
    dist, idx = mm.retrieve_host(x, x, 10, exclude_self=True)
    ref_d, ref_i = O.knn(x, x, 10, exclude_self=True)
    assert_knn_equal(dist, idx, ref_d, ref_i)
    d2, i2 = mm.retrieve(x, x, 10, exclude_self=True)
    assert np.array_equal(dist, d2) and np.array_equal(idx, i2)


def test_host_call_on_a_side_stream(mm, rs, monkeypatch):
    from multimodal_similarity_b200.retrieval import check_status, knn_host
    monkeypatch.setenv("MMSIM_KNN_SPLITS", "2")
    g, _ = clustered(rs, 12000, 128, 4)
    q, _ = clustered(rs, 256, 128, 4)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        dist, idx, st = knn_host(torch.from_numpy(q).pin_memory(), torch.from_numpy(g).pin_memory(), 5)
        check_status(st)
    side.synchronize()
    ref_d, ref_i = device_result(q, g, 5)
    assert torch.equal(dist.cpu(), ref_d) and torch.equal(idx.cpu(), ref_i)


def test_errors(mm):
    from multimodal_similarity_b200.retrieval import knn_host
    q = torch.zeros((4, 8))
    with pytest.raises(ValueError, match="CPU tensor"):
        knn_host(q.cuda(), q, 1)
    with pytest.raises(ValueError, match="-d but the gallery"):
        knn_host(q, torch.zeros((4, 9)), 1)
    with pytest.raises(ValueError, match="stage"):
        knn_host(q, q, 1, stage=(torch.zeros((4, 8), device="cuda"), torch.zeros((5, 8), device="cuda")))
    with pytest.raises(ValueError, match="unsupported"):
        knn_host(q, q, 200)
