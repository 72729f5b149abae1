"""``ShardedGallery.graphed``: one retrieve captured as a CUDA graph and replayed.  On one GPU the object degenerates to a
single shard (no collectives in the graph); the 2-GPU replay -- kernels + NCCL collectives in one graph -- is checked against
the eager call by ``bench.py --gpus 2 --graph`` before it is timed (``profiles/r2j_graph_n2.txt``)."""
import numpy as np
import pytest
import torch

from conftest import clustered
from oracle import retrieval_np as O

pytestmark = pytest.mark.gpu


def test_graphed_replay_equals_eager_and_oracle(rs):
    import multimodal_similarity_b200 as mm
    x, _ = clustered(rs, 9000, 128, 20)
    y, _ = clustered(rs, 700, 128, 20)
    sg = mm.ShardedGallery(x)
    q = torch.from_numpy(y[:350]).cuda()
    eager_d, eager_i = sg.retrieve(q, 25)
    g = sg.graphed(q, 25)
    try:
        d, i = g()
        torch.cuda.synchronize()
        assert torch.equal(d, eager_d) and torch.equal(i, eager_i)
        assert int(g.uncertified) == 0
        # new queries go into the captured buffer; the replay serves them
        d2, i2 = g(torch.from_numpy(y[350:]).cuda())
        torch.cuda.synchronize()
        ref_d, ref_i = O.knn(y[350:], x, 25)
        assert np.array_equal(d2.cpu().numpy(), ref_d) and np.array_equal(i2.cpu().numpy(), ref_i)
    finally:
        g.close()


def test_graphed_needs_device_queries(rs):
    import multimodal_similarity_b200 as mm
    x, _ = clustered(rs, 500, 64, 5)
    sg = mm.ShardedGallery(x)
    with pytest.raises(ValueError):
        sg.graphed(x[:10], 5)


def test_rerank_gather_forms_agree(rs, monkeypatch):
    """The re-rank kernel's opt-in 128-bit gather form (MMSIM_RR_VEC=1: two lanes per candidate, packed 64-bit sort; 8 | D <= 256)
    and its default 8-lane form return the same bits -- one- and two-leaf summation trees, ties (duplicated gallery rows),
    leave-one-out -- and both equal the oracle."""
    import multimodal_similarity_b200 as mm
    for ng, nq, d, k, excl in ((40000, 700, 128, 100, False), (30000, 500, 256, 50, False), (20000, 400, 200, 30, False),
                               (9000, 600, 64, 20, True), (5000, 200, 8, 10, False)):
        x = clustered(rs, ng, d, 30)[0]
        x[1::7] = x[0:-1:7][: len(x[1::7])]                        # exact duplicates: ties on (distance), broken by index
        g = torch.from_numpy(x).cuda()
        q = g[:nq].clone() if excl else torch.from_numpy(clustered(rs, nq, d, 30)[0]).cuda()
        monkeypatch.setenv("MMSIM_RR_VEC", "1")
        d1, i1 = mm.retrieve(q, g, k, exclude_self=excl)
        monkeypatch.delenv("MMSIM_RR_VEC", raising=False)
        d0, i0 = mm.retrieve(q, g, k, exclude_self=excl)
        assert torch.equal(d0, d1) and torch.equal(i0, i1), (ng, nq, d, k)
        ref_d, _ = O.knn(q[:64].cpu().numpy(), x, k, exclude_self=excl)
        assert np.array_equal(d1[:64].cpu().numpy(), ref_d)
