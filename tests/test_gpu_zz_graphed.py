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
