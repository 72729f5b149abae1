"""Device semi-hard mining (csrc/mining.cu through mining.select_triplets_facenet) against the reference's outputs
(tests/golden/mining_*.npz, made by the unmodified reference) and against the oracle on random inputs."""
import glob
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, clustered

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN, "mining_*.npz")))


@pytest.fixture(scope="module")
def mm():
    import multimodal_similarity_b200 as mm
    return mm


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("on_device", [False, True])
def test_golden_triplets_bit_exact(mm, name, on_device):
    import torch
    g = np.load(os.path.join(GOLDEN, f"mining_{name}.npz"))
    x = torch.from_numpy(g["x"]).cuda() if on_device else g["x"]
    dist = mm.cdist(mm.all_diffs(x, x), metric=str(g["metric"]))      # the reference call site, src/base_model.py:271
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    miner = mm.select_triplets_facenet_cub if name.startswith("cubcopy") else mm.select_triplets_facenet
    trip, active = miner(g["labels"], dist, int(g["triplet_per_batch"]), alpha=float(g["alpha"]),
                         num_negative=int(g["num_negative"]))
    if bool(g["empty_is_none"]):                                      # src/base_model_CUB.py:91
        assert trip is None and active is None
        return
    assert np.array_equal(np.asarray(trip, dtype=np.int64), g["triplets"])
    assert float(active) == float(g["active"])
    if len(trip) == 0:
        assert trip == [] and active == 0.


@pytest.mark.parametrize("n,d,classes,alpha", [(1, 8, 1, 0.2), (33, 16, 3, 0.3), (1000, 128, 7, 0.2), (513, 96, 50, 1.0)])
def test_counts_masks_and_picks_vs_oracle(mm, n, d, classes, alpha):
    import torch
    from oracle import mining_np as M
    from oracle import retrieval_np as O
    from multimodal_similarity_b200 import _lib, mining
    rs = np.random.RandomState(n)
    x, lab = clustered(rs, n, d, classes, noise=1.0, background=0.2)
    dist = O.cdist(O.all_diffs(x, x))
    if n > 40:
        dist[3, 5] = dist[3, 4]                 # an exact tie with the positive distance: strict '<' must exclude it
    pairs = np.stack([rs.randint(0, n, 400), rs.randint(0, n, 400)], axis=1).astype(np.int32)
    count, mask = mining.semihard_counts(dist, lab, pairs, alpha, return_mask=True)
    count, mask = count.cpu().numpy(), mask.cpu().numpy().view(np.uint32)
    bits = np.unpackbits(mask.view(np.uint8), axis=1, bitorder="little")[:, :n]
    picks = []
    for i, (a, p) in enumerate(pairs):
        want = M.semihard_set(dist, lab.astype(np.int64), a, p, np.float32(alpha))
        assert count[i] == len(want)
        assert np.array_equal(np.flatnonzero(bits[i]), want)
        for r in {0, len(want) // 2, len(want) - 1, len(want)}:
            picks.append((a, p, r, want[r] if 0 <= r < len(want) else -1))
    picks = np.asarray(picks, dtype=np.int64)
    dev = torch.device("cuda")
    dist_d = torch.from_numpy(dist).to(dev)
    lab_d = torch.from_numpy(lab.astype(np.int32)).to(dev)
    pk = torch.from_numpy(picks[:, :3].astype(np.int32)).contiguous().to(dev)
    out = torch.empty(len(picks), dtype=torch.int32, device=dev)
    rc = _lib.load().mmsim_semihard_pick_f32(dist_d.data_ptr(), n, n, lab_d.data_ptr(), pk.data_ptr(), len(picks), float(alpha),
                                             out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "pick")
    assert np.array_equal(out.cpu().numpy(), picks[:, 3])


def test_random_seeds_vs_oracle(mm):
    from oracle import mining_np as M
    rs = np.random.RandomState(99)
    x, lab = clustered(rs, 700, 128, 9, noise=0.9, background=0.3)
    lab_f = lab.astype(np.float32)                                   # float labels, as the trainers feed them
    dist = mm.pairwise_distance(x, x)
    for seed, tpb, nneg in ((1, 50, 3), (2, 1000, 1), (3, 7, 5)):
        random.seed(seed); np.random.seed(seed)
        want = M.select_triplets_facenet(lab_f, dist, tpb, 0.2, nneg)
        state_py, state_np = random.getstate(), np.random.get_state()[1].copy()
        random.seed(seed); np.random.seed(seed)
        got = mm.select_triplets_facenet(lab_f, dist, tpb, 0.2, nneg)
        assert [int(v) for v in got[0]] == [int(v) for v in want[0]] and got[1] == float(want[1])
        # both global RNGs are left where the reference leaves them
        assert random.getstate() == state_py and np.array_equal(np.random.get_state()[1], state_np)


def test_errors(mm):
    with pytest.raises(ValueError):
        mm.select_triplets_facenet(np.zeros(4), np.zeros((4, 5), np.float32), 3)
    assert mm.select_triplets_facenet(np.ones(4), np.zeros((4, 4), np.float32), 0) == ([], 0.)
