"""K6: shard-simulation on one GPU -- split the gallery into R shards, local top-k on each, same merge kernel,
compare with the unsharded result bit for bit."""
import numpy as np
import pytest
import torch

from conftest import clustered
from oracle import retrieval_np as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R", [2, 3, 8])
@pytest.mark.parametrize("exclude_self", [False, True])
def test_sharded_equals_unsharded(R, exclude_self, rs):
    import multimodal_similarity_b200 as mm
    from multimodal_similarity_b200.retrieval import knn_raw
    from multimodal_similarity_b200.sharded import merge_parts, shard_bounds
    x, _ = clustered(rs, 6001, 128, 11)
    g = torch.from_numpy(x).cuda()
    q = g[:300].clone() if exclude_self else torch.from_numpy(clustered(rs, 300, 128, 11)[0]).cuda()
    k = 37
    full_d, full_i = mm.retrieve(q, g, k, exclude_self=exclude_self)
    packed = torch.empty((R, 2, 300, k), dtype=torch.int32, device="cuda")
    bases = []
    for r in range(R):
        lo, hi = shard_bounds(6001, R, r)
        d, i, st = knn_raw(q, g[lo:hi].contiguous(), k, exclude_self, -lo)
        assert int(st[1]) == 0 and int(st[2]) == 0
        packed[r, 0] = d.view(torch.int32)
        packed[r, 1] = i
        bases.append(lo)
    md, mi = merge_parts(packed[:, 0].view(torch.float32), packed[:, 1], torch.tensor(bases, device="cuda"), k)
    assert torch.equal(md, full_d) and torch.equal(mi, full_i)
    ref_d, ref_i = O.knn(q.cpu().numpy(), x, k, exclude_self=exclude_self)
    assert np.array_equal(md.cpu().numpy(), ref_d)


def test_sharded_gallery_single_process(rs):
    import multimodal_similarity_b200 as mm
    x, _ = clustered(rs, 2000, 128, 5)
    sg = mm.ShardedGallery(x)
    d, i = sg.retrieve(x[:50], 10)
    ref_d, ref_i = O.knn(x[:50], x, 10)
    assert np.array_equal(d, ref_d) and i.dtype == np.int64


def _simulate_reduced(g, q, k, R, exclude_self, repair=False):
    """The reduced sharded protocol with R simulated ranks in one process (collectives replaced by torch.stack).
    repair=True adds the per-query exact fallback of the uncertified queries and returns the repaired result."""
    from multimodal_similarity_b200.sharded import (FALLBACK_CAP, ReducedShard, merge_certified, merge_pivots_into, patch_rows,
                                                    reduced_kp, shard_bounds)
    n, nq = g.shape[0], q.shape[0]
    kp = reduced_kp(R, k)
    shards, packed, pivs, bases = [], [], [], []
    for r in range(R):
        lo, hi = shard_bounds(n, R, r)
        shards.append(ReducedShard(g[lo:hi].contiguous(), lo))
        packed.append(torch.empty(ReducedShard.packed_elems(nq, kp), dtype=torch.int32, device="cuda"))
        bases.append(lo)
    for r in range(R):
        pivs.append(shards[r].stage1(q, k, kp, packed[r]))
    allpiv = torch.stack(pivs)                       # the all-gather
    for r in range(R):
        merge_pivots_into(allpiv, pivs[r])
    for r in range(R):
        shards[r].stage2(q, k, kp, exclude_self, 0, packed[r])
    gathered = torch.stack(packed)                   # the second all-gather
    whole = merge_certified(gathered, torch.tensor(bases, device="cuda"), nq, kp, k)
    # the production path merges by query slice (all-to-all + all-gather): simulate it and require identical results
    from multimodal_similarity_b200.sharded import merge_certified_slice, pack_slices, slice_rows, unpack_merged
    S = slice_rows(nq, R)
    sends = [pack_slices(packed[r], nq, kp, R) for r in range(R)]
    bases_t = torch.tensor(bases, device="cuda")
    res = []
    for r in range(R):                               # rank r receives slice r of every shard
        recv = torch.stack([sends[src][r] for src in range(R)])
        res.append(merge_certified_slice(recv, bases_t, max(0, min(S, nq - r * S)), S, kp, k))
    sd, si, unc, flag = unpack_merged(torch.stack(res), nq, S, k)
    assert int(unc) == int(whole[2][0]) and torch.equal(flag, whole[3])
    assert int((flag >= 0).sum()) == int(unc)
    # certified rows agree between the whole-batch merge and the slice merge (uncertified rows too: same kernel)
    assert torch.equal(sd, whole[0]) and torch.equal(si, whole[1])
    if repair and int(unc) > 0:
        assert int(unc) <= FALLBACK_CAP
        fbs = [shards[r].fallback(q, k, exclude_self, 0, flag, FALLBACK_CAP) for r in range(R)]
        allfb = torch.stack(fbs)                     # the third all-gather
        st = allfb[:, 2 * FALLBACK_CAP * k + FALLBACK_CAP:].cpu()
        assert bool((st[:, 0] == int(unc)).all()) and bool((st[:, 1] == st[:, 2]).all()), st
        patch_rows(sd, si, allfb, bases_t, FALLBACK_CAP, k)
        return (sd, si, whole[2], flag), kp
    return whole, kp


@pytest.mark.parametrize("R", [2, 4, 8])
@pytest.mark.parametrize("exclude_self", [False, True])
def test_reduced_protocol_equals_unsharded(R, exclude_self, rs):
    import multimodal_similarity_b200 as mm
    x, _ = clustered(rs, 40000, 128, 50)
    g = torch.from_numpy(x).cuda()
    q = g[:500].clone() if exclude_self else torch.from_numpy(clustered(rs, 500, 128, 50)[0]).cuda()
    k = 100
    (md, mi, status, _), kp = _simulate_reduced(g, q, k, R, exclude_self)
    assert kp < 128
    assert int(status[0]) == 0, f"{int(status[0])} uncertified queries with randomly ordered rows"
    full_d, full_i = mm.retrieve(q, g, k, exclude_self=exclude_self)
    assert torch.equal(md, full_d) and torch.equal(mi, full_i)


def test_reduced_protocol_with_grouped_queries(rs):
    """4,200 queries: every simulated shard sorts them by nearest anchor (shard mode always groups), the pivot lists they
    exchange are indexed by sweep position, results come back in the caller's order."""
    import multimodal_similarity_b200 as mm
    x, _ = clustered(rs, 50000, 64, 200)
    g = torch.from_numpy(x).cuda()
    q = torch.from_numpy(clustered(rs, 4200, 64, 200)[0]).cuda()
    (md, mi, status, _), kp = _simulate_reduced(g, q, 20, 4, False, repair=True)
    # Round 1 failed here on a fresh box with ONE uncertified query: the anchor assignment read an uninitialised part of the
    # workspace, so two shards could sort the queries differently and exchange misaligned pivot lists (fixed in
    # csrc/knn_tc.cu, "pack_min_kernel ... apack").  The certified count is a performance property; the RESULT is what must
    # hold, and it does either way now that uncertified queries are repaired one by one.
    assert int(status[0]) <= 2, f"{int(status[0])} uncertified queries with randomly ordered rows"
    full_d, full_i = mm.retrieve(q, g, 20)
    assert torch.equal(md, full_d) and torch.equal(mi, full_i)


def test_reduced_protocol_certificate_catches_adversarial_order(rs):
    """Gallery sorted by cluster: all neighbours of a query live in ONE shard, which re-ranks only kp < k of them.
    The global certificate must flag those queries (the caller then falls back to the exact-shards protocol)."""
    cent = rs.randn(8, 128).astype(np.float32) * 3
    lab = np.repeat(np.arange(8), 1500)
    x = cent[lab] + 0.3 * rs.randn(lab.size, 128).astype(np.float32)
    g = torch.from_numpy(x).cuda()
    q = torch.from_numpy(cent[[0, 3, 7]] + 0.3 * rs.randn(3, 128).astype(np.float32)).cuda()
    (md, mi, status, flag), kp = _simulate_reduced(g, q, 100, 8, False)
    assert kp < 100 and int(status[0]) == 3 and bool((flag >= 0).all())


@pytest.mark.parametrize("R,exclude_self", [(8, False), (4, True)])
def test_reduced_protocol_repairs_uncertified_queries_one_by_one(R, exclude_self, rs):
    """Same adversarial order, 900 queries of which only those near a shard-spanning cluster are certified: the per-query
    fallback (exact top-k inside every shard for the flagged queries + merge into those rows) restores the exact result."""
    import multimodal_similarity_b200 as mm
    cent = rs.randn(8, 128).astype(np.float32) * 3
    lab = np.repeat(np.arange(8), 1500)
    x = cent[lab] + 0.3 * rs.randn(lab.size, 128).astype(np.float32)
    g = torch.from_numpy(x).cuda()
    if exclude_self:
        q = g[:900].clone()
    else:
        q = torch.from_numpy(cent[rs.randint(0, 8, 900)] + 0.3 * rs.randn(900, 128).astype(np.float32)).cuda()
    (md, mi, status, flag), kp = _simulate_reduced(g, q, 100, R, exclude_self, repair=True)
    assert int(status[0]) > 0
    full_d, full_i = mm.retrieve(q, g, 100, exclude_self=exclude_self)
    assert torch.equal(md, full_d) and torch.equal(mi, full_i)


def test_shards_derive_the_same_sweep_order(rs):
    """Every shard sorts the queries by nearest anchor in its OWN workspace; the pivot lists they exchange are indexed by
    sweep position, so the permutations must be identical -- also when the workspaces start out with different garbage."""
    import ctypes
    from multimodal_similarity_b200 import _lib
    from multimodal_similarity_b200.sharded import ReducedShard, reduced_kp
    x, _ = clustered(rs, 30000, 64, 50)
    g = torch.from_numpy(x).cuda()
    q = torch.from_numpy(clustered(rs, 4200, 64, 50)[0]).cuda()
    kp = reduced_kp(2, 20)
    # the permutation itself: run the same shard data through both workspaces
    a = ReducedShard(g[:15000].contiguous(), 0)
    b = ReducedShard(g[:15000].contiguous(), 0)
    n = ctypes.c_size_t()
    _lib.check(_lib.load().mmsim_knn_workspace_bytes(4200, 15000, 64, 20, ctypes.byref(n)), "ws")
    a.ws = torch.zeros(n.value, dtype=torch.uint8, device="cuda")
    b.ws = torch.full((n.value,), 0x7f, dtype=torch.uint8, device="cuda")
    pa = a.stage1(q, 20, kp, torch.empty(ReducedShard.packed_elems(4200, kp), dtype=torch.int32, device="cuda")).clone()
    pb = b.stage1(q, 20, kp, torch.empty(ReducedShard.packed_elems(4200, kp), dtype=torch.int32, device="cuda")).clone()
    assert torch.equal(pa, pb)
