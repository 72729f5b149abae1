"""The N > 1 host logic of ShardedGallery under gloo with world_size 2 and 3 on CPU: row partition, the single
all-gather of packed (distance bits, local index) lists, index globalisation and the leave-one-out offset.
The CUDA steps (_local, _merge) are replaced by the oracle here -- tests may use the oracle; the product never does."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import clustered
from oracle import retrieval_np as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, exclude_self, out_dir, query_groups=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodal_similarity_b200.sharded import ShardedGallery, shard_bounds

    class CpuShardedGallery(ShardedGallery):
        def _to_device(self, x, device):
            return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))

        def _local(self, q, k, exclude_self, self_offset):
            packed = torch.empty((2, q.shape[0], k), dtype=torch.int32)
            g = self.shard.numpy()
            kk = min(k, max(g.shape[0] - (1 if exclude_self else 0), 0))
            d = np.full((q.shape[0], k), np.inf, np.float32)
            i = np.full((q.shape[0], k), -1, np.int64)
            if g.shape[0]:
                for qi in range(q.shape[0]):
                    dd = O.l2_to_all(q[qi].numpy(), g)
                    s = self_offset - self.lo + qi
                    if exclude_self and 0 <= s < g.shape[0]:
                        dd[s] = np.inf
                    order = np.lexsort((np.arange(dd.size), dd))[:k]
                    order = order[np.isfinite(dd[order])]
                    d[qi, :order.size] = dd[order]
                    i[qi, :order.size] = order
            packed[0] = torch.from_numpy(d).view(torch.int32)
            packed[1] = torch.from_numpy(i.astype(np.int32))
            return packed, None

        def _merge(self, gathered, bases, k):
            dists = gathered[:, 0].contiguous().view(torch.float32).numpy()     # [parts, Q, k]
            idx = gathered[:, 1].numpy().astype(np.int64)
            gidx = np.where(idx >= 0, idx + bases.numpy()[:, None, None], np.iinfo(np.int64).max)
            P, Q, _ = dists.shape
            dd = dists.transpose(1, 0, 2).reshape(Q, P * k)
            gg = gidx.transpose(1, 0, 2).reshape(Q, P * k)
            out_d = np.empty((Q, k), np.float32)
            out_i = np.empty((Q, k), np.int64)
            for qi in range(Q):
                o = np.lexsort((gg[qi], dd[qi]))[:k]
                out_d[qi], out_i[qi] = dd[qi][o], np.where(np.isfinite(dd[qi][o]), gg[qi][o], -1)
            return torch.from_numpy(out_d), torch.from_numpy(out_i)

    rs = np.random.RandomState(7)
    x, _ = clustered(rs, 1001, 32, 6)
    q = x[:40] if exclude_self else clustered(rs, 40, 32, 6)[0]
    sg = CpuShardedGallery(x, query_groups=query_groups)
    parts = world // query_groups
    assert (sg.parts, sg.part, sg.qgroup) == (parts, rank % parts, rank // parts)
    assert (sg.lo, sg.hi) == shard_bounds(1001, parts, rank % parts)
    d, i = sg.retrieve(q, 15, exclude_self=exclude_self)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), d=d, i=i)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("exclude_self", [False, True])
def test_sharded_gallery_gloo(world, exclude_self, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), exclude_self, str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(7)
    x, _ = clustered(rs, 1001, 32, 6)
    q = x[:40] if exclude_self else clustered(rs, 40, 32, 6)[0]
    ref_d, ref_i = O.knn(q, x, 15, exclude_self=exclude_self)
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for o in outs:                                     # identical on every rank and equal to the unsharded result
        assert np.array_equal(o["d"], ref_d) and np.array_equal(o["i"], ref_i)


@pytest.mark.parametrize("world,groups", [(2, 2), (4, 2), (4, 4)])
@pytest.mark.parametrize("exclude_self", [False, True])
def test_query_groups_gloo(world, groups, exclude_self, tmp_path):
    """2-D decomposition: `groups` query groups of world / groups gallery parts each -- sub-group all-gather of the shard
    lists, one world all-gather of the merged query slices; identical, unsharded-equal results on every rank."""
    mp.spawn(_worker, args=(world, _free_port(), exclude_self, str(tmp_path), groups), nprocs=world, join=True)
    rs = np.random.RandomState(7)
    x, _ = clustered(rs, 1001, 32, 6)
    q = x[:40] if exclude_self else clustered(rs, 40, 32, 6)[0]
    ref_d, ref_i = O.knn(q, x, 15, exclude_self=exclude_self)
    for r in range(world):
        o = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(o["d"], ref_d) and np.array_equal(o["i"], ref_i), r


def test_query_groups_resolution():
    from multimodal_similarity_b200.sharded import resolve_query_groups
    assert [resolve_query_groups("auto", w) for w in (1, 2, 3, 4, 6, 8)] == [1, 1, 1, 2, 3, 4]
    assert resolve_query_groups(4, 8) == 4 and resolve_query_groups(1, 3) == 1
    with pytest.raises(ValueError):
        resolve_query_groups(3, 8)


def test_shard_bounds_cover_and_partition():
    from multimodal_similarity_b200.sharded import shard_bounds
    for n in (0, 1, 7, 1000, 1_000_000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(w - 1))
            assert all(lo <= hi for lo, hi in b)


def test_slice_layout_addresses():
    """Host-side arithmetic of the reduced protocol's exchange buffer (what mmsim_knn_shard_f32 writes with slice_rows > 0 and
    the all-to-all sends): query q lives in block q // S at row q % S; the blocks of all ranks tile the padded query range."""
    from multimodal_similarity_b200.sharded import reduced_kp, slice_rows
    for nq, world, k in ((10, 4, 2), (7, 2, 5), (1, 3, 1), (100000, 8, 100)):
        S, kp = slice_rows(nq, world), reduced_kp(world, k)
        assert world * S >= nq and (S - 1) * world < nq
        stride = S * (2 * kp + 1)
        seen = set()
        for q in (0, nq // 2, nq - 1):
            blk, row = divmod(q, S)
            assert blk < world
            d0, i0, lb = blk * stride + row * kp, blk * stride + S * kp + row * kp, blk * stride + 2 * S * kp + row
            assert d0 + kp <= blk * stride + S * kp < i0 + kp <= blk * stride + 2 * S * kp <= lb < (blk + 1) * stride
            seen.add((blk, row))
        assert len(seen) == len({0, nq // 2, nq - 1})
        assert 1 <= kp <= 128 and (world == 1 or kp < 128 or k > 100)


def test_presharded_rows_must_follow_shard_bounds():
    """Every rank derives the other shards' row offsets from (total_rows, world): a different split is refused."""
    from multimodal_similarity_b200.sharded import ShardedGallery

    class CpuShardedGallery(ShardedGallery):
        def _to_device(self, x, device):
            return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))

    x = np.zeros((10, 4), np.float32)
    sg = CpuShardedGallery(x, presharded=True, row_offset=0, total_rows=10)      # world 1: the whole gallery
    assert (sg.lo, sg.hi, sg.total) == (0, 10, 10)
    with pytest.raises(ValueError, match="shard_bounds"):
        CpuShardedGallery(x[:7], presharded=True, row_offset=0, total_rows=10)
    with pytest.raises(ValueError, match="row_offset"):
        CpuShardedGallery(x, presharded=True)
