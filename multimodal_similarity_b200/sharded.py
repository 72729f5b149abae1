"""Gallery-sharded retrieval over the GPUs of one box (no counterpart in the single-process reference).

One process per GPU (``torch.distributed``, NCCL over NVLink).  The gallery rows are split contiguously over the
ranks, the queries are replicated.  Every rank runs the fused kNN kernel on its shard (exact, certified local
top-k), one all-gather exchanges the ``[Q, k]`` (distance f32, shard-local index i32) lists, and every rank runs the
same merge kernel ordered by (distance, global index) -- so the output is bit-identical on every rank and equal
to the single-GPU result by construction.  The training-loss kernels are not sharded (replicas only).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from ._util import is_numpy_like, stream_handle, to_cuda_f32
from .retrieval import check_status, knn_raw


def shard_bounds(n_rows: int, world: int, rank: int):
    """Contiguous row block of ``rank``: [lo, hi) with ceil(n/world) rows per rank (the last ranks may be short/empty)."""
    per = -(-n_rows // world)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def merge_parts(dist_parts: torch.Tensor, idx_parts: torch.Tensor, idx_base: torch.Tensor, k: int):
    """[parts, Q, k] shard-local results -> global (dist [Q,k] f32, idx [Q,k] i64) on the GPU (mmsim_knn_merge)."""
    lib = _lib.load()
    parts, nq = dist_parts.shape[0], dist_parts.shape[1]
    assert dist_parts.stride(1) == k and dist_parts.stride(2) == 1 and idx_parts.stride() == dist_parts.stride()
    dev = dist_parts.device
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mmsim_knn_merge(dist_parts.data_ptr(), idx_parts.data_ptr(), dist_parts.stride(0), idx_base.data_ptr(),
                                 parts, nq, k, out_d.data_ptr(), out_i.data_ptr(), stream_handle(dev))
    _lib.check(rc, "mmsim_knn_merge")
    return out_d, out_i


class ShardedGallery:
    """``ShardedGallery(gallery, group).retrieve(queries, k)`` -> identical (dist, idx) on every rank.

    gallery      the FULL [G, D] gallery (each rank keeps only its slice) or, with ``presharded=True``, this rank's
                 rows together with ``row_offset`` / ``total_rows``
    group        a torch.distributed process group (default: the world); without an initialised process group the
                 object degenerates to a single shard
    """

    def __init__(self, gallery, group=None, *, presharded=False, row_offset=None, total_rows=None, device=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if presharded:
            if row_offset is None or total_rows is None:
                raise ValueError("presharded=True needs row_offset and total_rows")
            self.lo, self.total = int(row_offset), int(total_rows)
            shard = gallery
        else:
            self.total = int(gallery.shape[0])
            self.lo, hi = shard_bounds(self.total, self.world, self.rank)
            shard = gallery[self.lo:hi]
        self.shard = self._to_device(shard, device)
        self.hi = self.lo + int(self.shard.shape[0])
        self.dim = int(gallery.shape[1])

    # -- the three steps; tests on CPU boxes replace _to_device/_local/_merge to exercise the sharding logic under gloo
    def _to_device(self, x, device):
        return to_cuda_f32(x, device)

    def _local(self, q, k, exclude_self, self_offset):
        """Exact top-k inside this rank's shard -> packed [2, Q, k] int32 (row 0 = distance bits, row 1 = local index)."""
        packed = torch.empty((2, q.shape[0], k), dtype=torch.int32, device=q.device)
        if self.shard.shape[0] == 0:
            packed[0] = torch.tensor(float("inf"), device=q.device).view(torch.int32)
            packed[1] = -1
            return packed, None
        d, i, status = knn_raw(q, self.shard, k, exclude_self, self_offset - self.lo)
        packed[0].copy_(d.view(torch.int32))
        packed[1].copy_(i)
        return packed, status

    def _merge(self, gathered, bases, k):
        return merge_parts(gathered[:, 0].view(torch.float32), gathered[:, 1], bases, k)

    def retrieve(self, queries, k, *, exclude_self=False, self_offset=0, check=True):
        as_numpy = is_numpy_like(queries)
        q = self._to_device(queries, self.shard.device)
        if q.shape[1] != self.dim:
            raise ValueError(f"queries are {q.shape[1]}-d but the gallery is {self.dim}-d")
        k = int(k)
        packed, status = self._local(q, k, exclude_self, self_offset)
        if self.world > 1:
            # dim-0 concatenation layout (accepted by both NCCL and gloo), viewed as [world, 2, Q, k] afterwards
            flat = torch.empty((self.world * 2,) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
            dist.all_gather_into_tensor(flat, packed, group=self.group)
            gathered = flat.view((self.world,) + tuple(packed.shape))
        else:
            gathered = packed.unsqueeze(0)
        per = -(-self.total // self.world)
        bases = torch.tensor([min(self.total, r * per) for r in range(self.world)], dtype=torch.int64, device=packed.device)
        out_d, out_i = self._merge(gathered, bases, k)
        if check and status is not None:
            check_status(status)
        if as_numpy:
            return out_d.cpu().numpy(), out_i.cpu().numpy()
        return out_d, out_i
