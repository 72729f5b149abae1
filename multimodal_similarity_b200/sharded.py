"""Gallery-sharded retrieval over the GPUs of one box (no counterpart in the single-process reference).

One process per GPU (``torch.distributed``, NCCL over NVLink).  The gallery rows are split contiguously over the
ranks, the queries are replicated.  Two protocols, both giving results bit-identical on every rank and equal to the
single-GPU result:

``exact-shards``  every rank computes the exact, locally certified top-k of its shard (``mmsim_knn_f32``), one
                  all-gather of the packed ``[2, Q, k]`` lists, the same merge kernel on every rank.
``reduced``       (default with CUDA tensors) the per-rank work that does not shrink with the number of shards is removed:
                  (1) the pivot pre-pass lists are all-gathered and merged, so all shards filter every query with ONE
                  threshold (about 1000 gallery rows below it in total, not per shard); (2) each shard re-ranks exactly
                  only ``kp`` << 128 candidates and reports a lower bound for everything it did not re-rank; (3) the
                  candidate lists are exchanged BY QUERY SLICE (one all-to-all: rank r receives every shard's lists for
                  queries [r S, (r+1) S)), each rank merges and certifies its slice only, and one all-gather hands every
                  rank the merged top-k of all queries -- the merge no longer repeats on every rank.  (4) Queries the
                  global certificate could not prove (adversarial row order, massive ties) are repaired ONE BY ONE:
                  every rank computes the exact top-k inside its shard for exactly those queries
                  (``mmsim_knn_shard_fallback_f32``: a tensor-core re-sweep bounded by the merged k-th distance), the
                  compact lists are all-gathered and merged into those rows.  Only when more than ``FALLBACK_CAP``
                  queries are uncertified does the whole call repeat as ``exact-shards`` -- the result is always exact.

``query_groups = R`` (2-D decomposition).  What does not shrink with the number of gallery shards is per-QUERY work -- operand
copies and grouping of the replicated queries, the pivot pre-pass, candidate selection, the exchanges (a third of the step at
8 shards).  With R > 1 the ranks form R groups of ``world / R``: the gallery is split over the ranks INSIDE a group (every
group holds the whole gallery), the queries are split over the groups, each group runs the protocol above on its queries with
collectives inside the group only, and ONE all-gather over all ranks assembles the result (every rank contributes the query
slice it merged).  Per-query work and the sweep both shrink with the number of ranks.  ``"auto"`` keeps two gallery parts
per group from 4 ranks on (memory: every row lives on R ranks).

The training-loss kernels are not sharded (replicas only).
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.distributed as dist

from . import _lib
from ._util import is_numpy_like, stream_handle, to_cuda_f32
from .retrieval import check_status, knn_raw

PH_PREP, PH_TENSOR, PH_RERANK, PH_FALLBACK, PH_PIVOT, PH_LADDER, PH_PREP_Q, PH_PREP_G = 1, 2, 4, 8, 16, 32, 128, 256
ZERO_COPY_RESULT = False  # retrieve_host: merge kernel stores its slice into pinned host memory instead of two copies (slower)
FALLBACK_CAP = 1024      # uncertified queries repaired one by one per call; beyond that the call repeats as exact-shards


def shard_bounds(n_rows: int, world: int, rank: int):
    """Contiguous row block of ``rank``: [lo, hi) with ceil(n/world) rows per rank (the last ranks may be short/empty)."""
    per = -(-n_rows // world)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def reduced_kp(world: int, k: int) -> int:
    """Candidates each shard re-ranks exactly in the reduced protocol.  With rows in random order a shard holds
    Binomial(128, 1/world) of the 128 best approximate keys: mean + 6 sigma + 8 (and never fewer than its share of k)."""
    p = 1.0 / world
    kp = int(math.ceil(128 * p + 6.0 * math.sqrt(128 * p * (1 - p)))) + 8
    return min(128, max(kp, -(-k // world) + 8))


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


class ReducedShard:
    """State of one shard in the reduced protocol (its own workspace: the pivot lists live inside it)."""

    def __init__(self, shard: torch.Tensor, lo: int, gallery_host: torch.Tensor | None = None):
        """gallery_host: the shard's rows in page-locked host memory -- every retrieval uploads them again, in pieces, under
        the rest of the pipeline (``mmsim_knn_shard_host_f32``); ``shard`` is then only the device staging buffer."""
        self.shard, self.lo = shard, lo
        self.gallery_host = gallery_host
        self.ws = None

    def _ws(self, nq, d, k):
        lib = _lib.load()
        n = ctypes.c_size_t()
        if self.shard.shape[0] == 0:       # an empty shard (more ranks than ceil-sized blocks) has no workspace and no lists
            return None
        size_fn = lib.mmsim_knn_workspace_bytes if self.gallery_host is None else lib.mmsim_knn_host_workspace_bytes
        _lib.check(size_fn(nq, self.shard.shape[0], d, k, ctypes.byref(n)), "mmsim_knn_workspace_bytes")
        if self.ws is None or self.ws.numel() < n.value:
            self.ws = torch.empty(n.value, dtype=torch.uint8, device=self.shard.device)
        off, nb = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(lib.mmsim_knn_pivot_region(nq, self.shard.shape[0], d, k, ctypes.byref(off), ctypes.byref(nb)),
                   "mmsim_knn_pivot_region")
        return self.ws[off.value:off.value + nb.value].view(torch.float32).view(-1, 16)

    def _call(self, q, k, kp, exclude_self, self_offset, phases, out, slice_rows=0, slice_stride=0):
        lib = _lib.load()
        dev = q.device
        d_, i_, lb_, st_ = out
        with torch.cuda.device(dev):
            if self.gallery_host is None:
                rc = lib.mmsim_knn_shard_f32(q.data_ptr(), q.shape[0], self.shard.data_ptr(), self.shard.shape[0], q.shape[1], k, kp,
                                             int(bool(exclude_self)), int(self_offset - self.lo), d_.data_ptr(), i_.data_ptr(),
                                             lb_.data_ptr(), st_.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                                             stream_handle(dev), phases, slice_rows, slice_stride)
            else:
                rc = lib.mmsim_knn_shard_host_f32(q.data_ptr(), q.shape[0], self.gallery_host.data_ptr(), self.shard.data_ptr(),
                                                  self.shard.shape[0], q.shape[1], k, kp, int(bool(exclude_self)),
                                                  int(self_offset - self.lo), d_.data_ptr(), i_.data_ptr(), lb_.data_ptr(),
                                                  st_.data_ptr(), self.ws.data_ptr(), self.ws.numel(), stream_handle(dev), phases,
                                                  slice_rows, slice_stride)
        _lib.check(rc, "mmsim_knn_shard_f32")

    def stage1(self, q, k, kp, send, status, phases=PH_PREP | PH_PIVOT):
        """Operand copies + pivot pre-pass (or the given part of them); returns this shard's pivot lists [rows, 16] (a view
        into the workspace; None for an empty shard)."""
        piv = self._ws(q.shape[0], q.shape[1], k)
        if piv is None:
            return None
        self._call(q, k, kp, False, 0, phases, (send, send, send, status))     # (these phases write no outputs)
        return piv

    def stage2(self, q, k, kp, exclude_self, self_offset, send, status, S):
        """Threshold ladder from the (merged) pivot lists, sweep, reduced exact re-rank.  The re-rank kernel writes the
        candidate lists straight into ``send`` [world, S (2 kp + 1)] int32, the buffer the all-to-all by query slice sends:
        for every destination rank | distance bits S x kp | shard-local indices S x kp | lower-bound bits S |."""
        stride = S * (2 * kp + 1)
        flat = send.view(-1)
        if self.shard.shape[0] == 0:
            v = send.view(-1, stride)
            v[:, :S * kp] = 2139095040            # +inf bits
            v[:, S * kp:2 * S * kp] = -1
            v[:, 2 * S * kp:] = 2139095040
            status.zero_()
            return
        d_ = flat.view(torch.float32)
        self._call(q, k, kp, exclude_self, self_offset, PH_LADDER | PH_TENSOR | PH_RERANK,
                   (d_, flat[S * kp:], d_[2 * S * kp:], status), S, stride)

    def fallback(self, q, k, exclude_self, self_offset, flag, cap):
        """Exact top-k inside this shard for the queries with flag >= 0 (``mmsim_knn_shard_fallback_f32``) -> one int32
        buffer | distance bits cap x k | shard-local indices cap x k | query of each slot cap | status 8 |."""
        dev = q.device
        buf = torch.empty(2 * cap * k + cap + 8, dtype=torch.int32, device=dev)
        d_ = buf[:cap * k]
        i_ = buf[cap * k:2 * cap * k]
        qq = buf[2 * cap * k:2 * cap * k + cap]
        st = buf[2 * cap * k + cap:]
        if self.shard.shape[0] == 0:
            d_.view(torch.float32).fill_(float("inf")); i_.fill_(-1); qq.zero_(); st.zero_()
            return buf
        lib = _lib.load()
        with torch.cuda.device(dev):
            rc = lib.mmsim_knn_shard_fallback_f32(q.data_ptr(), q.shape[0], self.shard.data_ptr(), self.shard.shape[0], q.shape[1], k,
                                                  int(bool(exclude_self)), int(self_offset - self.lo), flag.data_ptr(), cap,
                                                  d_.data_ptr(), i_.data_ptr(), qq.data_ptr(), st.data_ptr(), self.ws.data_ptr(),
                                                  self.ws.numel(), stream_handle(dev), int(self.gallery_host is not None))
        _lib.check(rc, "mmsim_knn_shard_fallback_f32")
        return buf


def merge_pivots_into(piv_parts: torch.Tensor, out: torch.Tensor):
    """[parts, rows, 16] pivot lists -> out [rows, 16]: the 16 smallest of the union (mmsim_knn_merge_pivots)."""
    lib = _lib.load()
    dev = out.device
    with torch.cuda.device(dev):
        rc = lib.mmsim_knn_merge_pivots(piv_parts.data_ptr(), piv_parts.shape[0], piv_parts.stride(0), out.shape[0],
                                        out.data_ptr(), stream_handle(dev))
    _lib.check(rc, "mmsim_knn_merge_pivots")


def slice_rows(nq: int, world: int) -> int:
    """Queries per merge slice: ceil(nq / world) (the last slices may be short or empty)."""
    return -(-nq // world)


def merge_certified_slice(recv: torch.Tensor, bases: torch.Tensor, n_rows: int, S: int, kp: int, k: int, idx32: bool, out=None):
    """[parts, S (2 kp + 1)] candidate lists of ONE query slice (what the all-to-all delivers) -> merged
    (dist [S, k] f32, idx [S, k] int32 or int64 GLOBAL indices, meta [S + 8] int32 = | flag S (float bits: the merged k-th
    distance of an uncertified query, -1 for a certified one or a row past the slice's end) | status 8 |)."""
    lib = _lib.load()
    dev = recv.device
    parts, stride = recv.shape[0], recv.stride(0)
    if out is None:
        out_d = torch.empty((S, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((S, k), dtype=torch.int32 if idx32 else torch.int64, device=dev)
    else:       # caller's buffers, e.g. page-locked host memory (device-accessible): the kernel's stores ARE the transfer
        out_d, out_i = out
        assert out_d.shape == (S, k) and out_i.shape == (S, k) and out_i.dtype == (torch.int32 if idx32 else torch.int64)
    meta = torch.zeros(S + 8, dtype=torch.int32, device=dev)
    meta[:S] = -1082130432                         # bits of -1.0f
    base = recv.data_ptr()
    if n_rows > 0:
        with torch.cuda.device(dev):
            rc = lib.mmsim_knn_merge_certified(base, base + S * kp * 4, stride, bases.data_ptr(), parts, n_rows, kp, k,
                                               base + 2 * S * kp * 4, stride, out_d.data_ptr(), out_i.data_ptr(),
                                               32 if idx32 else 64, meta.data_ptr() + S * 4, meta.data_ptr(), stream_handle(dev))
        _lib.check(rc, "mmsim_knn_merge_certified")
    return out_d, out_i, meta


def patch_rows(out_d, out_i, allfb: torch.Tensor, bases: torch.Tensor, cap: int, k: int):
    """[world, 2 cap k + cap + 8] gathered fallback buffers -> rows of (out_d, out_i int64) rewritten with the merge of the
    shards' exact lists (``mmsim_knn_merge_patch``).  Slot -> query map and count are rank 0's (identical on every rank whose
    shard is not empty; an empty shard contributes +inf lists)."""
    lib = _lib.load()
    dev = out_d.device
    world, stride = allfb.shape[0], allfb.stride(0)
    base = allfb.data_ptr()
    src = 0                                           # rank 0's shard is never empty (shard_bounds)
    row_map = base + (src * stride + 2 * cap * k) * 4
    count = base + (src * stride + 2 * cap * k + cap) * 4
    with torch.cuda.device(dev):
        rc = lib.mmsim_knn_merge_patch(base, base + cap * k * 4, stride, bases.data_ptr(), world, cap, k, count, row_map,
                                       out_d.data_ptr(), out_i.data_ptr(), stream_handle(dev))
    _lib.check(rc, "mmsim_knn_merge_patch")


class GraphedRetrieve:
    """See ``ShardedGallery.graphed``."""

    def __init__(self, sg, queries, k, exclude_self, self_offset):
        if not (torch.is_tensor(queries) and queries.is_cuda):
            raise ValueError("graphed() needs the queries as a CUDA tensor (the captured buffer)")
        self.sg, self.queries = sg, to_cuda_f32(queries, sg.shard.device)
        dev = self.queries.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):           # warm-up outside the capture: workspaces, kernel attributes, NCCL channels
            for _ in range(2):
                sg.retrieve(self.queries, k, exclude_self=exclude_self, self_offset=self_offset, check=False)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = sg.retrieve(self.queries, k, exclude_self=exclude_self, self_offset=self_offset, check=False)
            self.uncertified = sg.last_uncertified if sg.last_uncertified is not None else torch.zeros((), dtype=torch.int32, device=dev)
        self.protocol = sg.last_protocol

    def __call__(self, queries=None):
        if queries is not None:
            self.queries.copy_(queries, non_blocking=True)
        self.graph.replay()
        return self.out

    def close(self):
        """Release the captured graph (it holds NCCL kernels): call this on every rank BEFORE ``destroy_process_group`` --
        tearing the communicator down under a live graph hung a 2-GPU run of this round in its teardown."""
        torch.cuda.synchronize(self.queries.device)
        self.out = self.uncertified = None
        self.graph = None


def resolve_query_groups(query_groups, world: int) -> int:
    """``"auto"``: two gallery parts per query group from 4 ranks on (R = world / 2), otherwise one group."""
    if query_groups == "auto":
        return world // 2 if world >= 4 and world % 2 == 0 else 1
    r = int(query_groups)
    if r < 1 or world % r:
        raise ValueError(f"query_groups={query_groups} must divide the number of ranks ({world})")
    return r


class ShardedGallery:
    """``ShardedGallery(gallery, group).retrieve(queries, k)`` -> identical (dist, idx) on every rank.

    gallery      the FULL [G, D] gallery (each rank keeps only its slice) or, with ``presharded=True``, this rank's
                 rows together with ``row_offset`` / ``total_rows`` (the rows ``shard_bounds`` assigns to it)
    group        a torch.distributed process group (default: the world); without an initialised process group the
                 object degenerates to a single shard
    query_groups R: the ranks form R groups of ``parts = world / R``; rank r holds gallery part ``r % parts`` (the rows
                 ``shard_bounds(total_rows, parts, r % parts)``) and serves the queries of group ``r // parts``
                 (see the module docstring); 1 = every rank holds a different shard and sees every query; ``"auto"``
    """

    def __init__(self, gallery, group=None, *, presharded=False, row_offset=None, total_rows=None, device=None, query_groups=1):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.query_groups = resolve_query_groups(query_groups, self.world)
        self.parts = self.world // self.query_groups       # gallery parts == ranks per query group
        self.part = self.rank % self.parts
        self.qgroup = self.rank // self.parts
        self.sub_group = group                             # the ranks that share my queries (collectives of the protocols)
        if self.query_groups > 1 and self.parts > 1:
            members = dist.get_process_group_ranks(group) if group is not None else list(range(self.world))
            for g_ in range(self.query_groups):            # every rank creates every group (torch.distributed contract)
                sub = dist.new_group([members[g_ * self.parts + j] for j in range(self.parts)])
                if g_ == self.qgroup:
                    self.sub_group = sub
        if presharded:
            if row_offset is None or total_rows is None:
                raise ValueError("presharded=True needs row_offset and total_rows")
            self.lo, self.total = int(row_offset), int(total_rows)
            shard = gallery
            want = shard_bounds(self.total, self.parts, self.part)
            if (self.lo, self.lo + int(gallery.shape[0])) != want:
                # every rank derives the other shards' row offsets from (total_rows, parts) alone -- see _bases()
                raise ValueError(f"presharded rows [{self.lo}, {self.lo + int(gallery.shape[0])}) of rank {self.rank} are not "
                                 f"shard_bounds({self.total}, {self.parts}, {self.part}) = {want}")
        else:
            self.total = int(gallery.shape[0])
            self.lo, hi = shard_bounds(self.total, self.parts, self.part)
            shard = gallery[self.lo:hi]
        self.shard = self._to_device(shard, device)
        self.hi = self.lo + int(self.shard.shape[0])
        self.dim = int(gallery.shape[1])
        self._reduced = None
        self._copy_stream, self._q_all, self._host_out = None, None, None     # retrieve_host: copy stream, staging, pinned results
        self._reduced_host = None                                              # ... and the shard state fed from host memory
        self.last_protocol = None
        self.last_uncertified = None      # reduced protocol: device scalar, queries the global certificate did not prove
        self.last_repaired = 0            # ... and how many of them the last call repaired one by one

    def _my_queries(self, nq: int):
        """(S, q_lo, q_hi): rows per merge slice and the query range of my group.  Rank r merges the queries
        [r S, (r + 1) S): group g serves [g parts S, (g + 1) parts S)."""
        S = slice_rows(nq, self.world)
        q_lo = min(nq, self.qgroup * self.parts * S)
        return S, q_lo, min(nq, q_lo + self.parts * S)

    # -- steps of the exact-shards protocol; tests on CPU boxes replace them to exercise the sharding logic under gloo
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

    def _bases(self, device):
        """First gallery row of every part of my group (cached: a host -> device copy cannot be captured in a CUDA graph)."""
        key = str(device)
        if getattr(self, "_bases_cache", None) is None or self._bases_cache[0] != key:
            per = -(-self.total // self.parts)
            self._bases_cache = (key, torch.tensor([min(self.total, r * per) for r in range(self.parts)], dtype=torch.int64,
                                                   device=device))
        return self._bases_cache[1]

    def graphed(self, queries, k, *, exclude_self=False, self_offset=0):
        """``retrieve`` for a FIXED query buffer captured once as a CUDA graph (kernels + NCCL collectives) and replayed:
        ``g = sg.graphed(q, k); dist, idx = g()`` (``g(new_queries)`` copies them into the captured buffer first).  At 8 GPUs a
        100k-query step is ~4 ms made of ~30 kernels and 6 collectives issued from Python; replaying the graph removes the
        launch gaps between them.  Every rank must capture and replay in step (the collectives are part of the graph).
        The replay does not repair uncertified queries: ``g.uncertified`` (device scalar) says how many the last replay had --
        call ``retrieve`` when it is not zero."""
        return GraphedRetrieve(self, queries, int(k), exclude_self, self_offset)

    def _retrieve_exact_shards(self, q, k, exclude_self, self_offset, check):
        nq = q.shape[0]
        S, q_lo, q_hi = self._my_queries(nq)
        if self.query_groups > 1:
            q = q[q_lo:q_hi]
        else:
            q_lo = 0
        if nq == 0:
            dev = self.shard.device
            return torch.empty((0, k), dtype=torch.float32, device=dev), torch.empty((0, k), dtype=torch.int64, device=dev)
        packed, status = None, None
        if q.shape[0] > 0:
            packed, status = self._local(q, k, exclude_self, self_offset + q_lo)
        if check and self.world > 1:
            # finish queued exact scans BEFORE the collectives, and agree on the outcome: either every rank raises or none
            err = torch.zeros(1, dtype=torch.int32, device=self.shard.device)
            try:
                if status is not None:
                    check_status(status)
            except _lib.MmsimError:
                err += 1
            dist.all_reduce(err, op=dist.ReduceOp.MAX, group=self.group)
            if int(err):
                raise _lib.MmsimError("knn: the exact fallback failed on at least one gallery shard")
        elif check and status is not None:
            check_status(status)
        merged = None
        if packed is not None:
            if self.parts > 1:
                # dim-0 concatenation layout (accepted by both NCCL and gloo), viewed as [parts, 2, Q, k] afterwards
                flat = torch.empty((self.parts * 2,) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
                dist.all_gather_into_tensor(flat, packed, group=self.sub_group)
                gathered = flat.view((self.parts,) + tuple(packed.shape))
            else:
                gathered = packed.unsqueeze(0)
            merged = self._merge(gathered, self._bases(packed.device), k)
        if self.query_groups == 1:
            return merged
        # 2-D decomposition: every rank contributes the S rows of its group's result that it would merge in the reduced
        # protocol; one all-gather over ALL ranks assembles the queries in order
        dev = self.shard.device
        my_d = torch.full((S, k), float("inf"), dtype=torch.float32, device=dev)
        my_i = torch.full((S, k), -1, dtype=torch.int64, device=dev)
        if merged is not None:
            rows = merged[0][self.part * S:self.part * S + S]
            my_d[:rows.shape[0]] = rows
            my_i[:rows.shape[0]] = merged[1][self.part * S:self.part * S + S]
        out_d = torch.empty((self.world * S, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((self.world * S, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(out_d, my_d, group=self.group)
        dist.all_gather_into_tensor(out_i, my_i, group=self.group)
        return out_d[:nq], out_i[:nq]

    def _reduced_slice(self, q, k, exclude_self, self_offset, rs=None, gallery_queued=False, S=None, out=None):
        """Stages of the reduced protocol up to this rank's merged query slice: (dist [S, k], idx [S, k] global, meta).
        ``q``: the queries of my group; ``S``: rows per merge slice (default: ceil(len(q) / parts)).
        ``rs``: the shard state to use (end-to-end path: one whose gallery rows come from the host; its gallery copies were
        queued by the caller already -- ``gallery_queued`` -- so only the query half of the preparation is left)."""
        nq = q.shape[0]
        kp = reduced_kp(self.parts, k)
        if rs is None:
            if self._reduced is None:
                self._reduced = ReducedShard(self.shard, self.lo)
            rs = self._reduced
        dev = q.device
        if S is None:
            S = slice_rows(nq, self.parts)
        send = torch.empty((self.parts, S * (2 * kp + 1)), dtype=torch.int32, device=dev)
        status = torch.empty(8, dtype=torch.int32, device=dev)
        piv = rs.stage1(q, k, kp, send, status, phases=(PH_PREP_Q if gallery_queued else PH_PREP) | PH_PIVOT)
        rows = -(-nq // 128) * 128
        mine = piv if piv is not None else torch.full((rows, 16), float("inf"), device=dev)
        allpiv = torch.empty((self.parts * rows, 16), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allpiv, mine, group=self.sub_group)
        if piv is not None:
            merge_pivots_into(allpiv.view(self.parts, rows, 16), piv)
        rs.stage2(q, k, kp, exclude_self, self_offset, send, status, S)
        # merge by query slice: all-to-all of the candidate lists, merge + certify my slice
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.sub_group)
        n_mine = max(0, min(S, nq - self.part * S))
        if out is not None:      # (dist, idx int64) buffers of the caller: the merge kernel writes the slice there
            return merge_certified_slice(recv, self._bases(dev), n_mine, S, kp, k, False, out=out) + (S,)
        return merge_certified_slice(recv, self._bases(dev), n_mine, S, kp, k, self.total < 2 ** 31) + (S,)

    def _repair(self, q, k, exclude_self, self_offset, out_d, out_i, flag, n_unc):
        """Per-query repair of the rows of MY GROUP's queries the global certificate did not prove (``q``, ``out_d``,
        ``out_i``, ``flag``: the group's rows; identical decisions on every rank of the group).  Returns False when the call
        has to be repeated as exact-shards."""
        if n_unc > FALLBACK_CAP:
            return False
        dev = q.device
        fb = self._reduced.fallback(q, k, exclude_self, self_offset, flag, FALLBACK_CAP)
        allfb = torch.empty((self.parts, fb.numel()), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allfb.view(-1), fb, group=self.sub_group)
        st = allfb[:, 2 * FALLBACK_CAP * k + FALLBACK_CAP:].cpu()
        if bool((st[:, 1] > st[:, 2]).any()):         # a shard's streaming scan has queries left: identical view in the group
            return False
        patch_rows(out_d, out_i, allfb, self._bases(dev), FALLBACK_CAP, k)
        return True

    def _retrieve_reduced(self, q, k, exclude_self, self_offset, check):
        """Returns (dist, idx), or None when the call has to be repeated as exact-shards (more than FALLBACK_CAP
        uncertified queries in a group, or a shard's streaming scan still has queries queued)."""
        nq = q.shape[0]
        dev = q.device
        S, q_lo, q_hi = self._my_queries(nq)
        idx_dtype = torch.int32 if self.total < 2 ** 31 else torch.int64
        if q_hi > q_lo:
            my_d, my_i, meta, _ = self._reduced_slice(q[q_lo:q_hi], k, exclude_self, self_offset + q_lo, S=S)
        else:                                             # more query groups than slices with queries: nothing to do here
            my_d = torch.full((S, k), float("inf"), dtype=torch.float32, device=dev)
            my_i = torch.full((S, k), -1, dtype=idx_dtype, device=dev)
            meta = torch.zeros(S + 8, dtype=torch.int32, device=dev)
            meta[:S] = -1082130432
        # the merged slices go straight into the final arrays (three all-gathers over ALL ranks, no repacking)
        out_d = torch.empty((self.world * S, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((self.world * S, k), dtype=my_i.dtype, device=dev)
        allmeta = torch.empty((self.world, S + 8), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(out_d, my_d, group=self.group)
        dist.all_gather_into_tensor(out_i, my_i, group=self.group)
        dist.all_gather_into_tensor(allmeta.view(-1), meta, group=self.group)
        out_d, out_i = out_d[:nq], out_i[:nq]
        if out_i.dtype != torch.int64:
            out_i = out_i.to(torch.int64)
        uncertified = allmeta[:, S].sum()
        self.last_uncertified = uncertified           # device scalar, identical on every rank (part of the gathered buffer)
        if not check:
            return out_d, out_i                       # the caller inspects last_uncertified (bench.py does, after timing)
        per_rank = allmeta[:, S].cpu()
        n_unc = int(per_rank.sum())
        if n_unc == 0:
            return out_d, out_i
        per_group = per_rank.view(self.query_groups, self.parts).sum(1)
        if int(per_group.max()) > FALLBACK_CAP:       # identical view on every rank: all of them repeat the call
            return None
        flag = allmeta[:, :S].reshape(-1)[:nq].contiguous().view(torch.float32)
        ok = True
        if int(per_group[self.qgroup]) > 0:
            ok = self._repair(q[q_lo:q_hi], k, exclude_self, self_offset + q_lo, out_d[q_lo:q_hi], out_i[q_lo:q_hi],
                              flag[q_lo:q_hi].contiguous(), int(per_group[self.qgroup]))
        if self.query_groups == 1:
            if ok:
                self.last_repaired = n_unc
            return (out_d, out_i) if ok else None
        # 2-D decomposition: a group repaired its own rows only; hand them to the other groups (and agree on failures)
        rows = torch.nonzero(flag >= 0).view(-1)                                # all uncertified queries, ascending
        mine = (rows >= q_lo) & (rows < q_hi)
        pd = torch.where(mine[:, None], out_d[rows], torch.full_like(out_d[rows], float("-inf")))
        pi = torch.where(mine[:, None], out_i[rows], torch.full_like(out_i[rows], -1))
        bad = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(pd, op=dist.ReduceOp.MAX, group=self.group)            # exactly one group owns each row
        dist.all_reduce(pi, op=dist.ReduceOp.MAX, group=self.group)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
        if int(bad):
            return None
        out_d[rows] = pd
        out_i[rows] = pi
        self.last_repaired = n_unc
        return out_d, out_i

    def retrieve_host(self, queries_host, k, *, gallery_host=None, exclude_self=False, self_offset=0):
        """End-to-end form of ``retrieve`` for inputs in page-locked HOST memory (reduced protocol, CUDA only).

        queries_host   the full [Q, D] float32 query array, the same on every rank (SPMD).  Each rank uploads only ITS 1/world
                       slice over PCIe; the slices are all-gathered over NVLink.
        gallery_host   optional: this rank's shard rows in page-locked host memory, uploaded again by this call (the
                       benchmark's end-to-end contract: every input starts on the host).  The library copies it in pieces on
                       its own stream (``mmsim_knn_shard_host_f32``): the pivot sample first, then split by split, under the
                       query exchange, the query preparation and the sweeps of the earlier splits.
        Returns (dist [n, k] float32, idx [n, k] int64 global indices, (q_lo, q_hi)): THIS rank's query slice of the result in
        page-locked host memory -- the ranks' slices tile the queries, no rank downloads what another one already has.
        A query the global certificate cannot prove makes the call fall back to ``retrieve`` (every rank, same decision)."""
        if self.world == 1 or not self.shard.is_cuda:
            raise _lib.MmsimError("retrieve_host is the multi-GPU end-to-end path; use retrieval.retrieve_host on one GPU")
        if self.query_groups != 1:
            raise _lib.MmsimError("retrieve_host uploads every gallery row once: build the ShardedGallery with query_groups=1")
        k = int(k)
        dev = self.shard.device
        nq, d = queries_host.shape
        if d != self.dim:
            raise ValueError(f"queries are {d}-d but the gallery is {self.dim}-d")
        S = slice_rows(nq, self.world)
        q_lo, q_hi = min(nq, self.rank * S), min(nq, self.rank * S + S)
        if gallery_host is not None and (self._reduced_host is None or self._reduced_host.gallery_host is not gallery_host):
            self._reduced_host = ReducedShard(self.shard, self.lo, gallery_host)
        rs = self._reduced_host if gallery_host is not None else None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
        if self._q_all is None or tuple(self._q_all.shape) != (self.world * S, d):
            self._q_all = torch.zeros((self.world * S, d), dtype=torch.float32, device=dev)
        main = torch.cuda.current_stream(dev)
        start = torch.cuda.Event(); start.record(main)
        ev_q = torch.cuda.Event()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(start)          # earlier work on the caller's stream may still read these buffers
            if q_hi > q_lo:
                self._q_all[q_lo:q_hi].copy_(queries_host[q_lo:q_hi], non_blocking=True)
            ev_q.record()
        q = self._q_all[:nq]
        # the query slice is on the critical path (exchange -> preparation -> pre-pass -> sweep) and shares the PCIe link with
        # the shard: it goes first, alone -- the library's copy stream starts after everything queued on the caller's stream
        main.wait_event(ev_q)
        if rs is not None and self.shard.shape[0] > 0:
            # queue the shard's host -> device copies now (sample first, then split by split: the library's copy stream);
            # they run under the query exchange, the query preparation and the sweeps of the earlier splits
            kp = reduced_kp(self.world, k)
            dummy = torch.empty(8, dtype=torch.int32, device=dev)
            rs._ws(nq, d, k)
            rs._call(q, k, kp, False, 0, PH_PREP_G, (dummy, dummy, dummy, dummy))
        dist.all_gather_into_tensor(self._q_all, self._q_all[self.rank * S:self.rank * S + S], group=self.group)
        n = q_hi - q_lo
        if self._host_out is None or self._host_out[0].shape != (S, k):
            self._host_out = (torch.empty((S, k), dtype=torch.float32).pin_memory(), torch.empty((S, k), dtype=torch.int64).pin_memory())
        h_d, h_i = self._host_out
        if ZERO_COPY_RESULT:
            # the merge kernel writes this rank's slice (float32 distances, int64 global indices) straight into the page-locked
            # host buffers (device-accessible under UVA).  Measured at N = 2: 16.2 ms per step against 15.7 with the copies
            # below -- one warp per query storing 1.2 KB rows over PCIe is slower than the copy engine; kept as a switch.
            _, _, meta, S = self._reduced_slice(q, k, exclude_self, self_offset, rs=rs,
                                                gallery_queued=rs is not None and self.shard.shape[0] > 0, out=(h_d, h_i))
        else:
            my_d, my_i, meta, S = self._reduced_slice(q, k, exclude_self, self_offset, rs=rs,
                                                       gallery_queued=rs is not None and self.shard.shape[0] > 0)
            h_d[:n].copy_(my_d[:n], non_blocking=True)
            h_i[:n].copy_(my_i[:n].to(torch.int64), non_blocking=True)      # (widened on the device: 12 k x 100 words)
        # uncertified queries anywhere?  (a 4-byte all-reduce; the copies above are in flight meanwhile)
        unc = meta[S:S + 1].clone()
        dist.all_reduce(unc, group=self.group)
        self.last_uncertified = unc[0]
        if int(unc) != 0:                                 # synchronises: the host buffers are complete as well
            out_d, out_i = self.retrieve(q, k, exclude_self=exclude_self, self_offset=self_offset)
            return out_d[q_lo:q_hi].cpu(), out_i[q_lo:q_hi].cpu(), (q_lo, q_hi)
        return h_d[:n], h_i[:n], (q_lo, q_hi)

    def retrieve(self, queries, k, *, exclude_self=False, self_offset=0, check=True, protocol="auto"):
        as_numpy = is_numpy_like(queries)
        q = self._to_device(queries, self.shard.device)
        if q.shape[1] != self.dim:
            raise ValueError(f"queries are {q.shape[1]}-d but the gallery is {self.dim}-d")
        if protocol not in ("auto", "reduced", "exact-shards"):
            raise ValueError("protocol must be auto, reduced or exact-shards")
        k = int(k)
        out = None
        self.last_protocol = "exact-shards"
        self.last_repaired = 0
        if protocol != "exact-shards" and self.parts > 1 and q.is_cuda and reduced_kp(self.parts, k) < 128:
            out = self._retrieve_reduced(q, k, exclude_self, self_offset, check)
            if out is not None:
                self.last_protocol = "reduced"
        if out is None:
            out = self._retrieve_exact_shards(q, k, exclude_self, self_offset, check)
        out_d, out_i = out
        if as_numpy:
            return out_d.cpu().numpy(), out_i.cpu().numpy()
        return out_d, out_i
