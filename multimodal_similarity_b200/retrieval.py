"""Query x gallery retrieval -- drop-ins for the reference's evaluation path (src/utils.py:55-266).

``retrieve`` is the GPU form of the per-query ``retrieve_one`` loop: distances are the reference's float32
``np.linalg.norm(q - G, axis=1)`` bit for bit, neighbours are ordered by (distance, index).  The tcgen05 kernel
only *filters*; every returned distance is recomputed in the reference arithmetic and the result is certified
(or recomputed exactly), see csrc/knn_tc.cu.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._util import is_numpy_like, stream_handle, to_cuda_f32, workspace

RECALL_KS = (1, 2, 4, 8, 16, 32)  # src/utils.py:190-197


def late_fusion(*modalities):
    """Feature concatenation of independently normalised modalities (src/evaluate_late_fusion.py:115,
    src/evaluate_hallucination.py:59): d^2_fused = sum of the per-modality d^2."""
    if all(is_numpy_like(m) for m in modalities):
        return np.concatenate([np.asarray(m, dtype=np.float32) for m in modalities], axis=1)
    dev = next(m.device for m in modalities if torch.is_tensor(m) and m.is_cuda)
    return torch.cat([to_cuda_f32(m, dev) for m in modalities], dim=1)


def knn_raw(q: torch.Tensor, g: torch.Tensor, k: int, exclude_self: bool = False, self_offset: int = 0, *, phases: int = 63,
            out=None):
    """Low-level call: CUDA float32 tensors in, (dist [Q,k] f32, idx [Q,k] i32 shard-local, status [8] i32) out.
    Asynchronous on the current stream; ``status`` is checked by the callers that hand results to the user."""
    lib = _lib.load()
    nq, d = q.shape
    ng = g.shape[0]
    if g.shape[1] != d:
        raise ValueError(f"queries are {d}-d but the gallery is {g.shape[1]}-d")
    if not 1 <= k <= _lib.KNN_MAX_K - (1 if exclude_self else 0):
        raise ValueError(f"k={k} unsupported: 1 <= k <= {_lib.KNN_MAX_K - (1 if exclude_self else 0)}")
    dev = q.device
    nbytes = ctypes.c_size_t()
    _lib.check(lib.mmsim_knn_workspace_bytes(nq, ng, d, k, ctypes.byref(nbytes)), "mmsim_knn_workspace_bytes")
    ws = workspace("knn", nbytes.value, dev)
    if out is None:
        dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
        status = torch.empty(8, dtype=torch.int32, device=dev)
    else:
        dist, idx, status = out
    with torch.cuda.device(dev):
        rc = lib.mmsim_knn_f32_phases(q.data_ptr(), nq, g.data_ptr(), ng, d, k, int(bool(exclude_self)), int(self_offset),
                                      dist.data_ptr(), idx.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(),
                                      stream_handle(dev), int(phases))
    _lib.check(rc, "mmsim_knn_f32")
    status._mmsim_finish = _finisher(q, g, k, exclude_self, self_offset, dist, idx, status, ws)
    return dist, idx, status


def _finisher(q, g, k, exclude_self, self_offset, dist, idx, status, ws):
    """One more wave of the streaming exact scan for the queries a call left queued (``mmsim_knn_finish_f32``); q / g are
    the device arrays of that call, whose workspace still holds the queue."""
    def run():
        lib = _lib.load()
        dev = q.device
        with torch.cuda.device(dev):
            rc = lib.mmsim_knn_finish_f32(q.data_ptr(), q.shape[0], g.data_ptr(), g.shape[0], q.shape[1], k, int(bool(exclude_self)),
                                          int(self_offset), dist.data_ptr(), idx.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                          ws.numel(), stream_handle(dev))
        _lib.check(rc, "mmsim_knn_finish_f32")
    return run


def knn_host(q_host: torch.Tensor, g_host: torch.Tensor, k: int, exclude_self: bool = False, self_offset: int = 0, *,
             device=None, stage=None, out=None):
    """``knn_raw`` for arrays in HOST memory (``mmsim_knn_host_f32``): CPU float32 tensors in -- page-locked
    (``.pin_memory()``) for the transfers to overlap the kernels -- and the same (dist, idx, status) out.  The gallery
    goes over one split at a time while the previous split is being swept, so only the first split's copy is exposed.
    ``stage`` = (q_stage [Q,D], g_stage [G,D]) CUDA float32 buffers for the device copies (allocated when omitted; keep
    them between calls); ``out`` = (dist, idx, status) with dist / idx in device or page-locked host memory."""
    lib = _lib.load()
    _lib.require_cuda(torch)
    for name, t in (("q_host", q_host), ("g_host", g_host)):
        if not (torch.is_tensor(t) and t.device.type == "cpu" and t.dtype == torch.float32 and t.dim() == 2 and t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous 2-d float32 CPU tensor")
    nq, d = q_host.shape
    ng = g_host.shape[0]
    if g_host.shape[1] != d:
        raise ValueError(f"queries are {d}-d but the gallery is {g_host.shape[1]}-d")
    if not 1 <= k <= _lib.KNN_MAX_K - (1 if exclude_self else 0):
        raise ValueError(f"k={k} unsupported: 1 <= k <= {_lib.KNN_MAX_K - (1 if exclude_self else 0)}")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if stage is None:
        stage = (torch.empty((nq, d), dtype=torch.float32, device=dev), torch.empty((ng, d), dtype=torch.float32, device=dev))
    q_stage, g_stage = stage
    if tuple(q_stage.shape) != (nq, d) or tuple(g_stage.shape) != (ng, d) or not (q_stage.is_contiguous() and g_stage.is_contiguous()) \
            or q_stage.dtype != torch.float32 or g_stage.dtype != torch.float32 or q_stage.device != dev or g_stage.device != dev:
        raise ValueError("stage must be contiguous float32 CUDA tensors shaped like q_host and g_host")
    nbytes = ctypes.c_size_t()
    _lib.check(lib.mmsim_knn_host_workspace_bytes(nq, ng, d, k, ctypes.byref(nbytes)), "mmsim_knn_host_workspace_bytes")
    ws = workspace("knn_host", nbytes.value, dev)
    if out is None:
        dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
        status = torch.empty(8, dtype=torch.int32, device=dev)
    else:
        dist, idx, status = out
    with torch.cuda.device(dev):
        rc = lib.mmsim_knn_host_f32(q_host.data_ptr(), nq, g_host.data_ptr(), ng, d, k, int(bool(exclude_self)), int(self_offset),
                                    dist.data_ptr(), idx.data_ptr(), status.data_ptr(), q_stage.data_ptr(), g_stage.data_ptr(),
                                    ws.data_ptr(), ws.numel(), stream_handle(dev))
    _lib.check(rc, "mmsim_knn_host_f32")
    status._mmsim_finish = _finisher(q_stage, g_stage, k, exclude_self, self_offset, dist, idx, status, ws)
    return dist, idx, status


def check_status(status: torch.Tensor) -> int:
    """Synchronising check of a knn status word; returns the number of queries that took the exact fallback.
    status[1] > status[2]: queries are still queued for the streaming exact scan (the call itself runs one wave of 128) --
    further waves are launched here until none are left, so the result is exact on return whatever the input."""
    s = status.tolist()
    while s[2] < s[1]:
        finish = getattr(status, "_mmsim_finish", None)
        if finish is None:
            raise _lib.MmsimError(f"knn: {s[1] - s[2]} queries are still queued for the exact scan and this status word does "
                                  "not come from knn_raw / knn_host (call mmsim_knn_finish_f32 with the original arguments)")
        before = s[2]
        finish()
        s = status.tolist()
        if s[2] <= before:
            raise _lib.MmsimError(f"knn: the exact scan made no progress (status={s})")
    return s[0]


def retrieve(queries, gallery, k, *, exclude_self=False, self_offset=0, queries2=None, gallery2=None, check=True):
    """Top-k gallery rows for every query.

    Returns (dist [Q,k] float32, idx [Q,k] int64) -- the first k entries of the reference's
    ``retrieve_one(q, gallery)`` (dist[idx[:k]], idx[:k]) for each query, ties ordered by index.
    ``exclude_self`` removes gallery row ``self_offset + i`` for query i (leave-one-out, indices stay in the
    gallery's numbering).  ``queries2`` / ``gallery2`` are a second modality, fused by concatenation (late fusion).
    NumPy in -> NumPy out; CUDA tensors in -> CUDA tensors out.
    """
    as_numpy = is_numpy_like(queries) and is_numpy_like(gallery)
    if (queries2 is None) != (gallery2 is None):
        raise ValueError("late fusion needs both queries2 and gallery2")
    if queries2 is not None:
        queries, gallery = late_fusion(queries, queries2), late_fusion(gallery, gallery2)
    q = to_cuda_f32(queries)
    g = q if gallery is queries else to_cuda_f32(gallery, q.device)
    k = int(k)
    if k > _lib.KNN_MAX_K - (1 if exclude_self else 0) or q.shape[1] > _lib.KNN_MAX_D:
        dist, idx = _retrieve_exact_any(q, g, k, exclude_self, self_offset)
    else:
        dist, idx, status = knn_raw(q, g, k, exclude_self, self_offset)
        if check:
            check_status(status)
        idx = idx.to(torch.int64)
    if as_numpy:
        return dist.cpu().numpy(), idx.cpu().numpy()
    return dist, idx


def _retrieve_exact_any(q, g, k, exclude_self, self_offset):
    """The shapes outside the tensor-core pipeline (k > 112 -- up to the reference's full argsort -- or D > 256): exact
    distances in the reference's arithmetic (csrc/sqdist.cu, any D) for a block of queries at a time, IEEE square root, a
    stable device sort by distance (ties by index, like every other path), the first k columns.  O(G log G) per query:
    the per-query form of the reference (src/utils.py:73-74), not the fast path."""
    from .distance import pairwise_distance
    nq, ng = q.shape[0], g.shape[0]
    if k < 1 or k > ng - (1 if exclude_self else 0):
        raise ValueError(f"k={k} exceeds the {ng - (1 if exclude_self else 0)} gallery rows a query can be ranked against")
    out_d = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    block = max(1, min(nq, (1 << 27) // max(ng, 1)))            # <= 512 MB of distances at a time
    for lo in range(0, nq, block):
        hi = min(nq, lo + block)
        d = torch.sqrt(pairwise_distance(q[lo:hi], g))
        if exclude_self:
            rows = torch.arange(lo, hi, device=q.device)
            cols = rows + int(self_offset)
            ok = (cols >= 0) & (cols < ng)
            d[rows[ok] - lo, cols[ok]] = float("inf")
        sd, si = torch.sort(d, dim=1, stable=True)
        out_d[lo:hi] = sd[:, :k]
        out_i[lo:hi] = si[:, :k]
    return out_d, out_i


def _as_host_f32(x) -> torch.Tensor:
    if torch.is_tensor(x):
        return x.detach().to(device="cpu", dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


def retrieve_host(queries, gallery, k, *, exclude_self=False, self_offset=0, check=True, device=None):
    """``retrieve`` for embeddings that live in host memory (NumPy arrays or CPU tensors), as they do in the reference's
    evaluation scripts: (dist [Q,k] float32, idx [Q,k] int64) NumPy arrays out.  The transfers are part of the call and
    overlap it (``knn_host``): the gallery is swept split by split as it arrives and the re-rank kernel writes the result
    rows straight into page-locked host memory.  Page-locked inputs (``torch.Tensor.pin_memory``) overlap fully."""
    q, g = _as_host_f32(queries), _as_host_f32(gallery)
    if q.dim() != 2 or g.dim() != 2:
        raise ValueError(f"queries and gallery must be [rows, D], got {tuple(q.shape)} and {tuple(g.shape)}")
    k = int(k)
    _lib.require_cuda(torch)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    dist = torch.empty((q.shape[0], k), dtype=torch.float32).pin_memory()
    idx = torch.empty((q.shape[0], k), dtype=torch.int32).pin_memory()
    status = torch.empty(8, dtype=torch.int32, device=dev)
    knn_host(q, g, k, exclude_self, self_offset, device=dev, out=(dist, idx, status))
    if check:
        check_status(status)            # reads the status word back on the same stream: the host buffers are complete
    else:
        torch.cuda.current_stream(dev).synchronize()
    return dist.numpy().copy(), idx.numpy().astype(np.int64)


def retrieve_one(query, database, query_label=None, labels=None, normalize=False):
    """The reference's single-query call with the reference's return value (src/utils.py:55-81): ``dist`` [N] float32 =
    ``np.linalg.norm(query - database, axis=1)`` bit for bit (unsorted), ``idx`` [N] int64 = its argsort (ties by index;
    the reference's introsort leaves tie order unspecified) and ``ap`` = sklearn's average precision of
    ``labels == query_label`` scored by ``max(dist) - dist`` (None without labels, like the reference's ``ap = None``
    initialisation would suggest; the reference itself requires labels).  The distances come from the exact kernel in
    NumPy's summation order (csrc/sqdist.cu); the sort is a stable device sort.  For the k nearest rows of MANY queries
    use ``retrieve`` -- this per-query form exists for call-site compatibility."""
    if normalize:
        raise NotImplementedError("retrieve_one(normalize=True) is broken in the reference (undefined name, src/utils.py:69)")
    from .distance import pairwise_distance
    as_numpy = is_numpy_like(query) and is_numpy_like(database)
    q = to_cuda_f32(query).reshape(1, -1)
    g = to_cuda_f32(database, q.device)
    dist = torch.sqrt(pairwise_distance(q, g)[0])            # IEEE sqrt of the exactly summed squares == np.linalg.norm
    idx = torch.sort(dist, stable=True).indices               # ascending distance, ties by index
    ap = None
    if labels is not None:
        lab = np.squeeze(labels.cpu().numpy() if torch.is_tensor(labels) else np.asarray(labels))
        d_np = dist.cpu().numpy()
        ap = average_precision(lab == query_label, d_np.max() - d_np)       # :76-79
    if as_numpy:
        return dist.cpu().numpy(), idx.cpu().numpy(), ap
    return dist, idx, ap


def retrieve_one_topk(query, database, query_label=None, labels=None, k=None):
    """Top-k form of ``retrieve_one`` (an extension, not the reference's return value): (dist[idx[:k]], idx[:k], AP@k)
    with AP@k = ``sum_{r<=k} P(r) rel(r) / min(k, #positives)``; k defaults to min(N, 112)."""
    n = database.shape[0]
    k = min(n, _lib.KNN_MAX_K) if k is None else k
    q = np.asarray(query, dtype=np.float32).reshape(1, -1) if is_numpy_like(query) else query.reshape(1, -1)
    dist, idx = retrieve(q, database, k)
    dist, idx = dist[0], idx[0]
    ap = None
    if labels is not None:
        lab = labels.cpu().numpy() if torch.is_tensor(labels) else np.asarray(labels)
        ii = idx.cpu().numpy() if torch.is_tensor(idx) else idx
        rel = (np.squeeze(lab)[ii] == query_label)
        ap = average_precision_at_k(rel[None], np.array([(np.squeeze(lab) == query_label).sum()]))[0]
    return dist, idx, ap


def average_precision_at_k(rel: np.ndarray, n_pos: np.ndarray) -> np.ndarray:
    """AP@k = sum_{r<=k} P(r) rel(r) / min(k, #positives) for boolean ``rel`` [Q,k] (an extension: the reference
    only defines full-ranking AP).  nan where a query has no positive."""
    rel = np.asarray(rel, dtype=bool)
    k = rel.shape[1]
    cum = np.cumsum(rel, axis=1)
    prec = cum / np.arange(1, k + 1)[None, :]
    denom = np.minimum(k, np.asarray(n_pos)).astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(denom > 0, (prec * rel).sum(1) / denom, np.nan)


def recall_at_K(label_list, query_label, K=10):
    """1 iff any of the first K ranked labels equals the query label (src/utils.py:257-266)."""
    lab = label_list.cpu().numpy() if torch.is_tensor(label_list) else np.asarray(label_list)
    return 1 if np.sum(lab[:K] == query_label) > 0 else 0


def precision_at_recall(label_list, query_label, alpha=0.5):
    """Per-class fractions of the ranked prefix that reaches recall alpha (src/utils.py:231-255)."""
    lab = (label_list.cpu().numpy() if torch.is_tensor(label_list) else np.asarray(label_list)).tolist()
    query_label = int(query_label)
    target = int(alpha * sum(1 for l in lab if l == query_label))
    counts = dict.fromkeys(sorted(set(lab)), 0)
    depth = len(lab)
    for i, l in enumerate(lab):
        counts[l] += 1
        if counts[query_label] == target:
            depth = i + 1
            break
    frac = {c: n / depth for c, n in counts.items()}
    return frac[query_label], frac


def average_precision(y_true, score):
    """sklearn-compatible AP (ties grouped into one threshold); see evaluate() for the fused GPU form."""
    y = np.asarray(y_true).astype(bool).ravel()
    s = np.asarray(score).ravel()
    npos = int(y.sum())
    if npos == 0:
        return float("nan")
    order = np.argsort(-s, kind="mergesort")
    y, s = y[order], s[order]
    ends = np.r_[np.where(np.diff(s))[0], y.size - 1]
    tps = np.cumsum(y, dtype=np.float64)[ends]
    prec = tps / (ends + 1)
    rec = tps / npos
    return float(np.sum(np.diff(np.r_[0.0, rec]) * prec))


# --------------------------------------------------------------------------- leave-one-out evaluation (K5)
# Which kernels run the evaluation (tests force each one and compare the records):
#   "auto"       workspace form: tiled exact distances + shared-memory radix sort per query (N <= 24,576), else + segmented sort
#   "smem"       force the shared-memory sort path (csrc/eval_fast.cu)
#   "segmented"  force the segmented-sort path (csrc/eval_large.cu)
#   "fused"      the one-CTA-per-query kernel of round 1 (csrc/eval.cu: distances + bitonic sort fused, N <= 16385)
EVAL_PATH = "auto"
EVAL_FORCE_LARGE = False    # older switch: True == EVAL_PATH "segmented"
_EVAL_PATHS = {"auto": 0, "smem": 1, "segmented": 2}


def _loo_records(embeddings, labels, normalize, standardize, alpha, aligned, want_rank=False, queries=None, confusion=False,
                 want_hist=False):
    """Run the per-query evaluation kernels for every foreground row (or the rows in ``queries``); returns host-side
    records (one upload of the integer arrays, one read-back per result array).  ``confusion``: also accumulate the
    reference's confusion matrix on the device (``mmsim_evaluate_confusion_f32``) instead of handing back the [nq, C]
    histograms."""
    lib = _lib.load()
    lab_np = np.squeeze(labels.cpu().numpy() if torch.is_tensor(labels) else np.asarray(labels)).astype(np.int32)
    emb = to_cuda_f32(embeddings)
    if normalize:      # src/utils.py:104-105,154-155 (the reference normalises the caller's array in place)
        emb = emb / torch.linalg.vector_norm(emb, dim=1, keepdim=True)
    if standardize:    # :106-109,156-159 (float32 statistics, like NumPy on a float32 array)
        # the reference divides by (std + np.finfo(float).tiny) (:108,158): a constant column gives 0, not 0/0
        emb = (emb - emb.mean(dim=0)) / (emb.std(dim=0, unbiased=False) + float(np.finfo(float).tiny))
    emb = emb.contiguous()
    n, d = emb.shape
    dev = emb.device
    classes, cls_np = np.unique(lab_np, return_inverse=True)
    if queries is None:
        queries_np = np.nonzero(lab_np > 0)[0].astype(np.int32)      # only foreground rows are queries (:114,171)
    else:
        queries_np = np.asarray(queries, dtype=np.int32)
    nq, C = int(queries_np.size), int(classes.size)
    qcls_np = cls_np[queries_np].astype(np.int32)
    list_off = np.cumsum(np.bincount(qcls_np, minlength=C))[:C].astype(np.int32)
    list_off = np.r_[np.int32(0), list_off[:-1]].astype(np.int32)     # where each class's query list starts (confusion kernel)
    packed = torch.from_numpy(np.concatenate([lab_np, cls_np.astype(np.int32), queries_np, qcls_np, list_off,
                                              np.zeros(nq, np.int32)])).to(dev)
    lab, cls, queries, qcls = packed[:n], packed[n:2 * n], packed[2 * n:2 * n + nq], packed[2 * n + nq:2 * n + 2 * nq]
    list_off_d, lists_d = packed[2 * n + 2 * nq:2 * n + 2 * nq + C], packed[2 * n + 2 * nq + C:]
    ap = torch.empty(nq, dtype=torch.float64, device=dev)
    ints = torch.empty((4, max(nq, 1)), dtype=torch.int32, device=dev)          # npos, first, depth, hist[q, class of q]
    hist = torch.empty((max(nq, 1), C), dtype=torch.int32, device=dev)
    rank = torch.empty((nq, n - 1), dtype=torch.int32, device=dev) if want_rank else None
    path = "segmented" if EVAL_FORCE_LARGE else EVAL_PATH
    if path == "fused" and n > _lib.EVAL_SMEM_MAX_N:
        path = "auto"
    with torch.cuda.device(dev):
        if path == "fused":
            rc = lib.mmsim_evaluate_f32(emb.data_ptr(), lab.data_ptr(), cls.data_ptr(), n, d, C, queries.data_ptr(), nq,
                                        float(alpha), int(bool(aligned)), ap.data_ptr(), ints[0].data_ptr(), ints[1].data_ptr(),
                                        ints[2].data_ptr(), hist.data_ptr(), _lib.ptr(rank), stream_handle(dev))
        else:
            nbytes = ctypes.c_size_t()
            _lib.check(lib.mmsim_evaluate_large_workspace_bytes(n, nq, ctypes.byref(nbytes)), "mmsim_evaluate_large_workspace_bytes")
            ws = workspace("eval_large", nbytes.value, dev)
            rc = lib.mmsim_evaluate_ws_f32(emb.data_ptr(), lab.data_ptr(), cls.data_ptr(), n, d, C, queries.data_ptr(), nq,
                                           float(alpha), int(bool(aligned)), ap.data_ptr(), ints[0].data_ptr(),
                                           ints[1].data_ptr(), ints[2].data_ptr(), hist.data_ptr(), _lib.ptr(rank),
                                           ws.data_ptr(), ws.numel(), stream_handle(dev), _EVAL_PATHS[path])
        _lib.check(rc, "mmsim_evaluate_f32")
        rec = dict(classes=classes.tolist(), labels=lab_np, queries=queries_np, qcls=qcls_np, n=n)
        if nq:
            ints[3, :nq] = hist[:nq].gather(1, qcls.to(torch.int64)[:, None])[:, 0]
        if confusion:
            cm = torch.empty((C, C), dtype=torch.float32, device=dev)
            count = torch.empty(C, dtype=torch.int32, device=dev)
            _lib.check(lib.mmsim_evaluate_confusion_f32(hist.data_ptr(), ints[2].data_ptr(), ints[0].data_ptr(), qcls.data_ptr(),
                                                        nq, C, cm.data_ptr(), count.data_ptr(), lists_d.data_ptr(),
                                                        list_off_d.data_ptr(), stream_handle(dev)),
                       "mmsim_evaluate_confusion_f32")
            rec["cm"], rec["count"] = cm.cpu().numpy(), count.cpu().numpy()
    ints = ints.cpu().numpy()
    rec.update(ap=ap.cpu().numpy(), npos=ints[0][:nq], first=ints[1][:nq], depth=ints[2][:nq], selfhist=ints[3][:nq])
    if want_rank:
        rec["rank"] = rank
    if want_hist:
        rec["hist"] = hist.cpu().numpy()[:nq]
    return rec


def evaluate_simple(embeddings, labels, normalize=False, standardize=False, alpha=0.5, aligned=False):
    """Leave-one-out retrieval over all foreground rows -> (mAP, mPrec@alpha, R@1) (src/utils.py:83-138).

    ``aligned=False`` keeps the reference's label lookup for mPrec / R@1 (see csrc/eval_metrics.cuh); queries without any
    positive are skipped like the reference's nan-AP branch (:118-123)."""
    r = _loo_records(embeddings, labels, normalize, standardize, alpha, aligned)
    ok = r["npos"] > 0
    prec = r["selfhist"] / r["depth"]
    return np.mean(r["ap"][ok]), np.mean(prec[ok]), np.mean((r["first"][ok] < 1).astype(np.int64))


def evaluate(embeddings, labels, normalize=False, standardize=False, alpha=0.5, aligned=False):
    """Leave-one-out retrieval -> (mAP, mAP_event, mPrec, confusion, count, recall) exactly as src/utils.py:140-229:
    per-class mAP dict, confusion = {"confusion_matrix": float32 [C,C], "labels": [...]}, count int32 [C,1],
    recall = [R@1, R@2, R@4, R@8, R@16, R@32]."""
    r = _loo_records(embeddings, labels, normalize, standardize, alpha, aligned, confusion=True)
    classes, lab = r["classes"], r["labels"]
    ok = np.nonzero(r["npos"] > 0)[0]
    qcls = r["qcls"][ok]
    aps = r["ap"][ok]
    mAP = np.mean(aps)
    mPrec = np.mean(r["selfhist"][ok] / r["depth"][ok])                 # python int / int -> float64 (:252)
    # per-class mean AP, keyed in order of first appearance like the reference's dict (:206-212); np.mean over the class's
    # APs in query order == np.mean of the reference's list
    order = np.argsort(qcls, kind="stable")
    bounds = np.flatnonzero(np.r_[True, qcls[order][1:] != qcls[order][:-1], True])
    groups = sorted((order[a:b] for a, b in zip(bounds[:-1], bounds[1:])), key=lambda g: g[0])
    mAP_event = {classes[int(qcls[g[0]])]: np.mean(aps[g]) for g in groups}
    cm = r["cm"]                                                        # sequential float32 accumulation (:214-220), on the device
    count = r["count"].reshape(-1, 1).astype("int32")
    cm[1:] /= count[1:]                                                 # :222 (assumes class 0 is row 0)
    count[0] = (lab == 0).sum()                                         # :223
    confusion = {"confusion_matrix": cm, "labels": classes}
    recall = [float((r["first"][ok] < K).sum()) / len(ok) for K in RECALL_KS]
    return mAP, mAP_event, mPrec, confusion, count, recall


def full_ranking(embeddings, labels=None):
    """[N, N-1] int32 leave-one-out rankings (row-i-deleted numbering, ordered by (distance, index)) for every row:
    the argsort of the reference's retrieve_one, on the GPU."""
    n = embeddings.shape[0]
    lab = np.ones(n, np.int32) if labels is None else labels
    return _loo_records(embeddings, lab, False, False, 0.5, True, want_rank=True, queries=np.arange(n))["rank"]
