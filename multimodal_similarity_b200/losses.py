"""Fused batch-hard / lifted-structured losses -- drop-ins for src/networks.py:797-870.

Reference call sites (src/base_model_batchhard.py:115-124, src/base_model_lifted.py:115-119):

    diffs    = utils.all_diffs_tf(embedding, embedding)
    all_dist = utils.cdist_tf(diffs)
    loss, num_active, diff, weights, fp, cn = networks.batch_hard(all_dist, label_ph, margin)

The same three lines work here (``cdist_tf`` returns a lazy handle); ``batch_hard(embeddings, pids, ...)`` is the
direct form.  One cooperative CUDA launch computes the loss, the five auxiliary vectors, the mined indices and the
gradient w.r.t. the embeddings; ``loss.backward()`` only scales that gradient.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._util import stream_handle, to_cuda_f32, workspace
from .distance import LazyDists


class LossOutput(tuple):
    """The reference 6-tuple (loss, num_active, diff, weights, furthest_positive, closest_negative)
    plus ``.pos_idx`` / ``.neg_idx`` (mined column per row, -1 if none)."""
    pos_idx = None
    neg_idx = None


_ws_bytes: dict = {}


def _run(kind, emb, pids, soft, margin, weighted, need_grad):
    """One C-ABI call = one cooperative kernel launch.  Returns (scalars [2], vectors [4,N], indices [2,N], dE or None)."""
    lib = _lib.load()
    n, d = emb.shape
    dev = emb.device
    nbytes = _ws_bytes.get((n, d))
    if nbytes is None:
        c = ctypes.c_size_t()
        _lib.check(lib.mmsim_loss_workspace_bytes(n, d, ctypes.byref(c)), "mmsim_loss_workspace_bytes")
        nbytes = _ws_bytes[(n, d)] = c.value
    ws = workspace(f"loss{n}x{d}", nbytes, dev, zero=True)    # one zero-initialised workspace per layout (see mmsim.h)
    out = torch.empty((6, n + 2), dtype=torch.float32, device=dev)      # rows 0-3 vectors, 4-5 indices (int32 view), tail scalars
    vec = out[:4, :n]
    idx = out[4:6, :n].view(torch.int32)
    scal = out[0, n:n + 2]
    grad = torch.empty_like(emb) if need_grad else None
    base, pitch = out.data_ptr(), (n + 2) * 4
    if dev.index != torch.cuda.current_device():
        torch.cuda.set_device(dev)      # the library launches on the calling thread's current device
    rc = lib.mmsim_loss_f32(kind, emb.data_ptr(), pids.data_ptr(), n, d, int(soft), float(margin), int(bool(weighted)),
                            base + n * 4, base + n * 4 + 4, base, base + pitch, base + 2 * pitch, base + 3 * pitch,
                            base + 4 * pitch, base + 5 * pitch, _lib.ptr(grad), ws.data_ptr(), ws.numel(),
                            stream_handle(dev))
    _lib.check(rc, "mmsim_loss_f32")
    return scal, vec, idx, grad


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, pids, kind, soft, margin, weighted):
        scal, vec, idx, grad = _run(kind, emb, pids, soft, margin, weighted, ctx.needs_input_grad[0])
        ctx.save_for_backward(grad) if grad is not None else None
        ctx.has_grad = grad is not None
        loss, num_active = scal[0], scal[1]
        ctx.mark_non_differentiable(num_active, vec, idx)
        return loss, num_active, vec, idx

    @staticmethod
    def backward(ctx, g_loss, *_):
        if not ctx.has_grad:
            return (None,) * 6
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None, None, None, None


def _prepare(first, pids):
    """Accept embeddings [N,D] or the lazy handle of cdist_tf(all_diffs_tf(e, e))."""
    if isinstance(first, LazyDists):
        if first.metric != "squaredeuclidean":
            raise NotImplementedError("the fused losses take squared-Euclidean distances (the reference's default, "
                                      "src/base_model_batchhard.py:116); materialise other metrics with cdist()")
        d = first.diffs
        if d.a is not d.b and d.a_src is not d.b_src:
            raise ValueError("the losses need all_diffs_tf(e, e) of one batch with itself")
        src = d.a_src
        emb = src if torch.is_tensor(src) and src.is_cuda and src.dtype == torch.float32 else d.a
    else:
        emb = first
    if not (torch.is_tensor(emb) and emb.is_cuda and emb.dtype == torch.float32):
        emb = to_cuda_f32(emb)
    if emb.dim() != 2:
        raise ValueError(f"embeddings must be [N, D], got {tuple(emb.shape)}")
    pids = to_cuda_f32(pids, emb.device).reshape(-1)
    if pids.numel() != emb.shape[0]:
        raise ValueError("pids must have one label per embedding row")
    return emb, pids


def _wrap(outs) -> LossOutput:
    loss, num_active, vec, idx = outs
    res = LossOutput((loss, num_active, vec[0], vec[1], vec[2], vec[3]))
    res.pos_idx, res.neg_idx = idx[0], idx[1]
    return res


def batch_hard(dists_or_embeddings, pids, margin="soft", weighted=True) -> LossOutput:
    """Batch-hard triplet loss (src/networks.py:797-833): margin "soft" (softplus) or a float."""
    emb, pids = _prepare(dists_or_embeddings, pids)
    soft = isinstance(margin, str)
    if soft and margin != "soft":
        raise ValueError('margin must be "soft" or a number')
    return _wrap(_FusedLoss.apply(emb.contiguous(), pids, _lib.LOSS_BATCH_HARD, soft, 0.0 if soft else float(margin), weighted))


def lifted_loss(dists_or_embeddings, pids, margin, weighted=True) -> LossOutput:
    """The reference's lifted-structured variant (src/networks.py:835-870); num_active is the constant 1.0."""
    emb, pids = _prepare(dists_or_embeddings, pids)
    return _wrap(_FusedLoss.apply(emb.contiguous(), pids, _lib.LOSS_LIFTED, False, float(margin), weighted))


# --------------------------------------------------------------------------- tf.contrib metric losses (K8, K9)
class _ContribLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, labels, margin, name):
        lib = _lib.load()
        n, d = emb.shape
        dev = emb.device
        c = ctypes.c_size_t()
        _lib.check(getattr(lib, f"mmsim_{name}_workspace_bytes")(n, ctypes.byref(c)), f"mmsim_{name}_workspace_bytes")
        ws = workspace(name, c.value, dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(emb) if emb.requires_grad else None
        with torch.cuda.device(dev):
            rc = getattr(lib, f"mmsim_{name}_f32")(emb.data_ptr(), labels.data_ptr(), n, d, float(margin), loss.data_ptr(),
                                                   _lib.ptr(grad), ws.data_ptr(), ws.numel(), stream_handle(dev))
        _lib.check(rc, f"mmsim_{name}_f32")
        ctx.grad = grad
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return (None if ctx.grad is None else ctx.grad * g), None, None, None


def _contrib(name, labels, embeddings, margin):
    keep = torch.is_tensor(embeddings) and embeddings.is_cuda and embeddings.dtype == torch.float32
    emb = embeddings if keep else to_cuda_f32(embeddings)
    emb = emb if emb.is_contiguous() else emb.contiguous()
    lab = labels if torch.is_tensor(labels) else torch.as_tensor(labels)
    lab = lab.reshape(-1).to(device=emb.device).to(torch.int32).contiguous()     # class ids (the TF ops compare for equality)
    if emb.dim() != 2 or lab.numel() != emb.shape[0]:
        raise ValueError(f"{name} expects labels [N] and embeddings [N,D]; got {tuple(lab.shape)}, {tuple(emb.shape)}")
    return _ContribLoss.apply(emb, lab, float(margin), name)


def triplet_semihard_loss(labels, embeddings, margin=1.0):
    """``tf.contrib.losses.metric_learning.triplet_semihard_loss(labels, embeddings, margin)`` -- the loss the reference's
    CUB trainers select with ``--loss triplet`` (src/base_CUB.py:163-166) -- forward and backward on the device
    (csrc/semihard_loss.cu).  Same argument order as the TF function; returns a scalar CUDA tensor that supports
    ``.backward()`` when ``embeddings`` requires grad.  Parity with TF is unpinned (see the kernel's header)."""
    return _contrib("triplet_semihard", labels, embeddings, margin)


def lifted_struct_loss(labels, embeddings, margin=1.0):
    """``tf.contrib.losses.metric_learning.lifted_struct_loss(labels, embeddings, margin)`` (``--loss lifted`` of the CUB
    trainers, src/base_CUB.py:167-171), forward and backward on the device (csrc/lifted_struct.cu).  Parity unpinned."""
    return _contrib("lifted_struct", labels, embeddings, margin)
