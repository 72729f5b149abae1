"""Pairwise distances -- drop-ins for the reference's distance ops (src/utils.py:302-360).

The reference builds ``diff = all_diffs(a, b)`` ([N1,N2,D]) and reduces it with ``cdist(diff, metric)``.
Here ``all_diffs`` returns a lazy handle (the [N1,N2,D] tensor is never built) and ``cdist`` runs one CUDA kernel
that evaluates the reference's float32 arithmetic in NumPy's summation order, so
``cdist(all_diffs(a, b), metric)`` is bit-identical to the reference's NumPy twin for all three metrics.
The ``*_tf`` names are the same objects: the TF twins compute the same values, and the lazy handle is what
``losses.batch_hard`` / ``losses.lifted_loss`` accept in place of the reference's ``dists`` tensor.
"""
from __future__ import annotations

import torch

from . import _lib
from ._util import is_numpy_like, out_like_input, stream_handle, to_cuda_f32


class Diffs:
    """Lazy ``a[:, None, :] - b[None, :, :]`` (src/utils.py:311,322)."""

    def __init__(self, a, b):
        self.as_numpy = is_numpy_like(a) and is_numpy_like(b)
        self.a_src, self.b_src = a, b          # kept so the loss wrappers can differentiate w.r.t. the embeddings
        self.a = to_cuda_f32(a)
        self.b = self.a if b is a else to_cuda_f32(b, self.a.device)
        if self.a.dim() != 2 or self.b.dim() != 2 or self.a.shape[1] != self.b.shape[1]:
            raise ValueError(f"all_diffs expects [N1,D] and [N2,D], got {tuple(self.a.shape)} and {tuple(self.b.shape)}")

    @property
    def shape(self):
        return (self.a.shape[0], self.b.shape[0], self.a.shape[1])

    def materialize(self) -> torch.Tensor:
        """The actual [N1,N2,D] tensor (debug only)."""
        return self.a[:, None, :] - self.b[None, :, :]


class LazyDists:
    """``cdist_tf(all_diffs_tf(e, e))`` not yet computed: what the fused loss kernels take instead of [N,N] dists."""

    def __init__(self, diffs: Diffs, metric: str):
        self.diffs, self.metric = diffs, metric

    def materialize(self):
        return _cdist_now(self.diffs, self.metric)


def all_diffs(a, b) -> Diffs:
    return Diffs(a, b)


def _cdist_now(diff: Diffs, metric: str):
    if metric not in _lib.METRICS:
        raise NotImplementedError(metric)   # same exception type as the reference (src/utils.py:341)
    lib = _lib.load()
    a, b = diff.a, diff.b
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = lib.mmsim_sqdist_f32(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], a.shape[1], _lib.METRICS[metric],
                                  out.data_ptr(), out.shape[1], stream_handle(a.device))
    _lib.check(rc, "mmsim_sqdist_f32")
    return out_like_input(out, diff.as_numpy)


def cdist(diff, metric: str = "squaredeuclidean"):
    """[N1,N2] distances of a lazy difference handle; metric in {squaredeuclidean, euclidean, l1} (src/utils.py:324-341)."""
    if not isinstance(diff, Diffs):
        raise TypeError("cdist expects the handle returned by all_diffs(a, b); the [N1,N2,D] tensor is never materialised")
    return _cdist_now(diff, metric)


def all_diffs_tf(a, b) -> Diffs:
    return Diffs(a, b)


def cdist_tf(diff, metric: str = "squaredeuclidean") -> LazyDists:
    """Lazy distances for the loss path (src/utils.py:343-360): pass the result to batch_hard / lifted_loss."""
    if not isinstance(diff, Diffs):
        raise TypeError("cdist_tf expects the handle returned by all_diffs_tf(a, b)")
    if metric not in _lib.METRICS:
        raise NotImplementedError(metric)
    return LazyDists(diff, metric)


def pairwise_distance(a, b, metric: str = "squaredeuclidean"):
    """cdist(all_diffs(a, b), metric) in one call."""
    return cdist(all_diffs(a, b), metric)


def project_normalize(x, W, b=None, normalized=True, epsilon=1e-10):
    """The reference's embedding head in one kernel: ``tf.nn.l2_normalize(tf.nn.xw_plus_b(x, W, b), axis=-1, epsilon)``
    (src/networks.py:376-380 + src/base_model_CUB.py:197-201); ``normalized=False`` returns the logits
    (``cfg.normalized`` off).  x [N, K], W [K, E <= 256], b [E] or None -> [N, E] float32 (NumPy in -> NumPy out)."""
    as_numpy = is_numpy_like(x)
    xd = to_cuda_f32(x)
    Wd = to_cuda_f32(W, xd.device)
    bd = None if b is None else to_cuda_f32(b, xd.device)
    if xd.dim() != 2 or Wd.dim() != 2 or xd.shape[1] != Wd.shape[0] or (bd is not None and tuple(bd.shape) != (Wd.shape[1],)):
        raise ValueError(f"project_normalize expects x [N,K], W [K,E], b [E]; got {tuple(xd.shape)}, {tuple(Wd.shape)}")
    out = torch.empty((xd.shape[0], Wd.shape[1]), dtype=torch.float32, device=xd.device)
    with torch.cuda.device(xd.device):
        rc = _lib.load().mmsim_project_normalize_f32(xd.data_ptr(), xd.shape[0], xd.shape[1], Wd.data_ptr(), _lib.ptr(bd),
                                                     Wd.shape[1], int(bool(normalized)), float(epsilon), out.data_ptr(),
                                                     stream_handle(xd.device))
    _lib.check(rc, "mmsim_project_normalize_f32")
    return out_like_input(out, as_numpy)
