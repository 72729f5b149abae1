"""Semi-hard (FaceNet) triplet mining -- drop-in for ``utils.select_triplets_facenet`` (src/utils.py:430-496).

The reference walks (anchor, positive) pairs class by class (python ``random.shuffle`` + ``itertools.permutations``,
round robin over the foreground classes), tests every other row of the distance matrix for the semi-hard condition in
NumPy and draws up to ``num_negative`` negatives per pair with ``np.random.randint``.  Here the pair order and the
draws stay on the host -- they consume the two global RNGs exactly like the reference, so the same seeds give the same
triplets -- while the O(pairs x N) work runs on the device: one launch counts the semi-hard negatives of a chunk of
pairs (``mmsim_semihard_mask_f32``; only the counts come back, the draws depend on nothing else) and one launch
resolves the draws to row indices (``mmsim_semihard_pick_f32``).  The distance matrix never leaves the device when it
comes from ``pairwise_distance`` on CUDA tensors.
"""
from __future__ import annotations

import itertools
import random

import numpy as np
import torch

from . import _lib
from ._util import stream_handle, to_cuda_f32


def _pair_stream(lab, background=0):
    """(anchor, positive) pairs in the reference's visiting order (src/utils.py:445-472).  ``background`` is the label
    that gets no anchors (0 in utils.py:455); None gives every class anchors (src/base_model_CUB.py:50)."""
    idx_dict: dict = {}
    for i, l in enumerate(lab):
        idx_dict.setdefault(int(l), []).append(i)
    for key in idx_dict:
        random.shuffle(idx_dict[key])
    iters = {key: itertools.permutations(idx_dict[key], 2) for key in idx_dict
             if background is None or not key == background}
    while iters:
        for key in list(iters):
            try:
                yield next(iters[key])
            except StopIteration:
                del iters[key]


def semihard_counts(all_dist, lab, pairs, alpha=0.2, return_mask=False):
    """``len(all_neg)`` (src/utils.py:477-479) for an int [m,2] array of (anchor, positive) pairs; device tensors out."""
    dist = to_cuda_f32(all_dist)
    dev = dist.device
    n = dist.shape[0]
    if dist.dim() != 2 or dist.shape[1] != n:
        raise ValueError(f"all_dist must be [N,N], got {tuple(dist.shape)}")
    labels = torch.as_tensor(np.asarray([int(l) for l in lab], dtype=np.int32)).to(dev) if not torch.is_tensor(lab) \
        else lab.to(device=dev).to(torch.int32).contiguous()
    pairs = torch.as_tensor(np.ascontiguousarray(pairs, dtype=np.int32)).reshape(-1, 2).to(dev)
    m = pairs.shape[0]
    count = torch.empty(m, dtype=torch.int32, device=dev)
    mask = torch.empty((m, (n + 31) // 32), dtype=torch.int32, device=dev) if return_mask else None
    with torch.cuda.device(dev):
        rc = _lib.load().mmsim_semihard_mask_f32(dist.data_ptr(), n, dist.stride(0), labels.data_ptr(), pairs.data_ptr(), m,
                                                 float(alpha), _lib.ptr(mask), count.data_ptr(), stream_handle(dev))
    _lib.check(rc, "mmsim_semihard_mask_f32")
    return (count, mask) if return_mask else count


def select_triplets_facenet(lab, all_dist, triplet_per_batch, alpha=0.2, num_negative=3, *, background=0,
                            empty=([], 0.)):
    """Reference signature and return value: ``(triplet_input_idx, mean semi-hard count)`` -- a flat list
    ``[anchor, positive, negative, ...]`` of at most ``3 * triplet_per_batch`` ints, ``([], 0.)`` when nothing is found.
    Same triplets as the reference for the same ``random`` / ``np.random`` seeds.

    The keyword-only arguments select the reference's near-copies: ``background=None, empty=(None, None)`` is the
    CUB trainers' version (``select_triplets_facenet_cub`` below)."""
    dist = to_cuda_f32(all_dist)
    dev = dist.device
    n = dist.shape[0]
    if dist.dim() != 2 or dist.shape[1] != n or len(lab) != n:
        raise ValueError(f"all_dist must be [N,N] with N = len(lab), got {tuple(dist.shape)} and {len(lab)}")
    lib = _lib.load()
    labels = torch.as_tensor(np.asarray([int(l) for l in lab], dtype=np.int32)).to(dev)
    want = int(triplet_per_batch) * 3
    stream = _pair_stream(lab, background)          # shuffles now, like the reference, even if nothing is asked for
    picks: list = []                    # (anchor, positive, r)
    all_neg_count: list = []
    chunk = max(256, 2 * int(triplet_per_batch))
    done = want <= 0
    while not done:
        pairs = list(itertools.islice(stream, chunk))
        if not pairs:
            break
        chunk *= 2
        pairs_d = torch.as_tensor(np.asarray(pairs, dtype=np.int32)).to(dev)
        count_d = torch.empty(len(pairs), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.mmsim_semihard_mask_f32(dist.data_ptr(), n, dist.stride(0), labels.data_ptr(), pairs_d.data_ptr(),
                                             len(pairs), float(alpha), 0, count_d.data_ptr(), stream_handle(dev))
        _lib.check(rc, "mmsim_semihard_mask_f32")
        for (an, pos), c in zip(pairs, count_d.cpu().tolist()):
            all_neg_count.append(c)
            for _ in range(min(c, num_negative)):
                picks.append((an, pos, np.random.randint(c)))
                if len(picks) * 3 >= want:
                    done = True
                    break
            if done:
                break
    if not picks:
        return tuple(list(e) if isinstance(e, list) else e for e in empty)
    picks_d = torch.as_tensor(np.asarray(picks, dtype=np.int32)).to(dev)
    neg_d = torch.empty(len(picks), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mmsim_semihard_pick_f32(dist.data_ptr(), n, dist.stride(0), labels.data_ptr(), picks_d.data_ptr(), len(picks),
                                         float(alpha), neg_d.data_ptr(), stream_handle(dev))
    _lib.check(rc, "mmsim_semihard_pick_f32")
    neg = neg_d.cpu().tolist()
    if min(neg) < 0:
        raise _lib.MmsimError("semi-hard pick out of range (count and pick kernels disagree)")
    triplet_input_idx = []
    for (an, pos, _), ng in zip(picks, neg):
        triplet_input_idx.extend([an, pos, ng])
    return triplet_input_idx, float(np.mean(all_neg_count))


def select_triplets_facenet_cub(lab, all_dist, triplet_per_batch, alpha=0.2, num_negative=3):
    """The CUB trainers' copy (src/base_model_CUB.py:25-91, debug_CUB.py:22-88, pddm_CUB.py:26-92): label 0 is an
    ordinary class and an empty result is ``(None, None)``; otherwise identical to ``select_triplets_facenet``."""
    return select_triplets_facenet(lab, all_dist, triplet_per_batch, alpha, num_negative, background=None,
                                   empty=(None, None))
