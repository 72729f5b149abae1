"""torch-CPU restatement of the reference's TensorFlow loss graph.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  TensorFlow cannot run here
and the reference ships no tests, so TensorFlow's own kernels are never
executed: this file follows src/utils.py:302-311,343-360 and
src/networks.py:797-870 op for op, in the reference's materialising form
([N,N,D] difference tensor), and is pinned by the hand-checked known-answer
tests of SURVEY.md Appendix B (tests/test_oracle_kat.py) and by the outputs of
the reference's own functions executed unmodified under oracle/tf_shim.py
(tests/golden/loss_*.npz, tests/test_oracle_losses_golden.py).  The tf.contrib
section at the end has no reference source to execute: PARITY UNPINNED.
Gradients come from torch autograd through the
same ops; ``amax``/``amin`` split the gradient evenly among ties exactly like
TF's ``_MinOrMaxGrad``.

dtype: pass float32 tensors for the parity target, float64 for the tolerance
arbiter.
"""
from __future__ import annotations

import torch


def all_diffs_tf(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """src/utils.py:302-311."""
    return a.unsqueeze(1) - b.unsqueeze(0)


def cdist_tf(diff: torch.Tensor, metric: str = "squaredeuclidean") -> torch.Tensor:
    """src/utils.py:343-360."""
    if metric == "squaredeuclidean":
        return diff.square().sum(-1)
    if metric == "euclidean":
        return (diff.square().sum(-1) + 1e-12).sqrt()
    if metric == "l1":
        return diff.abs().sum(-1)
    raise NotImplementedError(metric)


def _masks(pids: torch.Tensor):
    same = pids.unsqueeze(1) == pids.unsqueeze(0)          # networks.py:802-803
    neg = ~same                                            # :804
    pos = same ^ torch.eye(pids.shape[0], dtype=torch.bool)  # :805-806
    return same, neg, pos


def _weights(pids, neg, dtype, weighted):
    n = pids.shape[0]
    fg = (pids != 0).to(dtype)                             # networks.py:821
    if weighted:
        w = neg.to(dtype).sum(1) * fg                      # :824-825
        w = w / w.sum()                                    # :826
    else:
        w = torch.full((n,), 1.0 / n, dtype=dtype)         # :828
    return w, fg


def _softplus(x):
    return torch.clamp(x, min=0) + torch.log1p(torch.exp(-x.abs()))


def batch_hard(dists: torch.Tensor, pids: torch.Tensor, margin, weighted: bool = True):
    """src/networks.py:797-833.  Returns the reference 6-tuple
    (loss, num_active, diff, weights, furthest_positive, closest_negative).

    Empty negative set: reduce_min over an empty boolean_mask gives the reducer's
    initial value; we state it as +inf (TF>=1.13) -- the loss term is 0 either way.
    ``weighted=False`` makes the reference raise NameError at :831 (foreground_mask
    undefined); we define num_active with the same foreground mask instead.
    """
    dt = dists.dtype
    _, neg, pos = _masks(pids)
    fp = (dists * pos.to(dt)).amax(dim=1)                                  # :808
    inf = torch.full_like(dists, float("inf"))
    cn = torch.where(neg, dists, inf).amin(dim=1)                          # :809-810
    x = fp - cn                                                            # :812
    if margin == "soft":
        diff = _softplus(x)                                                # :814
    else:
        diff = torch.clamp(x + float(margin), min=0.0)                     # :816
    w, fg = _weights(pids, neg, dt, weighted)
    loss = (diff * w).sum()                                                # :830
    num_active = ((diff * fg) > 1e-5).to(dt).sum() / fg.sum()              # :831
    return loss, num_active, diff, w, fp, cn


def lifted_loss(dists: torch.Tensor, pids: torch.Tensor, margin, weighted: bool = True):
    """src/networks.py:835-870 (the author's variant).  num_active is the constant 1.0."""
    dt = dists.dtype
    _, neg, pos = _masks(pids)
    fp = torch.logsumexp(dists * pos.to(dt), dim=1)                        # :846 -- over ALL N columns
    ninf = torch.full_like(dists, float("-inf"))
    has_neg = neg.any(dim=1)
    masked = torch.where(neg, float(margin) - dists, ninf)                 # :847-848
    # rows without negatives: logsumexp(empty) = -inf and contributes no gradient
    safe = torch.where(has_neg.unsqueeze(1), masked, torch.zeros_like(masked))
    cn = torch.where(has_neg, torch.logsumexp(safe, dim=1), torch.full_like(fp, float("-inf")))
    diff = torch.clamp(torch.where(has_neg, fp + cn, torch.zeros_like(fp)), min=0.0)  # :850-852
    w, _ = _weights(pids, neg, dt, weighted)
    loss = (diff * w).sum()                                                # :866
    return loss, 1.0, diff, w, fp, cn


def mined_indices(dists: torch.Tensor, pids: torch.Tensor):
    """Hardest-positive / hardest-negative column per row, first index among ties.
    Rows without positives (negatives) get -1."""
    _, neg, pos = _masks(pids)
    dt = dists.dtype
    pv = torch.where(pos, dists, torch.full_like(dists, -1.0))
    pos_idx = pv.argmax(dim=1)
    pos_idx = torch.where(pos.any(1), pos_idx, torch.full_like(pos_idx, -1))
    nv = torch.where(neg, dists, torch.full_like(dists, float("inf")))
    neg_idx = nv.argmin(dim=1)
    neg_idx = torch.where(neg.any(1), neg_idx, torch.full_like(neg_idx, -1))
    return pos_idx, neg_idx


def loss_and_grad(kind: str, emb: torch.Tensor, pids: torch.Tensor, margin, weighted=True):
    """Forward + backward through the materialising graph, as the reference trains
    (src/base_model_batchhard.py:115-124, src/base_model_lifted.py:115-119)."""
    e = emb.detach().clone().requires_grad_(True)
    d = cdist_tf(all_diffs_tf(e, e))
    fn = batch_hard if kind == "batch_hard" else lifted_loss
    out = fn(d, pids.to(e.dtype), margin, weighted)
    out[0].backward()
    return out, e.grad.detach(), d.detach()


# --------------------------------------------------------------------------- tf.contrib metric_learning (third party)
# The reference's CUB trainers call tensorflow.contrib.losses.python.metric_learning.triplet_semihard_loss
# (src/base_CUB.py:8,163-166).  TensorFlow 1.x is neither vendored by the reference nor installable here, and the
# reference holds no test or golden vector for it: PARITY UNPINNED.  What follows restates the published TF 1.x source
# (tensorflow/contrib/losses/python/metric_learning/metric_loss_ops.py: pairwise_distance, masked_minimum,
# masked_maximum, triplet_semihard_loss) op for op in torch, so autograd gives the TF gradient as well.
def contrib_pairwise_distance(feature: torch.Tensor, squared: bool = False) -> torch.Tensor:
    sq = (feature * feature).sum(1, keepdim=True)
    d2 = sq + sq.t() - 2.0 * feature @ feature.t()
    d2 = torch.clamp(d2, min=0.0)
    err = (d2 <= 0.0).to(feature.dtype)
    d = d2 if squared else torch.sqrt(d2 + err * 1e-16)
    d = d * (1.0 - err)
    n = feature.shape[0]
    return d * (torch.ones_like(d) - torch.eye(n, dtype=d.dtype))


def contrib_masked_maximum(data, mask, dim=1):
    axis_min = data.amin(dim, keepdim=True)
    return ((data - axis_min) * mask).amax(dim, keepdim=True) + axis_min


def contrib_masked_minimum(data, mask, dim=1):
    axis_max = data.amax(dim, keepdim=True)
    return ((data - axis_max) * mask).amin(dim, keepdim=True) + axis_max


def contrib_triplet_semihard_loss(labels: torch.Tensor, embeddings: torch.Tensor, margin: float = 1.0) -> torch.Tensor:
    labels = labels.reshape(-1, 1)
    pdist = contrib_pairwise_distance(embeddings, squared=True)
    adjacency = labels == labels.t()
    adjacency_not = ~adjacency
    b = labels.numel()
    pdist_tile = pdist.repeat(b, 1)
    mask = adjacency_not.repeat(b, 1) & (pdist_tile > pdist.t().reshape(-1, 1))
    mask_final = (mask.to(pdist.dtype).sum(1, keepdim=True) > 0.0).reshape(b, b).t()
    adjacency_not_f = adjacency_not.to(pdist.dtype)
    mask_f = mask.to(pdist.dtype)
    negatives_outside = contrib_masked_minimum(pdist_tile, mask_f).reshape(b, b).t()
    negatives_inside = contrib_masked_maximum(pdist, adjacency_not_f).repeat(1, b)
    semi_hard_negatives = torch.where(mask_final, negatives_outside, negatives_inside)
    loss_mat = margin + pdist - semi_hard_negatives
    mask_positives = adjacency.to(pdist.dtype) - torch.eye(b, dtype=pdist.dtype)
    num_positives = mask_positives.sum()
    return torch.clamp(loss_mat * mask_positives, min=0.0).sum() / num_positives


def contrib_lifted_struct_loss(labels: torch.Tensor, embeddings: torch.Tensor, margin: float = 1.0) -> torch.Tensor:
    """metric_loss_ops.lifted_struct_loss (src/base_CUB.py:167-171), restated op for op from the TF 1.x source."""
    labels = labels.reshape(-1, 1)
    pairwise_distances = contrib_pairwise_distance(embeddings)
    adjacency = labels == labels.t()
    adjacency_not = ~adjacency
    b = labels.numel()
    diff = margin - pairwise_distances
    mask = adjacency_not.to(diff.dtype)
    row_minimums = diff.amin(1, keepdim=True)
    row_negative_maximums = ((diff - row_minimums) * mask).amax(1, keepdim=True) + row_minimums
    max_elements = torch.maximum(row_negative_maximums, row_negative_maximums.t())
    diff_tiled = diff.repeat(b, 1)
    mask_tiled = mask.repeat(b, 1)
    max_elements_vect = max_elements.t().reshape(-1, 1)
    loss_exp_left = (torch.exp(diff_tiled - max_elements_vect) * mask_tiled).sum(1, keepdim=True).reshape(b, b)
    loss_mat = max_elements + torch.log(loss_exp_left + loss_exp_left.t())
    loss_mat = loss_mat + pairwise_distances
    mask_positives = adjacency.to(diff.dtype) - torch.eye(b, dtype=diff.dtype)
    num_positives = mask_positives.sum() / 2.0
    return 0.25 * (torch.clamp(loss_mat * mask_positives, min=0.0) ** 2).sum() / num_positives
