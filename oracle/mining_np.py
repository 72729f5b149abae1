"""CPU restatement of the reference's semi-hard triplet miner -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``utils.select_triplets_facenet`` (src/utils.py:430-496).  PINNED: ``oracle/make_golden_mining.py`` ran the
unmodified reference function (``np.NaN`` re-added for NumPy 2, tensorflow stubbed) with fixed ``random`` /
``np.random`` seeds and committed inputs + outputs as ``tests/golden/mining_*.npz``.
"""
from __future__ import annotations

import itertools
import random

import numpy as np


def semihard_set(all_dist, lab_int, an_idx, pos_idx, alpha):
    """The reference's ``all_neg`` (src/utils.py:473-479): ascending rows of other classes with
    pos_dist < d < pos_dist + alpha, both tests in the dtype of all_dist."""
    pos_dist = all_dist[an_idx, pos_idx]
    neg_dist = np.copy(all_dist[an_idx])
    neg_dist[lab_int == lab_int[an_idx]] = np.nan
    with np.errstate(invalid="ignore"):
        return np.where(np.logical_and(neg_dist - pos_dist < alpha, pos_dist < neg_dist))[0]


def select_triplets_facenet(lab, all_dist, triplet_per_batch, alpha=0.2, num_negative=3, cub=False):
    """``cub=True`` restates the CUB trainers' copy (src/base_model_CUB.py:25-91): label 0 also gets anchors (:50) and
    an empty result is (None, None) (:91); pinned by ``tests/golden/mining_cubcopy*.npz``."""
    lab_int = np.asarray([int(l) for l in lab])
    idx_dict = {}
    for i, l in enumerate(lab_int):                               # :444-450
        idx_dict.setdefault(int(l), []).append(i)
    for key in idx_dict:                                          # :451-452
        random.shuffle(idx_dict[key])
    iters = {key: itertools.permutations(idx_dict[key], 2) for key in idx_dict
             if cub or key != 0}                                  # :455-458
    out, counts = [], []
    while len(out) < triplet_per_batch * 3 and iters:             # :462-465
        for key in list(iters):
            try:
                an_idx, pos_idx = next(iters[key])
            except StopIteration:                                 # :470-473
                del iters[key]
                continue
            all_neg = semihard_set(all_dist, lab_int, an_idx, pos_idx, alpha)
            counts.append(len(all_neg))
            for _ in range(min(len(all_neg), num_negative)):      # :483-490
                out.extend([an_idx, pos_idx, all_neg[np.random.randint(len(all_neg))]])
                if len(out) >= triplet_per_batch * 3:
                    return out, np.mean(counts)
    if len(out) > 0:
        return out, np.mean(counts)
    return (None, None) if cub else ([], 0.)
