"""Generate tests/golden/mining_*.npz by running the UNMODIFIED reference miner.

Build container only (``python -m oracle.make_golden_mining``): imports /root/reference/src/utils.py verbatim
(tensorflow stubbed; ``np.NaN``, removed in NumPy 2, re-added as an alias of ``np.nan`` -- src/utils.py:477 uses it).
Each fixture stores the inputs, the seeds given to ``random`` and ``np.random``, and the function's return value.
"""
from __future__ import annotations

import os
import random

import numpy as np

from .make_golden import OUT, clustered, load_reference_utils

CASES = {
    # name: (n, d, classes, background fraction, noise, triplet_per_batch, alpha, num_negative, metric)
    "hdd": (300, 128, 6, 0.4, 0.5, 200, 0.2, 3, "squaredeuclidean"),
    "cub": (333, 128, 40, 0.0, 2.0, 3000, 0.3, 2, "squaredeuclidean"),     # many classes, wide semi-hard bands
    "euclid": (257, 64, 5, 0.2, 0.9, 64, 0.1, 1, "euclidean"),
    "starved": (120, 32, 4, 0.3, 0.05, 5000, 0.01, 3, "squaredeuclidean"),  # tight clusters: pairs run out before the batch is full
    "none": (60, 16, 3, 0.0, 0.0, 10, 0.2, 3, "squaredeuclidean"),         # zero noise: no semi-hard negative at all
}


def main():
    utils = load_reference_utils()
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    rs = np.random.RandomState(4321)
    for name, (n, d, c, bg, noise, tpb, alpha, nneg, metric) in CASES.items():
        x, lab = clustered(rs, n, d, c, noise=noise, background=bg)
        if name == "none":
            lab = (np.arange(n) % c + 1).astype(np.int32)
            x = np.eye(c, d, dtype=np.float32)[lab - 1]
        dist = utils.cdist(utils.all_diffs(x, x), metric=metric)
        seed = 1000 + len(name)
        random.seed(seed)
        np.random.seed(seed)
        with np.errstate(invalid="ignore"):
            trip, active = utils.select_triplets_facenet(lab, dist, tpb, alpha=alpha, num_negative=nneg)
        np.savez_compressed(os.path.join(OUT, f"mining_{name}.npz"), x=x, labels=lab, seed=seed, metric=metric,
                            triplet_per_batch=tpb, alpha=alpha, num_negative=nneg,
                            triplets=np.asarray(trip, dtype=np.int64), active=np.float64(active))
        print(name, len(trip) // 3, "triplets, mean semi-hard count", active)


if __name__ == "__main__":
    main()
