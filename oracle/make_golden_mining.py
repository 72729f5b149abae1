"""Generate tests/golden/mining_*.npz by running the UNMODIFIED reference miner.

Build container only (``python -m oracle.make_golden_mining``): imports /root/reference/src/utils.py verbatim
(tensorflow stubbed; ``np.NaN``, removed in NumPy 2, re-added as an alias of ``np.nan`` -- src/utils.py:477 uses it).
Each fixture stores the inputs, the seeds given to ``random`` and ``np.random``, and the function's return value.

``mining_cubcopy*.npz`` come from the CUB trainers' copy of the miner (src/base_model_CUB.py:25-91).  That module
imports tensorflow.contrib at the top, so the one function definition is parsed out of the file where it lies (``ast``)
and executed unmodified; nothing of it is stored in this repo.
"""
from __future__ import annotations

import ast
import itertools
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


CUB_CASES = {
    # same tuple as CASES; "cubcopy" has 40% of the rows labelled 0, which this copy mines like any other class
    "cubcopy": (280, 64, 7, 0.4, 0.8, 400, 0.25, 3, "squaredeuclidean"),
    "cubcopy_none": (60, 16, 3, 0.0, 0.0, 10, 0.2, 3, "squaredeuclidean"),
}


def load_cub_miner(path="/root/reference/src/base_model_CUB.py", name="select_triplets_facenet"):
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    fn = [node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name == name]
    ns = {"np": np, "random": random, "itertools": itertools}
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)
    return ns[name]


def run_cases(cases, utils, miner, rs):
    for name, (n, d, c, bg, noise, tpb, alpha, nneg, metric) in cases.items():
        x, lab = clustered(rs, n, d, c, noise=noise, background=bg)
        if name.endswith("none"):
            lab = (np.arange(n) % c + 1).astype(np.int32)
            x = np.eye(c, d, dtype=np.float32)[lab - 1]
        dist = utils.cdist(utils.all_diffs(x, x), metric=metric)
        seed = 1000 + len(name)
        random.seed(seed)
        np.random.seed(seed)
        with np.errstate(invalid="ignore"):
            trip, active = miner(lab, dist, tpb, alpha=alpha, num_negative=nneg)
        empty_is_none = trip is None
        np.savez_compressed(os.path.join(OUT, f"mining_{name}.npz"), x=x, labels=lab, seed=seed, metric=metric,
                            triplet_per_batch=tpb, alpha=alpha, num_negative=nneg, empty_is_none=empty_is_none,
                            triplets=np.asarray([] if trip is None else trip, dtype=np.int64),
                            active=np.float64(0. if active is None else active))
        print(name, 0 if trip is None else len(trip) // 3, "triplets, mean semi-hard count", active)


def main():
    utils = load_reference_utils()
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    run_cases(CASES, utils, utils.select_triplets_facenet, np.random.RandomState(4321))
    run_cases(CUB_CASES, utils, load_cub_miner(), np.random.RandomState(8765))


if __name__ == "__main__":
    main()
