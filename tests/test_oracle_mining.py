"""The oracle's semi-hard miner against outputs of the unmodified reference (oracle/make_golden_mining.py)."""
import glob
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mining_np as M
from oracle import retrieval_np as O

CASES = sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN, "mining_*.npz")))


def load_case(name):
    g = np.load(os.path.join(GOLDEN, f"mining_{name}.npz"))
    dist = O.cdist(O.all_diffs(g["x"], g["x"]), str(g["metric"]))
    return g, dist


def test_fixtures_present():
    assert set(CASES) >= {"hdd", "cub", "euclid", "starved", "none", "cubcopy", "cubcopy_none"}


@pytest.mark.parametrize("name", CASES)
def test_select_triplets_facenet_matches_reference(name):
    g, dist = load_case(name)
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    trip, active = M.select_triplets_facenet(g["labels"], dist, int(g["triplet_per_batch"]), alpha=float(g["alpha"]),
                                             num_negative=int(g["num_negative"]), cub=name.startswith("cubcopy"))
    if bool(g["empty_is_none"]):                                  # src/base_model_CUB.py:91
        assert trip is None and active is None
        return
    assert np.array_equal(np.asarray(trip, dtype=np.int64), g["triplets"])
    assert float(active) == float(g["active"])
    if name == "cubcopy":                                         # label 0 is mined too in this copy (:50)
        assert (g["labels"][np.asarray(trip[0::3])] == 0).any()


def test_live_reference_when_mounted():
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not mounted (GPU box)")
    from oracle.make_golden import clustered, load_reference_utils
    utils = load_reference_utils()
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    rs = np.random.RandomState(7)
    x, lab = clustered(rs, 150, 32, 5, noise=0.8, background=0.25)
    dist = utils.cdist(utils.all_diffs(x, x))
    for seed in (0, 1):
        random.seed(seed); np.random.seed(seed)
        with np.errstate(invalid="ignore"):
            want = utils.select_triplets_facenet(lab, dist, 90, alpha=0.15, num_negative=2)
        random.seed(seed); np.random.seed(seed)
        got = M.select_triplets_facenet(lab, dist, 90, alpha=0.15, num_negative=2)
        assert [int(v) for v in got[0]] == [int(v) for v in want[0]] and got[1] == want[1]
