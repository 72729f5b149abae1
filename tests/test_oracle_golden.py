"""The oracle's NumPy half against outputs of the unmodified reference (tests/golden, oracle/make_golden.py)."""
import glob
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN, golden
from oracle import retrieval_np as O


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "dist_d*.npz"))))
def test_distance_bit_exact(path):
    g = np.load(path)
    diff = O.all_diffs(g["a"], g["b"])
    for metric, key in (("squaredeuclidean", "sq"), ("euclidean", "eu"), ("l1", "l1")):
        got = O.cdist(diff, metric)
        assert got.dtype == g[key].dtype
        assert np.array_equal(got, g[key]), metric
    assert np.array_equal(O.pairwise_distance(g["a"], g["b"], chunk=5), g["sq"])


@pytest.mark.parametrize("d", [7, 32, 128, 130, 256, 1024])
def test_pairwise_sum_emulation_matches_numpy_bits(d):
    """App. A.4: the summation order the CUDA exact-distance kernels reproduce."""
    g = golden(f"dist_d{d}.npz")
    a, b = g["a"], g["b"]
    for i, j in ((0, 0), (3, 5), (36, 52), (11, 7)):
        e = O.exact_l2_emulated(a[i], b[j])
        assert e * e == pytest.approx(float(g["sq"][i, j]), rel=1e-6)
        dd = (a[i] - b[j]).astype(np.float32)
        assert O.pairwise_sum_f32(dd * dd).tobytes() == g["sq"][i, j].tobytes()
        assert e.tobytes() == np.linalg.norm(a[i:i + 1] - b[j:j + 1], axis=1)[0].tobytes()


@pytest.mark.parametrize("name", ["small", "fused", "odd"])
def test_retrieve_one(name):
    g = golden(f"retrieve_{name}.npz")
    x, lab = g["x"], g["labels"]
    for n, q in enumerate(g["queries"]):
        db, gl = np.delete(x, q, 0), np.delete(lab, q)
        ql = lab[q] if lab[q] > 0 else 1
        dist, order, ap = O.retrieve_one(x[q], db, ql, gl)
        assert np.array_equal(dist, g["dist"][n])
        # argsort is unstable: compare the sorted distance sequence, and indices outside ties
        assert np.array_equal(dist[order], g["dist"][n][g["order"][n]])
        sd = dist[order]
        untied = np.r_[True, sd[1:] != sd[:-1]] & np.r_[sd[:-1] != sd[1:], True]
        assert np.array_equal(order[untied], g["order"][n][untied])
        assert ap == pytest.approx(float(g["ap"][n]), rel=0, abs=1e-12)


def test_average_precision_matches_sklearn(rs):
    from sklearn.metrics import average_precision_score
    for _ in range(20):
        n = rs.randint(5, 200)
        y = rs.rand(n) < 0.3
        if not y.any():
            y[0] = True
        s = np.round(rs.rand(n), 1 if rs.rand() < 0.5 else 6).astype(np.float32)  # force ties half the time
        assert O.average_precision(y, s) == pytest.approx(average_precision_score(y, s), abs=1e-12)
    assert np.isnan(O.average_precision(np.zeros(5, bool), rs.rand(5)))


@pytest.mark.parametrize("name", ["hdd", "cub", "fused"])
def test_evaluate(name):
    g = golden(f"eval_{name}.npz")
    mAP, mAP_event, mPrec, confusion, count, recall = O.evaluate(g["x"], g["labels"], alpha=float(g["alpha"]))
    assert mAP == pytest.approx(float(g["mAP"]), abs=1e-12)
    assert mPrec == pytest.approx(float(g["mPrec"]), abs=1e-12)
    assert sorted(mAP_event) == g["mAP_event_keys"].tolist()
    assert np.allclose([mAP_event[k] for k in sorted(mAP_event)], g["mAP_event_vals"], atol=1e-12)
    assert confusion["labels"] == g["confusion_labels"].tolist()
    assert np.array_equal(confusion["confusion_matrix"], g["confusion"])
    assert np.array_equal(count, g["count"])
    assert np.allclose(recall, g["recall"], atol=0)
    s = O.evaluate_simple(g["x"], g["labels"], alpha=float(g["alpha"]))
    assert np.allclose(s, g["simple"], atol=1e-12)


def test_evaluate_switches():
    g = golden("eval_switches.npz")
    assert np.allclose(O.evaluate_simple(g["x"].copy(), g["labels"], normalize=True), g["normalize"], atol=1e-12)
    assert np.allclose(O.evaluate_simple(g["x"].copy(), g["labels"], standardize=True), g["standardize"], atol=1e-12)


def test_metric_helpers():
    g = golden("metrics.npz")
    n = m = 0
    for row in g["rankings"]:
        for ql in (1, 2):
            assert [O.recall_at_K(row, ql, K) for K in (1, 2, 4, 8)] == g["recall"][n].tolist()
            n += 1
            for alpha in (0.0, 0.3, 0.5, 1.0):
                p, dct = O.precision_at_recall(row, ql, alpha)
                assert p == g["prec"][m]
                ks = [k for k in g["prec_keys"][m].tolist() if k >= 0]
                assert sorted(dct) == ks
                assert [dct[k] for k in ks] == g["prec_vals"][m][:len(ks)].tolist()
                m += 1


def test_knn_matches_full_ranking(rs):
    from conftest import clustered
    x, _ = clustered(rs, 300, 64, 5)
    d, i = O.knn(x[:20], x, 10, exclude_self=True)
    for q in range(20):
        full = O.l2_to_all(x[q], x)
        full[q] = np.inf
        assert np.array_equal(d[q], np.sort(full)[:10])
        assert q not in i[q]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree only exists in the build container")
def test_live_reference_agrees(rs):
    """When the reference is mounted, run it live on fresh inputs (beyond the committed fixtures)."""
    from oracle.make_golden import load_reference_utils, clustered
    utils = load_reference_utils()
    x, lab = clustered(rs, 150, 96, 4, background=0.3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = utils.evaluate(x.copy(), lab.copy())
    got = O.evaluate(x, lab)
    assert got[0] == pytest.approx(ref[0], abs=1e-12) and got[2] == pytest.approx(ref[2], abs=1e-12)
    assert np.array_equal(got[3]["confusion_matrix"], ref[3]["confusion_matrix"])
    assert got[5] == ref[5]
