"""K5: utils.evaluate / utils.evaluate_simple on the GPU against golden outputs of the unmodified reference."""
import numpy as np
import pytest

from conftest import clustered, golden
from oracle import retrieval_np as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["hdd", "cub", "fused"])
def test_evaluate_golden(name):
    import multimodal_similarity_b200 as mm
    g = golden(f"eval_{name}.npz")
    mAP, mAP_event, mPrec, confusion, count, recall = mm.evaluate(g["x"], g["labels"], alpha=float(g["alpha"]))
    assert mAP == pytest.approx(float(g["mAP"]), abs=1e-12)
    assert mPrec == pytest.approx(float(g["mPrec"]), abs=1e-12)
    assert sorted(mAP_event) == g["mAP_event_keys"].tolist()
    assert np.allclose([mAP_event[k] for k in sorted(mAP_event)], g["mAP_event_vals"], atol=1e-12)
    assert confusion["labels"] == g["confusion_labels"].tolist()
    assert np.array_equal(confusion["confusion_matrix"], g["confusion"])        # float32, bit for bit
    assert np.array_equal(count, g["count"]) and count.dtype == np.int32
    assert recall == g["recall"].tolist()
    s = mm.evaluate_simple(g["x"], g["labels"], alpha=float(g["alpha"]))
    assert np.allclose(s, g["simple"], atol=1e-12)


def test_evaluate_switches_golden():
    import multimodal_similarity_b200 as mm
    g = golden("eval_switches.npz")
    # normalisation / standardisation run in torch on the GPU: distances can differ in the last bit from NumPy's,
    # so these compare at metric level
    assert np.allclose(mm.evaluate_simple(g["x"].copy(), g["labels"], normalize=True), g["normalize"], atol=2e-3)
    assert np.allclose(mm.evaluate_simple(g["x"].copy(), g["labels"], standardize=True), g["standardize"], atol=2e-3)


@pytest.mark.parametrize("n,d,c,bg,alpha", [(64, 16, 3, 0.3, 0.5), (513, 128, 11, 0.0, 0.3), (1025, 96, 6, 0.5, 1.0), (300, 256, 40, 0.1, 0.0)])
def test_evaluate_vs_oracle(n, d, c, bg, alpha, rs):
    import multimodal_similarity_b200 as mm
    x, lab = clustered(rs, n, d, c, background=bg)
    for aligned in (False, True):
        ref = O.evaluate(x, lab, alpha=alpha, aligned=aligned)
        got = mm.evaluate(x, lab, alpha=alpha, aligned=aligned)
        assert got[0] == pytest.approx(ref[0], abs=1e-12) and got[2] == pytest.approx(ref[2], abs=1e-12)
        assert got[1].keys() == ref[1].keys()
        # (assert_array_equal: a class that never occurs gives the reference's 0/0 = NaN row, equal NaNs are parity)
        np.testing.assert_array_equal(got[3]["confusion_matrix"], ref[3]["confusion_matrix"])
        assert np.array_equal(got[4], ref[4])
        assert got[5] == ref[5]
    # exact duplicates: distance / score ties.  AP groups ties into one threshold so it is order independent; the
    # rank-based metrics depend on the reference's unstable argsort inside a tie and are not compared here.
    x[7] = x[3]
    x[11] = x[3]
    ref = O.evaluate(x, lab, alpha=alpha)
    got = mm.evaluate(x, lab, alpha=alpha)
    assert got[0] == pytest.approx(ref[0], abs=1e-12)
    assert np.allclose([got[1][k] for k in sorted(got[1])], [ref[1][k] for k in sorted(ref[1])], atol=1e-12)


def test_full_ranking_matches_reference_argsort(rs):
    import multimodal_similarity_b200 as mm
    x, _ = clustered(rs, 400, 128, 5)
    rank = mm.full_ranking(x).cpu().numpy()
    for i in (0, 1, 200, 399):
        dist = O.l2_to_all(x[i], np.delete(x, i, 0))
        assert np.array_equal(dist[rank[i]], np.sort(dist))
        assert np.array_equal(rank[i], np.lexsort((np.arange(399), dist)))


def test_cub_shape_recall(rs):
    """BASELINE config 3 in miniature: GoogleNet-shaped features projected to 128-d, labels 101.., all foreground."""
    import multimodal_similarity_b200 as mm
    feats = rs.randn(20, 1024)[rs.randint(0, 20, 1200)] + rs.randn(1200, 1024)
    emb = (feats @ (rs.randn(1024, 128) / 32)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    lab = (rs.randint(0, 20, 1200) + 101).astype(np.int32)
    ref = O.evaluate(emb, lab)
    got = mm.evaluate(emb, lab)
    assert got[0] == pytest.approx(ref[0], abs=1e-12) and got[5] == ref[5]


# ------------------------------------------------------------------------- the three kernel paths of the evaluation
@pytest.fixture(params=["smem", "segmented", "fused"])
def eval_path(request):
    from multimodal_similarity_b200 import retrieval
    retrieval.EVAL_PATH = request.param
    yield request.param
    retrieval.EVAL_PATH = "auto"


@pytest.mark.parametrize("name", ["hdd", "cub", "fused"])
def test_every_path_golden(name, eval_path):
    """Tiled distances + shared-memory radix sort (csrc/eval_fast.cu), + segmented sort (csrc/eval_large.cu) and the fused
    one-CTA-per-query kernel (csrc/eval.cu), each against the reference's own outputs."""
    import multimodal_similarity_b200 as mm
    g = golden(f"eval_{name}.npz")
    got = mm.evaluate(g["x"].copy(), g["labels"].copy(), alpha=float(g["alpha"]))
    assert got[0] == pytest.approx(float(g["mAP"]), abs=1e-12) and got[2] == pytest.approx(float(g["mPrec"]), abs=1e-12)
    np.testing.assert_array_equal(got[3]["confusion_matrix"], g["confusion"])
    assert np.array_equal(got[4], g["count"]) and got[5] == g["recall"].tolist()
    s = mm.evaluate_simple(g["x"].copy(), g["labels"].copy(), alpha=float(g["alpha"]))
    assert np.allclose(s, g["simple"], atol=1e-12)


def _records(path, x, lab, alpha, aligned, **kw):
    from multimodal_similarity_b200 import retrieval
    retrieval.EVAL_PATH = path
    try:
        return retrieval._loo_records(x, lab, False, False, alpha, aligned, want_rank=True, want_hist=True, **kw)
    finally:
        retrieval.EVAL_PATH = "auto"


# sizes on both sides of every launch shape of the sort kernel (512, 2048, 4096, 6144 items) and feature widths with one
# summation leaf (16, 100: 8-wide body + tail), two (160, 256), a deeper tree (250) and the sequential form (7)
@pytest.mark.parametrize("n,d,c,bg,alpha", [(64, 16, 3, 0.3, 0.5), (1300, 96, 6, 0.5, 1.0), (300, 256, 40, 0.1, 0.0), (512, 100, 5, 0.2, 0.5),
                                            (513, 160, 9, 0.0, 0.7), (2049, 250, 4, 0.1, 0.5), (4097, 7, 12, 0.0, 0.5),
                                            (6145, 32, 30, 0.3, 0.25)])
def test_paths_agree_record_by_record(n, d, c, bg, alpha, rs):
    x, lab = clustered(rs, n, d, c, background=bg)
    x[7] = x[3]; x[11] = x[3]                       # exact ties: every path orders by (distance, index)
    qs = None if n <= 1300 else np.nonzero(lab > 0)[0][:: max(1, n // 200)]
    for aligned in (False, True):
        ref = _records("fused", x, lab, alpha, aligned, queries=qs)
        for path in ("smem", "segmented"):
            got = _records(path, x, lab, alpha, aligned, queries=qs)
            for key in ("npos", "first", "depth", "hist", "selfhist"):
                assert np.array_equal(got[key], ref[key]), (path, key)
            assert np.allclose(got["ap"], ref["ap"], rtol=0, atol=1e-12)
            assert np.array_equal(got["rank"].cpu().numpy(), ref["rank"].cpu().numpy())


def test_tiled_distances_are_numpy_bits(rs):
    """The register-tiled distance kernel feeds both sort paths: the ranking's distances must be the reference's float32
    np.linalg.norm bits for every summation shape (one leaf, leaf + tail, two leaves, deeper trees)."""
    for n, d in ((700, 128), (333, 100), (257, 160), (130, 256), (90, 250), (65, 8), (50, 1024)):
        x, lab = clustered(rs, n, d, 4)
        rec = _records("smem", x, lab, 0.5, True, queries=[0, n // 2, n - 1])
        rank = rec["rank"].cpu().numpy()
        for n_, i in enumerate((0, n // 2, n - 1)):
            dist = O.l2_to_all(x[i], np.delete(x, i, 0))
            assert np.array_equal(rank[n_], np.lexsort((np.arange(n - 1), dist))), (n, d, i)


@pytest.mark.parametrize("path", ["auto", "segmented"])
def test_large_gallery_vs_oracle(rs, path):
    """BASELINE config 4's leave-one-out form: 20,000 fused 2 x 128-d rows, HDD-style labels (7 classes: thousands of
    positives per query).  A handful of queries against the oracle's per-query arithmetic."""
    cam, lab = clustered(rs, 20000, 128, 7, first_label=0)          # label 0 = background
    sens, _ = clustered(rs, 20000, 128, 7)
    x = np.concatenate((cam, sens), axis=1)
    qs = [int(q) for q in np.nonzero(lab > 0)[0][[0, 1, 777, 5000, -1]]]
    for aligned in (False, True):
        rec = _records(path, x, lab, 0.5, aligned, queries=qs)
        for n_, i in enumerate(qs):
            gl = np.delete(lab, i)
            dist, order, ap = O.retrieve_one(x[i], np.delete(x, i, 0), lab[i], gl)
            ranked = O._ranked_labels(lab, gl, order, aligned)
            assert rec["ap"][n_] == pytest.approx(ap, abs=1e-12)
            first = int(np.argmax(ranked == lab[i])) if (ranked == lab[i]).any() else len(ranked)
            assert rec["first"][n_] == first
            p, conf = O.precision_at_recall(ranked, lab[i], 0.5)
            depth = rec["depth"][n_]
            assert rec["hist"][n_][rec["classes"].index(int(lab[i]))] / depth == pytest.approx(p, abs=1e-12)
            for c, v in conf.items():
                assert rec["hist"][n_][rec["classes"].index(int(c))] / depth == pytest.approx(v, abs=1e-12)


def test_beyond_the_shared_memory_sort(rs):
    """N = 30,000 > 24,576: the workspace form falls through to the segmented sort on its own."""
    x, lab = clustered(rs, 30000, 32, 50)
    qs = [0, 12345, 29999]
    rec = _records("auto", x, lab, 0.5, True, queries=qs)
    rank = rec["rank"].cpu().numpy()
    for n_, i in enumerate(qs):
        gl = np.delete(lab, i)
        dist, order, ap = O.retrieve_one(x[i], np.delete(x, i, 0), lab[i], gl)
        assert np.array_equal(rank[n_], np.lexsort((np.arange(29999), dist)))
        assert rec["ap"][n_] == pytest.approx(ap, abs=1e-12)


def test_evaluate_host_assembly_matches_oracle_on_few_classes(rs):
    """Few classes, many queries per class: the device-side confusion accumulation adds thousands of float32 rows per class
    in query order -- bit-identical to the reference's python loop."""
    import multimodal_similarity_b200 as mm
    x, lab = clustered(rs, 3000, 64, 4, background=0.2)
    ref = O.evaluate(x, lab, alpha=0.5)
    got = mm.evaluate(x, lab, alpha=0.5)
    assert got[0] == pytest.approx(ref[0], abs=1e-12) and got[2] == pytest.approx(ref[2], abs=1e-12)
    assert list(got[1].keys()) == list(ref[1].keys())
    assert np.allclose(list(got[1].values()), list(ref[1].values()), rtol=0, atol=1e-12)
    np.testing.assert_array_equal(got[3]["confusion_matrix"], ref[3]["confusion_matrix"])
    assert np.array_equal(got[4], ref[4]) and got[5] == ref[5]
