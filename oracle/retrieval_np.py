"""NumPy restatement of the reference's host-side distance + retrieval evaluation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the
reference lines it restates; arithmetic that decides bits (the fp32 distance,
the score ``max(dist) - dist``) is done with the same NumPy primitives in the
same order so results are bit-identical to the reference (pinned by
tests/golden, see oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np

RECALL_KS = (1, 2, 4, 8, 16, 32)  # src/utils.py:190-197


# --------------------------------------------------------------------------- distances
def all_diffs(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """[N1,D],[N2,D] -> [N1,N2,D] broadcast difference (src/utils.py:313-322)."""
    return a[:, None, :] - b[None, :, :]


def cdist(diff: np.ndarray, metric: str = "squaredeuclidean") -> np.ndarray:
    """Reduce the last axis of a difference tensor (src/utils.py:324-341)."""
    if metric == "squaredeuclidean":
        return np.sum(np.square(diff), axis=-1)
    if metric == "euclidean":
        return np.sqrt(np.sum(np.square(diff), axis=-1) + 1e-12)
    if metric == "l1":
        return np.sum(np.abs(diff), axis=-1)
    raise NotImplementedError(metric)


def pairwise_distance(a, b, metric="squaredeuclidean", chunk=256):
    """cdist(all_diffs(a, b)) without materialising the whole [N1,N2,D] tensor."""
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.result_type(a, b))
    for s in range(0, a.shape[0], chunk):
        out[s:s + chunk] = cdist(all_diffs(a[s:s + chunk], b), metric)
    return out


def l2_to_all(query: np.ndarray, database: np.ndarray) -> np.ndarray:
    """The reference's retrieval distance, bit for bit (src/utils.py:73)."""
    return np.linalg.norm(query.reshape(1, -1) - database, axis=1)


def pairwise_sum_f32(s) -> np.float32:
    """Pure-Python emulation of NumPy's float32 pairwise add.reduce.

    Only for small known-answer tests: it documents the summation order the
    CUDA re-rank kernel reproduces.
    """
    s = [np.float32(x) for x in s]
    n = len(s)
    if n < 8:
        r = np.float32(0.0) if n == 0 else s[0]
        for x in s[1:]:
            r = np.float32(r + x)
        return r
    if n <= 128:
        r = list(s[:8])
        i = 8
        while i + 8 <= n:
            for k in range(8):
                r[k] = np.float32(r[k] + s[i + k])
            i += 8
        res = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3]))
                         + np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
        for x in s[i:]:
            res = np.float32(res + x)
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return np.float32(pairwise_sum_f32(s[:n2]) + pairwise_sum_f32(s[n2:]))


def exact_l2_emulated(q, x) -> np.float32:
    """sqrt(pairwise(fl(fl(q-x)^2))) in float32 -- scalar emulation of l2_to_all for one pair."""
    q = np.asarray(q, np.float32)
    x = np.asarray(x, np.float32)
    d = (q - x).astype(np.float32)
    return np.float32(np.sqrt(pairwise_sum_f32((d * d).astype(np.float32))))


# --------------------------------------------------------------------------- metrics
def average_precision(y_true: np.ndarray, score: np.ndarray) -> float:
    """Restatement of sklearn.metrics.average_precision_score for binary labels.

    AP = sum_n (R_n - R_{n-1}) P_n over thresholds at *distinct* score values,
    scores sorted descending with a stable mergesort (sklearn _binary_clf_curve).
    Returns nan when there is no positive (2018 sklearn behaviour the reference's
    nan-skip at src/utils.py:118-123 was written for).
    """
    y_true = np.asarray(y_true).astype(bool).ravel()
    score = np.asarray(score).ravel()
    npos = int(y_true.sum())
    if npos == 0:
        return float("nan")
    order = np.argsort(score, kind="mergesort")[::-1]
    ys = y_true[order]
    ss = score[order]
    distinct = np.where(np.diff(ss))[0]
    thr_idx = np.r_[distinct, ys.size - 1]
    tps = np.cumsum(ys, dtype=np.float64)[thr_idx]
    fps = 1 + thr_idx - tps
    precision = tps / (tps + fps)
    recall = tps / npos
    prev = np.r_[0.0, recall[:-1]]
    return float(np.sum((recall - prev) * precision))


def recall_at_K(ranked_labels: np.ndarray, query_label, K: int = 10) -> int:
    """1 iff any of the first K ranked labels matches (src/utils.py:257-266)."""
    return int(np.any(ranked_labels[:K] == query_label))


def precision_at_recall(ranked_labels: np.ndarray, query_label, alpha: float = 0.5):
    """Walk the ranking until int(alpha * #pos) positives were seen (src/utils.py:231-255).

    Returns (precision of the query class, {class: fraction of the walked prefix}).
    Quirk kept: the stop test runs after each increment on the query-class counter,
    so with int(alpha*#pos) == 0 the walk stops at rank 1 when the first item is not
    a positive and otherwise never stops (depth = N).
    """
    ranked = np.asarray(ranked_labels).tolist()
    query_label = int(query_label)
    target = int(alpha * sum(1 for l in ranked if l == query_label))
    classes = sorted(set(ranked))
    counts = {c: 0 for c in classes}
    if query_label not in counts:       # reference raises KeyError here (no positive at all)
        raise KeyError(query_label)
    depth = len(ranked)
    for i, l in enumerate(ranked):
        counts[l] += 1
        if counts[query_label] == target:
            depth = i + 1
            break
    frac = {c: counts[c] / depth for c in classes}
    return frac[query_label], frac


def retrieve_one(query, database, query_label=None, labels=None):
    """dist, ascending order, AP of one query (src/utils.py:55-81; normalize=False path)."""
    dist = l2_to_all(query, database)
    idx = np.argsort(dist)
    ap = None
    if labels is not None:
        ap = average_precision(np.squeeze(labels == query_label), np.squeeze(np.max(dist) - dist))
    return dist, idx, ap


def _prep(embeddings, normalize, standardize):
    if normalize:
        embeddings = embeddings / np.linalg.norm(embeddings, axis=1).reshape(-1, 1)
    if standardize:
        mu = np.mean(embeddings, axis=0)
        std = np.std(embeddings, axis=0) + np.finfo(float).tiny
        embeddings = (embeddings - mu) / std
    return embeddings


def _ranked_labels(labels, gl, order, aligned):
    """Labels in ranked order as the reference forms them.

    REFERENCE QUIRK (src/utils.py:128,132,185,190): ``sorted_idx`` indexes the
    gallery with row i deleted, but the reference looks the labels up in the FULL
    ``labels`` array, so every ranked item at original position >= i gets the label
    of its predecessor.  mAP is unaffected (computed inside retrieve_one with the
    deleted labels); mPrec, the confusion matrix and recall@K are.  ``aligned=False``
    reproduces the reference; ``aligned=True`` is the intended behaviour.
    """
    return gl[order] if aligned else labels[order]


def evaluate_simple(embeddings, labels, normalize=False, standardize=False, alpha=0.5, aligned=False):
    """Leave-one-out retrieval -> (mAP, mPrec@alpha, R@1) (src/utils.py:83-138)."""
    embeddings = _prep(embeddings, normalize, standardize)
    labels = np.squeeze(labels)
    aps, precs, hits = [], [], []
    for i in range(embeddings.shape[0]):
        if labels[i] <= 0:
            continue
        gl = np.delete(labels, i)
        _, order, ap = retrieve_one(embeddings[i], np.delete(embeddings, i, 0), labels[i], gl)
        if np.isnan(ap):
            continue
        ranked = _ranked_labels(labels, gl, order, aligned)
        aps.append(ap)
        precs.append(precision_at_recall(ranked, labels[i], alpha)[0])
        hits.append(recall_at_K(ranked, labels[i], 1))
    return np.mean(aps), np.mean(precs), np.mean(hits)


def evaluate(embeddings, labels, normalize=False, standardize=False, alpha=0.5, aligned=False):
    """Leave-one-out retrieval -> (mAP, mAP_event, mPrec, confusion, count, recall[6]) (src/utils.py:140-229)."""
    embeddings = _prep(embeddings, normalize, standardize)
    labels = np.squeeze(labels)
    classes = sorted(set(labels.tolist()))
    aps, qlab, precs, confs = [], [], [], []
    hits = [0] * len(RECALL_KS)
    for i in range(embeddings.shape[0]):
        if labels[i] <= 0:
            continue
        gl = np.delete(labels, i)
        _, order, ap = retrieve_one(embeddings[i], np.delete(embeddings, i, 0), labels[i], gl)
        if np.isnan(ap):
            continue
        ranked = _ranked_labels(labels, gl, order, aligned)
        aps.append(ap)
        qlab.append(int(labels[i]))
        p, conf = precision_at_recall(ranked, labels[i], alpha)
        precs.append(p)
        confs.append(conf)
        for n, K in enumerate(RECALL_KS):
            hits[n] += recall_at_K(ranked, labels[i], K)
    mAP = np.mean(aps)
    mPrec = np.mean(precs)
    per_class = {}
    for ap, l in zip(aps, qlab):
        per_class.setdefault(l, []).append(ap)
    mAP_event = {l: np.mean(v) for l, v in per_class.items()}
    cm = np.zeros((len(classes), len(classes)), dtype="float32")
    count = np.zeros((len(classes), 1), dtype="int32")
    for conf, l in zip(confs, qlab):
        r = classes.index(l)
        for c, v in conf.items():
            cm[r, classes.index(c)] += v
        count[r] += 1
    cm[1:] /= count[1:]                     # reference assumes class 0 is row 0 (src/utils.py:222)
    count[0] = (labels == 0).sum()
    confusion = {"confusion_matrix": cm, "labels": classes}
    recall = [float(h) / len(qlab) for h in hits]
    return mAP, mAP_event, mPrec, confusion, count, recall


# --------------------------------------------------------------------------- kNN (the sharded path's oracle)
def knn(queries, gallery, k, exclude_self=False, self_offset=0):
    """Exact top-k by the reference distance, ties broken by smaller index.

    Returns (dist [Q,k] float32, idx [Q,k] int64).  With exclude_self, gallery row
    ``self_offset + i`` is removed for query i and indices stay in the *full*
    gallery numbering (the reference's np.delete numbering is idx - (idx > i)).
    """
    Q = queries.shape[0]
    out_d = np.empty((Q, k), np.float32)
    out_i = np.empty((Q, k), np.int64)
    for i in range(Q):
        d = l2_to_all(queries[i], gallery)
        if exclude_self:
            d = d.copy()
            d[self_offset + i] = np.inf
        order = np.lexsort((np.arange(d.size), d))[:k]
        out_d[i] = d[order]
        out_i[i] = order
    return out_d, out_i


def late_fusion(a, b):
    """Feature concatenation of two modalities (src/evaluate_late_fusion.py:115)."""
    return np.concatenate((a, b), axis=1)
