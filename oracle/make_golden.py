"""Generate tests/golden/*.npz by running the UNMODIFIED reference (NumPy half).

Run in the build container only (``python -m oracle.make_golden``): it imports
``/root/reference/src/utils.py`` verbatim with ``tensorflow`` stubbed by an empty
module (the functions used here never touch tf).  The GPU box has no
/root/reference; it reads the committed fixtures.

Fixtures (all seeds fixed, float32 inputs):
  dist_*.npz      utils.cdist(utils.all_diffs(a, b)) for the three metrics, several D
  retrieve_*.npz  utils.retrieve_one: dist, argsort order, AP
  eval_*.npz      utils.evaluate / utils.evaluate_simple full return tuples
  metrics.npz     utils.recall_at_K / utils.precision_at_recall on hand-made rankings
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def load_reference_utils():
    if not os.path.isdir("/root/reference/src"):
        raise RuntimeError("reference tree not present; golden fixtures can only be regenerated in the build container")
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    sys.path.insert(0, "/root/reference/src")
    import utils  # noqa: the reference's src/utils.py
    return utils


def clustered(rs, n, d, n_classes, first_label=1, noise=0.5, background=0.0):
    """Synthetic embeddings in the shape SURVEY.md 8(d) prescribes: class centroids + noise, L2-normalised."""
    cent = rs.randn(n_classes, d).astype(np.float32)
    lab = rs.randint(0, n_classes, size=n)
    x = cent[lab] + noise * rs.randn(n, d).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    labels = (lab + first_label).astype(np.int32)
    if background > 0:
        labels[rs.rand(n) < background] = 0
    return x.astype(np.float32), labels


def main():
    utils = load_reference_utils()
    os.makedirs(OUT, exist_ok=True)
    rs = np.random.RandomState(12345)  # configs/base_config.py:15

    # ---- distances
    for d in (2, 7, 32, 128, 130, 256, 1024):
        a = rs.randn(37, d).astype(np.float32)
        b = rs.randn(53, d).astype(np.float32)
        if d == 32:  # exact duplicates -> exact zeros / ties
            b[5] = a[3]
            b[6] = b[7]
        diff = utils.all_diffs(a, b)
        np.savez_compressed(os.path.join(OUT, f"dist_d{d}.npz"), a=a, b=b,
                            sq=utils.cdist(diff), eu=utils.cdist(diff, "euclidean"), l1=utils.cdist(diff, "l1"))

    # ---- retrieve_one
    for name, n, d in (("small", 300, 128), ("fused", 2000, 256), ("odd", 257, 160)):
        x, lab = clustered(rs, n, d, 7, background=0.3 if name != "odd" else 0.0)
        qs = [0, 1, 17, n - 1]
        dist, order, ap = [], [], []
        for q in qs:
            db = np.delete(x, q, 0)
            gl = np.delete(lab, q)
            ql = lab[q] if lab[q] > 0 else 1
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                dd, oo, aa = utils.retrieve_one(x[q], db, ql, gl)
            dist.append(dd); order.append(oo); ap.append(aa)
        np.savez_compressed(os.path.join(OUT, f"retrieve_{name}.npz"), x=x, labels=lab, queries=np.array(qs),
                            dist=np.stack(dist), order=np.stack(order), ap=np.array(ap, np.float64))

    # ---- evaluate / evaluate_simple
    cases = {
        "hdd": clustered(rs, 400, 128, 6, background=0.45),          # HDD style: background class 0 present
        "cub": clustered(rs, 500, 128, 20, first_label=101),         # CUB style: labels 101.., all foreground
        "fused": None,
    }
    cam, lab = clustered(rs, 350, 128, 6, background=0.4)
    sens = clustered(rs, 350, 128, 6)[0]
    cases["fused"] = (np.concatenate((cam, sens), axis=1), lab)       # evaluate_late_fusion.py:115
    for name, (x, lab) in cases.items():
        for alpha in (0.5,):
            mAP, mAP_event, mPrec, confusion, count, recall = utils.evaluate(x.copy(), lab.copy(), alpha=alpha)
            s_mAP, s_mPrec, s_r1 = utils.evaluate_simple(x.copy(), lab.copy(), alpha=alpha)
            keys = sorted(mAP_event)
            np.savez_compressed(
                os.path.join(OUT, f"eval_{name}.npz"), x=x, labels=lab, alpha=alpha,
                mAP=np.float64(mAP), mPrec=np.float64(mPrec),
                mAP_event_keys=np.array(keys), mAP_event_vals=np.array([mAP_event[k] for k in keys], np.float64),
                confusion=confusion["confusion_matrix"], confusion_labels=np.array(confusion["labels"]),
                count=count, recall=np.array(recall, np.float64),
                simple=np.array([s_mAP, s_mPrec, s_r1], np.float64))
    # normalize / standardize switches
    x, lab = clustered(rs, 200, 64, 5, background=0.2)
    x = (x * (1 + rs.rand(200, 1))).astype(np.float32)               # not unit norm
    r_n = utils.evaluate_simple(x.copy(), lab.copy(), normalize=True)
    r_s = utils.evaluate_simple(x.copy(), lab.copy(), standardize=True)
    np.savez_compressed(os.path.join(OUT, "eval_switches.npz"), x=x, labels=lab,
                        normalize=np.array(r_n, np.float64), standardize=np.array(r_s, np.float64))

    # ---- small metric helpers
    rankings = np.array([[3, 1, 1, 2, 1, 0, 1, 2, 2, 1],
                         [1, 1, 2, 3, 0, 0, 1, 2, 3, 1],
                         [2, 2, 2, 2, 1, 3, 3, 0, 0, 2]], dtype=np.int32)
    rk, par, pad_k, pad_v = [], [], [], []
    for row in rankings:
        for ql in (1, 2):
            rk.append([utils.recall_at_K(row, ql, K) for K in (1, 2, 4, 8)])
            for alpha in (0.0, 0.3, 0.5, 1.0):
                p, dct = utils.precision_at_recall(row, ql, alpha)
                par.append(p)
                ks = sorted(dct)
                pad_k.append(ks + [-1] * (4 - len(ks)))
                pad_v.append([dct[k] for k in ks] + [0.0] * (4 - len(ks)))
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), rankings=rankings, recall=np.array(rk),
                        prec=np.array(par, np.float64), prec_keys=np.array(pad_k), prec_vals=np.array(pad_v, np.float64))
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f:24s} {os.path.getsize(os.path.join(OUT, f)) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
