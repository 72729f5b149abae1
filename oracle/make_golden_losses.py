"""Generate tests/golden/loss_*.npz by EXECUTING the reference's TensorFlow loss code, unmodified, under oracle/tf_shim.py.

Run in the build container only (``python -m oracle.make_golden_losses``): it imports ``/root/reference/src/networks.py`` and
``/root/reference/src/utils.py`` verbatim with the torch-backed shim registered as ``tensorflow`` and calls, exactly as the
trainers do (src/base_model_batchhard.py:115-124, src/base_model_lifted.py:115-119),

    dists = utils.cdist_tf(utils.all_diffs_tf(emb, emb))
    loss, num_active, diff, weights, fp, cn = networks.batch_hard(dists, pids, margin)      # or networks.lifted_loss

in float32, then ``loss.backward()`` for d loss / d emb.  The GPU box has no /root/reference; it reads the committed files.
What this does and does not pin is stated in oracle/tf_shim.py.

Fixtures (NumPy RandomState(12345), float32, SURVEY.md 8(d) shapes):
  loss_cfg1_soft / loss_cfg1_m02     batch-hard, 256 x 128, 32 classes x 8, margin "soft" / 0.2 (+ _noisy_: open hinges)
  loss_cfg1_bg_soft                  ... with class id 0 present (background rows get weight 0)
  loss_cfg2_lifted                   lifted, 512 x 128, HDD-style counts {0:200, 1:160, 2:50, 3:50, 4:25, 5:20, 6:7}, margin 1.0
  loss_lifted_unweighted             lifted, weighted=False (batch_hard with weighted=False raises NameError in the reference)
  loss_ties_*                        duplicated rows: tied hardest positives / negatives (even gradient split)
  loss_single_*                      a class with one member (no positive: fp = 0) next to ordinary classes
  loss_raw_*                         un-normalised embeddings (--no_normalized), odd width (D = 37)
  loss_kat_*                         SURVEY.md Appendix B's four points
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from oracle import tf_shim

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF_SRC = "/root/reference/src"


def clustered(rs, labels, d, noise=0.5, normalize=True):
    classes = np.unique(labels)
    cent = {c: rs.randn(d).astype(np.float32) for c in classes}
    x = np.stack([cent[c] for c in labels]) + noise * rs.randn(len(labels), d).astype(np.float32)
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def run_reference(networks, utils, kind, emb, pids, margin, weighted=True):
    e = torch.from_numpy(emb).clone().requires_grad_(True)
    p = torch.from_numpy(pids.astype(np.float32))            # label_ph is float32 (src/base_model_batchhard.py:102)
    dists = utils.cdist_tf(utils.all_diffs_tf(e, e))
    fn = networks.batch_hard if kind == "batch_hard" else networks.lifted_loss
    loss, num_active, diff, weights, fp, cn = fn(dists, p, margin, weighted)
    loss.backward()
    f = lambda t: (t.detach().numpy() if torch.is_tensor(t) else np.float32(t))
    keep = f(dists) if emb.shape[0] <= 64 else f(dists)[:8]   # the full matrix only for the small cases (file size)
    return dict(emb=emb, pids=pids.astype(np.float32), kind=kind, margin=str(margin), weighted=weighted,
                dists=keep, loss=f(loss), num_active=f(num_active), diff=f(diff), weights=f(weights),
                furthest_positive=f(fp), closest_negative=f(cn), d_emb=e.grad.numpy())


def cases(rs):
    lab1 = rs.permutation(np.repeat(np.arange(1, 33), 8))
    e1 = clustered(rs, lab1, 128)
    yield "cfg1_soft", "batch_hard", e1, lab1, "soft", True
    yield "cfg1_m02", "batch_hard", e1, lab1, 0.2, True      # (well separated clusters: every hinge is closed, loss 0)
    e1n = clustered(rs, lab1, 128, noise=1.5)                 # noisier: open hinges
    yield "cfg1_noisy_m02", "batch_hard", e1n, lab1, 0.2, True
    yield "cfg1_noisy_soft", "batch_hard", e1n, lab1, "soft", True
    lab1b = lab1.copy()
    lab1b[lab1b <= 4] = 0                                     # four classes become background
    yield "cfg1_bg_soft", "batch_hard", e1, lab1b, "soft", True
    yield "cfg1_bg_lifted", "lifted", e1, lab1b, 1.0, True
    counts = {0: 200, 1: 160, 2: 50, 3: 50, 4: 25, 5: 20, 6: 7}
    lab2 = rs.permutation(np.concatenate([np.full(n, c) for c, n in counts.items()]))
    e2 = clustered(rs, lab2, 128)
    yield "cfg2_lifted", "lifted", e2, lab2, 1.0, True
    yield "cfg2_batch_hard_soft", "batch_hard", e2, lab2, "soft", True
    lab3 = rs.permutation(np.repeat(np.arange(1, 9), 6))
    e3 = clustered(rs, lab3, 64)
    yield "lifted_unweighted", "lifted", e3, lab3, 1.0, False
    # ties: rows 1, 2 duplicate row 0's positive / negative partners
    e4 = clustered(rs, lab3, 64, noise=1.5)                   # noisy enough that hinges at margin 0.2 are open
    same = np.flatnonzero(lab3 == lab3[0])
    other = np.flatnonzero(lab3 == lab3[np.flatnonzero(lab3 != lab3[0])[0]])     # the members of one other class
    e4[same[2]] = e4[same[1]]                                 # tied positives of row 0 (and of each other's class mates)
    e4[other[1]] = e4[other[0]]                               # tied negatives for every anchor outside their class
    yield "ties_m02", "batch_hard", e4, lab3, 0.2, True
    yield "ties_soft", "batch_hard", e4, lab3, "soft", True
    yield "ties_lifted", "lifted", e4, lab3, 1.0, True
    # a class with a single member: no positives for that row
    lab5 = lab3.copy()
    lab5[0] = 99
    yield "single_soft", "batch_hard", e3, lab5, "soft", True
    yield "single_m02", "batch_hard", e3, lab5, 0.2, True
    yield "single_lifted", "lifted", e3, lab5, 1.0, True
    # un-normalised embeddings, odd width
    lab6 = rs.permutation(np.repeat(np.arange(0, 5), 7))
    e6 = clustered(rs, lab6, 37, noise=1.5, normalize=False) * 3.0
    yield "raw_soft", "batch_hard", e6, lab6, "soft", True
    yield "raw_m05", "batch_hard", e6, lab6, 0.5, True
    yield "raw_lifted", "lifted", e6, lab6, 1.0, True
    kat = np.array([[0, 0], [1, 0], [0, 2], [3, 0]], dtype=np.float32)
    yield "kat_m02", "batch_hard", kat, np.array([1, 1, 2, 2]), 0.2, True
    yield "kat_soft", "batch_hard", kat, np.array([1, 1, 2, 2]), "soft", True
    yield "kat_lifted", "lifted", kat, np.array([1, 1, 2, 2]), 1.0, True
    yield "kat_bg_soft", "batch_hard", kat, np.array([0, 1, 1, 2]), "soft", True


def main():
    if not os.path.isdir(REF_SRC):
        raise RuntimeError("reference tree not present; golden fixtures can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)                                  # one summation order, whatever the host
    sys.path.insert(0, REF_SRC)
    try:
        with tf_shim.installed():
            import networks  # noqa: the reference's src/networks.py, unmodified
            import utils     # noqa: the reference's src/utils.py, unmodified
            assert networks.__file__.startswith(REF_SRC) and utils.__file__.startswith(REF_SRC)
            rs = np.random.RandomState(12345)                 # configs/base_config.py:15
            for name, kind, emb, pids, margin, weighted in cases(rs):
                rec = run_reference(networks, utils, kind, emb, np.asarray(pids), margin, weighted)
                np.savez_compressed(os.path.join(OUT, f"loss_{name}.npz"), **rec)
                print(f"loss_{name}: {kind} N={emb.shape[0]} D={emb.shape[1]} margin={margin} loss={float(rec['loss']):.6f}")
    finally:
        sys.path.remove(REF_SRC)


if __name__ == "__main__":
    main()
