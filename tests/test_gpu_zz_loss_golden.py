"""K2/K3 against the reference ITSELF: the fused batch-hard / lifted kernels compared with tests/golden/loss_*.npz -- the outputs
of the reference's unmodified ``networks.batch_hard`` / ``networks.lifted_loss`` over ``utils.cdist_tf(utils.all_diffs_tf(e, e))``
executed under the torch-backed ``tensorflow`` stand-in (oracle/tf_shim.py, oracle/make_golden_losses.py).  Same tolerances as
tests/test_gpu_loss.py (north star: loss 1e-4 relative, gradients 1e-3 of the gradient scale); the fixtures are float32 runs, so
1e-5 of slack covers their own round-off.  (The un-normalised ``raw_*`` fixtures are checked on the CPU side only.)"""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = [f for f in sorted(glob.glob(os.path.join(GOLDEN, "loss_*.npz"))) if "loss_raw_" not in f]
LOSS_RTOL = 1e-4
GRAD_RTOL = 1e-3


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[5:-4] for f in FILES])
def test_kernels_match_the_executed_reference(path):
    import multimodal_similarity_b200 as mm
    z = np.load(path)
    kind, weighted = str(z["kind"]), bool(z["weighted"])
    margin = "soft" if str(z["margin"]) == "soft" else float(str(z["margin"]))
    e = torch.from_numpy(z["emb"]).cuda().requires_grad_(True)
    fn = mm.batch_hard if kind == "batch_hard" else mm.lifted_loss
    out = fn(e, torch.from_numpy(z["pids"]).cuda(), margin, weighted)
    out[0].backward()
    grad = e.grad.detach().cpu().numpy()
    n = z["emb"].shape[0]

    assert float(out[0]) == pytest.approx(float(z["loss"]), rel=LOSS_RTOL + 1e-5, abs=1e-7)
    for name, got in zip(("diff", "weights", "furthest_positive", "closest_negative"), out[2:]):
        a = np.broadcast_to(got.detach().cpu().numpy().astype(np.float64), (n,))
        b = np.broadcast_to(z[name].astype(np.float64), (n,))
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), name
        assert np.allclose(a[fin], b[fin], rtol=LOSS_RTOL + 1e-5, atol=2e-6), (name, np.abs(a[fin] - b[fin]).max())
    assert float(out[1]) == pytest.approx(float(z["num_active"]), abs=1.5 / n)      # a row within 1e-5 of the threshold may flip
    ref = z["d_emb"].astype(np.float64)
    scale = np.abs(ref).max()
    if scale == 0:                      # every hinge closed in the reference run: no gradient here either
        assert np.abs(grad).max() <= 1e-7
    else:
        assert np.abs(grad - ref).max() / scale < GRAD_RTOL + 1e-5
