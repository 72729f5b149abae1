"""The certificate's error budget, measured.  The tcgen05 sweep logs approximate keys
key = |g~|^2 - 2 q~.g~ (fp16 operands, fp32 accumulation in the tensor core, fp32 norm added in the epilogue) and the
certificate (csrc/knn_tc.cu, knn_rerank_kernel) assumes
    |key_logged - key_exact(q~, g~)| <= delta = 4 (Dp + 8) 2^-24 (|q~|^2 + max |g~|^2)
and that every gallery row whose key lies below the final threshold by more than delta IS in the log.  Both are checked here
against float64 arithmetic on the very fp16 operand copies the kernel left in its workspace, on adversarial magnitudes."""
import ctypes
import zlib

import numpy as np
import pytest

from oracle import retrieval_np as O

pytestmark = pytest.mark.gpu

CASES = [  # name, D, query scale, gallery scale, spread of per-row norms (log10)
    ("unit-128", 128, 1.0, 1.0, 0.0),
    ("unit-256", 256, 1.0, 1.0, 0.0),
    ("tiny-1e-3", 128, 1e-3, 1e-3, 0.0),
    ("large-1e3", 256, 1e3, 1e3, 0.0),
    ("mixed-norms", 100, 1.0, 1.0, 3.0),               # row norms from 1e-3 to 1e3 in one gallery
    ("near-fp16-overflow", 256, 2.0e4, 4.0e4, 0.0),    # |2 q_i| and |g_i| up to ~6e4 < 65504
    ("large-queries-small-gallery", 160, 1e3, 1e-2, 0.0),
]


def _plan(lib, nq, ng, d, k, sms):
    out = (ctypes.c_int64 * 21)()
    assert lib.mmsim_knn_plan(nq, ng, d, k, sms, out, 21) == 0
    names = ["Dp", "katoms", "n_qblocks", "n_tiles", "n_splits", "tiles_per_split", "grid", "logcap", "use_pivots", "n_sample_tiles",
             "n_sample", "total_bytes", "off_log_cnt", "off_log_tau", "sweepq", "host_splits", "off_log", "off_qh", "off_gh",
             "off_gpack", "n_anchor"]
    return dict(zip(names, list(out)))


@pytest.mark.parametrize("name,d,qs,gs,spread", CASES, ids=[c[0] for c in CASES])
def test_logged_keys_within_delta_and_log_complete(name, d, qs, gs, spread):
    import torch
    import multimodal_similarity_b200 as mm
    from multimodal_similarity_b200 import _lib, _util
    from multimodal_similarity_b200.retrieval import knn_raw, check_status
    rs = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
    nq, ng, k = 384, 30000, 100
    cent = rs.randn(40, d)
    g = cent[rs.randint(0, 40, ng)] + 0.7 * rs.randn(ng, d)
    q = cent[rs.randint(0, 40, nq)] + 0.7 * rs.randn(nq, d)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    if name == "near-fp16-overflow":          # scale so that the largest component sits just below the fp16 range
        g *= gs / np.abs(g).max() * 0.99
        q *= qs / np.abs(q).max() * 0.99
    else:
        g *= gs * 10.0 ** (spread * (rs.rand(ng, 1) - 0.5))
        q *= qs * 10.0 ** (spread * (rs.rand(nq, 1) - 0.5))
    q, g = q.astype(np.float32), g.astype(np.float32)

    dev = torch.device("cuda", 0)
    qd, gd = torch.from_numpy(q).to(dev), torch.from_numpy(g).to(dev)
    # the result is exact whatever the certificate decided
    dist, idx, status = knn_raw(qd, gd, k)
    fallback = check_status(status)
    rd, ri = O.knn(q, g, k)
    assert np.array_equal(dist.cpu().numpy(), rd)
    assert _only_tie_swaps(q, g, idx.cpu().numpy().astype(np.int64), ri, rd)
    # again without the fallback phase (its re-sweep reuses the log for the uncertified queries): the sweep's own log
    _, _, status = knn_raw(qd, gd, k, phases=63 & ~8)
    torch.cuda.synchronize()
    assert int(status[0]) == fallback

    lib = _lib.load()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    p = _plan(lib, nq, ng, d, k, sms)
    assert p["n_anchor"] == 0 and p["use_pivots"] == 1          # sweep order == query order; thresholds in force
    ws = _util.workspace("knn", 0, dev)
    raw = ws.cpu().numpy()
    Dp, S, cap = p["Dp"], p["n_splits"], p["logcap"]
    qh = raw[p["off_qh"]:p["off_qh"] + nq * Dp * 2].view(np.float16).reshape(nq, Dp).astype(np.float64)      # = -2 q~
    gh = raw[p["off_gh"]:p["off_gh"] + ng * Dp * 2].view(np.float16).reshape(ng, Dp).astype(np.float64)
    assert np.isfinite(qh).all() and np.isfinite(gh).all()
    cnt = raw[p["off_log_cnt"]:p["off_log_cnt"] + nq * S * 4].view(np.int32).reshape(nq, S)
    tau = raw[p["off_log_tau"]:p["off_log_tau"] + nq * S * 4].view(np.float32).reshape(nq, S)
    log = raw[p["off_log"]:p["off_log"] + nq * S * cap * 8].view(np.uint32).reshape(nq, S, cap, 2)

    gn = (gh * gh).sum(1)                                       # |g~|^2 exactly (float64)
    qn = (qh * qh).sum(1) / 4.0                                 # |q~|^2
    delta = 4.0 * (Dp + 8) * 2.0 ** -24 * (qn + gn.max())       # per query, the kernel's formula
    key_exact = gn[None, :] + qh @ gh.T                         # [nq, ng] float64: |g~|^2 - 2 q~.g~

    worst = 0.0
    n_logged = n_overflow = 0
    for i in range(nq):
        rows_logged, overflow = [], False
        for s in range(S):
            c = int(cnt[i, s])
            if c > cap:                    # the log dropped entries: the kernel treats the query as uncertified (exact fallback)
                overflow, c = True, cap
            ent = log[i, s, :c]
            keys = ent[:, 0].copy().view(np.float32).astype(np.float64)
            rows = ent[:, 1].astype(np.int64)
            assert (rows < ng).all()
            err = np.abs(keys - key_exact[i, rows])
            worst = max(worst, float((err / delta[i]).max()) if c else 0.0)
            assert (err <= delta[i]).all(), (name, i, s, float(err.max()), float(delta[i]))
            rows_logged.append(rows)
            n_logged += c
        n_overflow += overflow
        rows_logged = np.concatenate(rows_logged) if rows_logged else np.zeros(0, np.int64)
        assert np.unique(rows_logged).size == rows_logged.size                     # no row logged twice
        # completeness: whatever the epilogue's reordered test m < tau - min|g|^2 did, no row below the smallest threshold
        # in force by more than delta may be missing (this is the statement the certificate uses)
        tmin = float(tau[i].min())
        if np.isfinite(tmin) and not overflow:
            must = np.nonzero(key_exact[i] < tmin - delta[i])[0]
            missing = np.setdiff1d(must, rows_logged)
            assert missing.size == 0, (name, i, missing[:5], tmin, float(delta[i]))
    assert n_logged >= nq * k and n_overflow <= fallback
    print(f"{name}: {n_logged} logged keys, worst |error| / delta = {worst:.3f}, exact-fallback queries {fallback}, "
          f"of them {n_overflow} with a full log")


def _only_tie_swaps(q, g, got, ref, rd):
    """indices may differ only where the reference distances tie"""
    for i, j in zip(*np.nonzero(got != ref)):
        d_got = O.l2_to_all(q[i], g[got[i, j]][None])[0]
        if d_got != rd[i, j]:
            return False
    return True
