"""Diagnostics for tests/test_gpu_certificate.py: the worst logged keys of one case, term by term."""
import sys, zlib, ctypes
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_certificate import CASES, _plan
from multimodal_similarity_b200 import _lib, _util
from multimodal_similarity_b200.retrieval import knn_raw, check_status
name = sys.argv[1] if len(sys.argv) > 1 else "mixed-norms"
_, d, qs, gs, spread = [c for c in CASES if c[0] == name][0]
rs = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
nq, ng, k = 384, 30000, 100
cent = rs.randn(40, d)
g = cent[rs.randint(0, 40, ng)] + 0.7 * rs.randn(ng, d)
q = cent[rs.randint(0, 40, nq)] + 0.7 * rs.randn(nq, d)
g /= np.linalg.norm(g, axis=1, keepdims=True); q /= np.linalg.norm(q, axis=1, keepdims=True)
g *= gs * 10.0 ** (spread * (rs.rand(ng, 1) - 0.5)); q *= qs * 10.0 ** (spread * (rs.rand(nq, 1) - 0.5))
q, g = q.astype(np.float32), g.astype(np.float32)
dev = torch.device("cuda", 0)
dist, idx, status = knn_raw(torch.from_numpy(q).to(dev), torch.from_numpy(g).to(dev), k, phases=63 & ~8); torch.cuda.synchronize()
print("uncertified", int(status[0]))
lib = _lib.load()
p = _plan(lib, nq, ng, d, k, torch.cuda.get_device_properties(0).multi_processor_count)
print(p)
raw = _util.workspace("knn", 0, dev).cpu().numpy()
Dp, S, cap = p["Dp"], p["n_splits"], p["logcap"]
qh = raw[p["off_qh"]:p["off_qh"] + nq * Dp * 2].view(np.float16).reshape(nq, Dp).astype(np.float64)
gh = raw[p["off_gh"]:p["off_gh"] + ng * Dp * 2].view(np.float16).reshape(ng, Dp).astype(np.float64)
nt = p["n_tiles"]
pack = raw[p["off_gpack"]:p["off_gpack"] + nt * 320 * 4].view(np.float32).reshape(nt, 320)
gnorm32 = pack[:, :256].reshape(-1)[:ng].astype(np.float64)
cnt = raw[p["off_log_cnt"]:p["off_log_cnt"] + nq * S * 4].view(np.int32).reshape(nq, S)
tau = raw[p["off_log_tau"]:p["off_log_tau"] + nq * S * 4].view(np.float32).reshape(nq, S)
log = raw[p["off_log"]:p["off_log"] + nq * S * cap * 8].view(np.uint32).reshape(nq, S, cap, 2)
gn = (gh * gh).sum(1); qn = (qh * qh).sum(1) / 4
print("max |gnorm32 - gn| / gn", np.abs(gnorm32 - gn).max(), (np.abs(gnorm32 - gn) / gn).max())
delta = 4.0 * (Dp + 8) * 2.0 ** -24 * (qn + gn.max())
acc_exact = qh @ gh.T
recs = []
for i in range(nq):
    for s in range(S):
        c = min(int(cnt[i, s]), cap)
        ent = log[i, s, :c]
        keys = ent[:, 0].copy().view(np.float32).astype(np.float64); rows = ent[:, 1].astype(np.int64)
        err = keys - (gn[rows] + acc_exact[i, rows])
        for e in np.argsort(-np.abs(err))[:2]:
            recs.append((abs(err[e]) / delta[i], i, int(rows[e]), keys[e], gn[rows[e]], gnorm32[rows[e]], acc_exact[i, rows[e]], keys[e] - gnorm32[rows[e]], qn[i], delta[i], cnt[i, s], tau[i, s]))
recs.sort(reverse=True)
print("ratio query row key_logged gn_exact gn_fp32 acc_exact acc_implied qn delta cnt tau")
for r in recs[:12]:
    print(" ".join(f"{v:.6g}" for v in r))
