"""A/B of the product sweep with 16 epilogue warps (two TMEM loads in flight each) against 8 (four each), same box, same
process, alternating; the full call's result is checked to be identical first."""
import os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200.retrieval import knn_raw, check_status

dev = torch.device("cuda")
for D in (128, 256):
    g = synth_torch(1_000_000, D, 1000, 12345, dev)
    q = synth_torch(100_000, D, 1000, 12346, dev, centroid_seed=12345)
    res = {}
    for nepi in ("16", "8"):
        os.environ["MMSIM_KNN_NEPI"] = nepi
        d, i, st = knn_raw(q, g, 100)
        fb = check_status(st)
        res[nepi] = (d.clone(), i.clone(), fb)
    same = torch.equal(res["16"][0], res["8"][0]) and torch.equal(res["16"][1], res["8"][1])
    print(f"D={D}: results identical: {same}; fallback queries {res['16'][2]} / {res['8'][2]}", flush=True)
    out = knn_raw(q, g, 100)
    for rep in range(2):
        for nepi in ("16", "8"):
            os.environ["MMSIM_KNN_NEPI"] = nepi
            for _ in range(2):
                knn_raw(q, g, 100, phases=2, out=out)
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5):
                knn_raw(q, g, 100, phases=2, out=out)
            t.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(t) / 5
            print(f"D={D} NEPI={nepi}: sweep {ms:.2f} ms -> {2 * 1e5 * 1e6 * D / ms / 1e9:.0f} TFLOP/s", flush=True)
    del g, q, out, res
    torch.cuda.empty_cache()
