"""A/B of the product sweep: one query block per CTA (MMSIM_KNN_DUAL=0) against two (default), same box, same process,
alternating; the full call's results are checked to be identical first."""
import os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200.retrieval import knn_raw, check_status

dev = torch.device("cuda")
for D in (128, 64):
    g = synth_torch(1_000_000, D, 1000, 12345, dev)
    q = synth_torch(100_000, D, 1000, 12346, dev, centroid_seed=12345)
    res = {}
    for dual in ("0", "1"):
        os.environ["MMSIM_KNN_DUAL"] = dual
        d, i, st = knn_raw(q, g, 100)
        fb = check_status(st)
        res[dual] = (d.clone(), i.clone(), fb)
    same = torch.equal(res["0"][0], res["1"][0]) and torch.equal(res["0"][1], res["1"][1])
    print(f"D={D}: results identical: {same}; fallback queries {res['0'][2]} / {res['1'][2]}", flush=True)
    out = knn_raw(q, g, 100)
    for rep in range(2):
        for dual in ("0", "1"):
            os.environ["MMSIM_KNN_DUAL"] = dual
            for _ in range(2):
                knn_raw(q, g, 100, phases=2, out=out)
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5):
                knn_raw(q, g, 100, phases=2, out=out)
            t.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(t) / 5
            print(f"D={D} DUAL={dual}: sweep {ms:.2f} ms -> {2 * 1e5 * 1e6 * D / ms / 1e9:.0f} TFLOP/s", flush=True)
    del g, q, out, res
    torch.cuda.empty_cache()
