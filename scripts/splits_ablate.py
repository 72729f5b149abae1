"""Sweep-kernel time vs number of gallery splits (MMSIM_KNN_SPLITS) on one box."""
import os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200.retrieval import knn_raw, check_status

dev = torch.device("cuda")
g = synth_torch(1_000_000, 128, 1000, 12345, dev)
q = synth_torch(100_000, 128, 1000, 12346, dev, centroid_seed=12345)   # same mixture as the gallery
for S in (0, 1, 2, 3, 4, 6, 8, 12, 16):
    if S:
        os.environ["MMSIM_KNN_SPLITS"] = str(S)
    out = knn_raw(q, g, 100)
    fb = check_status(out[2])
    for _ in range(2):
        knn_raw(q, g, 100, phases=2, out=out)
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        knn_raw(q, g, 100, phases=2, out=out)
    t.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(t) / 3
    s.record()
    for _ in range(3):
        knn_raw(q, g, 100, out=out)
    t.record()
    torch.cuda.synchronize()
    print(f"splits={S or 'auto'}: sweep {ms:.2f} ms ({2 * 1e5 * 1e6 * 128 / ms / 1e9:.0f} TFLOP/s), full step {s.elapsed_time(t) / 3:.2f} ms, fallback {fb}")
