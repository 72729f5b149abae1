"""Ablation of the kNN sweep kernel (phase mask 2 only) under the experiment flags, in one process on one box."""
import os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200.retrieval import knn_raw

dev = torch.device("cuda")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = synth_torch(1_000_000, D, 1000, 12345, dev)
q = synth_torch(100_000, D, 1000, 12346, dev, centroid_seed=12345)   # same mixture as the gallery
out = knn_raw(q, g, 100)
torch.cuda.synchronize()
legacy = not os.environ.get("MMSIM_KNN_SWEEP", "").startswith("q")     # default: the gallery-streaming kernel
for flags in ((0, 8, 0) if os.environ.get("MMSIM_KNN_PAIR") == "1" else (0, 8, 16, 32, 0) if not legacy else (0, 8, 2, 4, 0)):
    os.environ["MMSIM_SWEEP_FLAGS"] = str(flags)
    for _ in range(2):
        knn_raw(q, g, 100, phases=2, out=out)
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(4):
        knn_raw(q, g, 100, phases=2, out=out)
    t.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(t) / 4
    print(f"D={D} flags={flags}: sweep {ms:.2f} ms  ->  {2 * 1e5 * 1e6 * D / ms / 1e9:.0f} TFLOP/s")
