"""Why are half-gallery shards so much more efficient?  Plain single-GPU sweep on 1M / 500k / 250k / 125k rows."""
import os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200.retrieval import knn_raw, check_status

dev = torch.device("cuda")
gfull = synth_torch(1_000_000, 128, 1000, 12345, dev)
q = synth_torch(100_000, 128, 1000, 12346, dev, centroid_seed=12345)   # same mixture as the gallery
for G in (1_000_000, 500_000, 250_000, 125_000):
    g = gfull[:G].contiguous()
    for flags in (0, 2, 4):
        os.environ["MMSIM_SWEEP_FLAGS"] = "0"
        out = knn_raw(q, g, 100)
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
        print(f"G={G} flags={flags}: sweep {ms:.2f} ms -> {2 * 1e5 * G * 128 / ms / 1e9:.0f} TFLOP/s ({ms * 1e6 / G:.1f} ns/row)")
