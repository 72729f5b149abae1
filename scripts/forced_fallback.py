"""What the exact fallback costs at the headline size: every m-th query is forced to count as uncertified
(MMSIM_KNN_FORCE_FALLBACK=m) -- ~1,000 and ~10,000 fallbacks in a 100k-query call; results must stay identical."""
import os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200.retrieval import knn_raw, check_status

dev = torch.device("cuda")
g = synth_torch(1_000_000, 128, 1000, 12345, dev)
q = synth_torch(100_000, 128, 1000, 12346, dev, centroid_seed=12345)


def run(reps=3):
    d, i, st = knn_raw(q, g, 100); n = check_status(st)
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        d, i, st = knn_raw(q, g, 100)
    t.record(); torch.cuda.synchronize()
    return s.elapsed_time(t) / reps, n, d.clone(), i.clone()


base_ms, n0, d0, i0 = run()
print(f"no forced fallback: {base_ms:.2f} ms per call, {n0} fallback queries")
for m in (100, 10):
    os.environ["MMSIM_KNN_FORCE_FALLBACK"] = str(m)
    ms, n, d, i = run()
    print(f"every {m}th query forced ({n} fallbacks): {ms:.2f} ms per call (+{ms - base_ms:.2f} ms), identical: {torch.equal(d, d0) and torch.equal(i, i0)}")
del os.environ["MMSIM_KNN_FORCE_FALLBACK"]
