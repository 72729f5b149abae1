"""Timings of the evaluation configs of BASELINE.json that are not the headline bench line (SURVEY.md 8(d) cfg 3 and 4):
cfg 3  CUB-style all-pairs leave-one-out evaluate, 5,924 x 128-d                   (fused one-CTA-per-query kernel)
cfg 4  late-fusion leave-one-out evaluate, 20,000 x (128 + 128)-d                 (gallery-scale path, csrc/eval_large.cu)
cfg 4  late-fusion retrieval 20,000 queries x 200,000 gallery, 2 x 128-d, top-50  (tcgen05 sweep + exact re-rank)"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import multimodal_similarity_b200 as mm
from conftest import clustered

rs = np.random.RandomState(12345)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out

x, lab = clustered(rs, 5924, 128, 100, first_label=101)
ms, out = timed(lambda: mm.evaluate(x, lab))
print(f"cfg3 evaluate 5924 x 128 (host arrays in, tuples out): {ms:.1f} ms  ({5924 / ms * 1e3:.0f} queries/s)  mAP={out[0]:.4f} R@1={out[5][0]:.4f}")
xd = torch.from_numpy(x).cuda()
ms, out = timed(lambda: mm.evaluate(xd, lab))
print(f"cfg3 evaluate 5924 x 128 (embeddings resident): {ms:.1f} ms")

cam, lab = clustered(rs, 20000, 128, 7, first_label=0)
sens, _ = clustered(rs, 20000, 128, 7)
fused = mm.late_fusion(cam, sens)
ms, out = timed(lambda: mm.evaluate(fused, lab), reps=1)
print(f"cfg4 leave-one-out evaluate 20000 x 256 fused: {ms:.0f} ms  ({int((lab > 0).sum()) / ms * 1e3:.0f} queries/s)  mAP={out[0]:.4f}")

cam, lab = clustered(rs, 220000, 128, 7, first_label=0)
sens, _ = clustered(rs, 220000, 128, 7)
cq, cg = torch.from_numpy(cam[:20000]).cuda(), torch.from_numpy(cam[20000:]).cuda()
sq, sg = torch.from_numpy(sens[:20000]).cuda(), torch.from_numpy(sens[20000:]).cuda()
ms, out = timed(lambda: mm.retrieve(cq, cg, 50, queries2=sq, gallery2=sg))
print(f"cfg4 late-fusion retrieval 20000 x 200000 x 256, top-50 (resident): {ms:.2f} ms  ({20000 / ms * 1e3:.0f} queries/s)")
