"""End-to-end time of the host-buffer call (mmsim_knn_host_f32) against the number of gallery splits, next to the
device-resident call: python scripts/host_splits.py [splits ...]   (0 = the plan's own choice)"""
import os
import sys

import torch

sys.path.insert(0, ".")
from bench import synth_pair_torch
from multimodal_similarity_b200.retrieval import check_status, knn_host, knn_raw

dev = torch.device("cuda")
g, q = synth_pair_torch(1_000_000, 100_000, 128, 1000, 12345, dev)[:2]
qh, gh = q.cpu().pin_memory(), g.cpu().pin_memory()
k = 100
res_d = torch.empty((q.shape[0], k), dtype=torch.float32).pin_memory()
res_i = torch.empty((q.shape[0], k), dtype=torch.int32).pin_memory()
st = torch.empty(8, dtype=torch.int32, device=dev)
stage = (torch.empty_like(q), torch.empty_like(g))


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ref = None
for s in [int(v) for v in sys.argv[1:]] or [0, 4, 5, 6, 8]:
    if s:
        os.environ["MMSIM_KNN_SPLITS"] = str(s)
    else:
        os.environ.pop("MMSIM_KNN_SPLITS", None)
    ms_dev = timed(lambda: knn_raw(q, g, k))
    ms_host = timed(lambda: knn_host(qh, gh, k, stage=stage, out=(res_d, res_i, st)))
    fb = check_status(st)
    if ref is None:
        ref = (res_d.clone(), res_i.clone())
    same = torch.equal(res_d, ref[0]) and torch.equal(res_i, ref[1])
    os.environ["MMSIM_HOST_ONE_STREAM"] = "1"
    ms_one = timed(lambda: knn_host(qh, gh, k, stage=stage, out=(res_d, res_i, st)))
    os.environ.pop("MMSIM_HOST_ONE_STREAM")
    same = same and torch.equal(res_d, ref[0]) and torch.equal(res_i, ref[1])
    print(f"splits {s}: device-resident {ms_dev:.2f} ms, host buffers {ms_host:.2f} ms (sweeps on one stream: {ms_one:.2f} ms), fallback queries {fb}, same result {same}", flush=True)
