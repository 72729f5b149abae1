"""Sweep time vs number of logged candidates (initial threshold rank 12 / 6 / 3) on one box."""
import ctypes, os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200 import _lib
from multimodal_similarity_b200._util import _ws_cache
from multimodal_similarity_b200.retrieval import knn_raw

lib = _lib.load()
dev = torch.device("cuda")
g = synth_torch(1_000_000, 128, 1000, 12345, dev)
q = synth_torch(100_000, 128, 1000, 12346, dev)
out = knn_raw(q, g, 100)
ws = [v for k_, v in _ws_cache.items() if k_[0] == "knn"][0]
pl = (ctypes.c_int64 * 14)()
lib.mmsim_knn_plan(100_000, 1_000_000, 128, 100, 148, pl, 14)
for flags in (0, 8, 16, 2):
    os.environ["MMSIM_SWEEP_FLAGS"] = str(flags)
    for _ in range(2):
        knn_raw(q, g, 100, phases=2, out=out)
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(4):
        knn_raw(q, g, 100, phases=2, out=out)
    t.record()
    torch.cuda.synchronize()
    S, rows, off = pl[4], pl[2] * 128, pl[12]
    c = ws[off: off + rows * S * 4].view(torch.int32).view(rows, S)[:100_000].float().sum(1)
    print(f"flags={flags}: sweep {s.elapsed_time(t) / 4:.2f} ms, candidates/query mean {c.mean().item():.0f} max {c.max().item():.0f}, queries with < 128: {(c < 128).sum().item()}")
