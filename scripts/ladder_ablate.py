"""Threshold-ladder ranks (MMSIM_LADDER="init,mid,low") vs sweep time, whole-call time and exact-fallback count.
Same-mixture queries, in random order and sorted by cluster label."""
import os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_pair_torch
from multimodal_similarity_b200.retrieval import knn_raw

dev = torch.device("cuda")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g, q, ql = synth_pair_torch(1_000_000, 100_000, D, 1000, 12345, dev, return_labels=True)
qs = q[torch.argsort(ql)].contiguous()

def timed(fn, reps=3):
    for _ in range(2):
        fn()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / reps

for ladder in (sys.argv[2:] or ["12,6,3", "8,5,3", "8,4,2", "7,4,2", "6,4,2"]):
    os.environ["MMSIM_LADDER"] = ladder
    for name, qq in (("random", q), ("sorted", qs)):
        out = knn_raw(qq, g, 100)
        whole = timed(lambda: knn_raw(qq, g, 100, out=out))
        unc = int(out[2][0])
        knn_raw(qq, g, 100, out=out)
        sweep = timed(lambda: knn_raw(qq, g, 100, phases=2, out=out))
        print(f"D={D} ladder={ladder:7s} {name}: whole call {whole:6.2f} ms, sweep {sweep:6.2f} ms, exact-fallback queries {unc}", flush=True)
