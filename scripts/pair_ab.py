"""A/B of the sweep as single CTAs vs CTA pairs (MMSIM_KNN_PAIR=1, tcgen05 cta_group::2): correctness against the
single-CTA result, then sweep time.  Run each setting in its own process (the switch is read once)."""
import os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_pair_torch
from multimodal_similarity_b200.retrieval import knn_raw

dev = torch.device("cuda")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g, q = synth_pair_torch(1_000_000, 100_000, D, 1000, 12345, dev)
out = knn_raw(q, g, 100)
torch.cuda.synchronize()
print("pair" if os.environ.get("MMSIM_KNN_PAIR") == "1" else "single", "status", out[2][:3].tolist(),
      "checksum", float(out[0].double().sum()), int(out[1].long().sum()))
if len(sys.argv) > 2:
    torch.save((out[0].cpu(), out[1].cpu()), sys.argv[2])
for _ in range(2):
    knn_raw(q, g, 100, phases=2, out=out)
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(4):
    knn_raw(q, g, 100, phases=2, out=out)
t.record()
torch.cuda.synchronize()
ms = s.elapsed_time(t) / 4
print(f"D={D}: sweep {ms:.2f} ms -> {2 * 1e5 * 1e6 * D / ms / 1e9:.0f} TFLOP/s")
