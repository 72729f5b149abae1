"""One full-size kNN call (for ncu captures): python scripts/one_call.py [D] [sorted]"""
import sys
import torch
sys.path.insert(0, ".")
from bench import synth_pair_torch
from multimodal_similarity_b200.retrieval import knn_raw

dev = torch.device("cuda")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g, q, ql = synth_pair_torch(1_000_000, 100_000, D, 1000, 12345, dev, return_labels=True)
if len(sys.argv) > 2 and sys.argv[2] == "sorted":
    q = q[torch.argsort(ql)].contiguous()
out = knn_raw(q, g, 100)
torch.cuda.synchronize()
print("status", out[2].tolist())
