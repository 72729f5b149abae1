"""One leave-one-out evaluate call at cfg 3 (or N D C from argv) for ncu captures."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import multimodal_similarity_b200 as mm
from conftest import clustered

n, d, c = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (5924, 128, 100)
x, lab = clustered(np.random.RandomState(12345), n, d, c, first_label=101 if c > 50 else 0)
out = mm.evaluate(x, lab)
torch.cuda.synchronize()
print("mAP", out[0], "R@1", out[5][0])
