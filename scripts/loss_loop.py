"""Tiny driver for profiling the fused loss kernels: a few eager calls of batch-hard (256 x 128) and lifted (512 x 128)."""
import sys
import torch
sys.path.insert(0, ".")
from multimodal_similarity_b200.losses import _run

dev = torch.device("cuda")
for n, kind, soft, margin in ((256, 0, True, 0.0), (512, 1, False, 1.0)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    e = torch.randn(n, 128, generator=g, device=dev)
    e = e / e.norm(dim=1, keepdim=True)
    pids = (torch.arange(n, device=dev) % 32 + 1).float() if n == 256 else (torch.arange(n, device=dev) % 7).float()
    for _ in range(6):
        out = _run(kind, e, pids, soft, margin, True, True)
    torch.cuda.synchronize()
    print(n, float(out[0][0]))
