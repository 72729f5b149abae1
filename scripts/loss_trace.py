"""Phase timeline of the fused loss kernel (CTA 0 globaltimer stamps written into the tail of the workspace)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from multimodal_similarity_b200 import _lib
from multimodal_similarity_b200.losses import _run
from multimodal_similarity_b200._util import _ws_cache

dev = torch.device("cuda")
names = ["start", "staged", "dist done", "reduced(arrive)", "barrier passed", "W done", "phase2 done", "last CTA end"]
for n, kind, soft, margin in ((256, 0, True, 0.0), (512, 1, False, 1.0)):
    e = torch.randn(n, 128, device=dev); e = e / e.norm(dim=1, keepdim=True)
    pids = (torch.arange(n, device=dev) % 32 + 1).float() if n == 256 else (torch.arange(n, device=dev) % 7).float()
    for _ in range(5):
        _run(kind, e, pids, soft, margin, True, True)
    torch.cuda.synchronize()
    ws = [v for k, v in _ws_cache.items() if k[0] == f"loss{n}x128"][0]
    nb = ctypes.c_size_t(); _lib.load().mmsim_loss_workspace_bytes(n, 128, ctypes.byref(nb))
    tr = ws[nb.value - 1024: nb.value].view(torch.int64)
    # the trace block is the last 64-byte region (256-byte aligned) of the layout
    t = ws[: nb.value].view(torch.int64)[-32:].cpu().tolist()
    st = [x for x in t if x > 0][:8]
    print(f"N={n}:", [f"{names[i]} +{(st[i] - st[0]) / 1e3:.2f}us" for i in range(len(st))])
