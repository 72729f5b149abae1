"""Host-side glue shared by the reference-signature wrappers: array conversion, stream handle, workspace cache."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_ws_cache: dict = {}


def to_cuda_f32(x, device=None) -> torch.Tensor:
    """numpy / torch (any device) -> contiguous float32 CUDA tensor (no copy when already in that form)."""
    _lib.require_cuda(torch)
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    elif not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x, dtype=np.float32))
    if device is None:
        device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return x.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def is_numpy_like(x) -> bool:
    return not torch.is_tensor(x)


def stream_handle(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def workspace(tag: str, nbytes: int, device, zero: bool = False) -> torch.Tensor:
    """Per (tag, device, stream) byte buffer, grown on demand.  The C-ABI never allocates: we pass this in.
    zero=True hands out a zero-filled buffer on (re)allocation (the loss kernels keep theirs zeroed between calls)."""
    key = (tag, device.index, stream_handle(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        alloc = torch.zeros if zero else torch.empty
        buf = alloc(max(int(nbytes), 1024), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def out_like_input(t: torch.Tensor, as_numpy: bool):
    return t.cpu().numpy() if as_numpy else t
