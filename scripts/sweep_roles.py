"""Per-role cycle breakdown of the kNN sweep from a MMSIM_DEBUG_BUILD=1 library:
    MMSIM_DEBUG_BUILD=1 MMSIM_LIB_OUT=$PWD/multimodal_similarity_b200/libmmsim_dbg.so python -m multimodal_similarity_b200.build --force
    MMSIM_LIB=$PWD/multimodal_similarity_b200/libmmsim_dbg.so python scripts/sweep_roles.py [D]
Counters are sums over the 148 CTAs of one sweep launch (lane 0 of the MMA warp, lane 0 of the first epilogue warp)."""
import ctypes, os, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200 import _lib
from multimodal_similarity_b200.retrieval import knn_raw

dev = torch.device("cuda")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = synth_torch(1_000_000, D, 1000, 12345, dev)
q = synth_torch(100_000, D, 1000, 12346, dev, centroid_seed=12345)
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * 16)()
out = knn_raw(q, g, 100)
torch.cuda.synchronize()
for flags in ("0", "8"):
    os.environ["MMSIM_SWEEP_FLAGS"] = flags
    knn_raw(q, g, 100, phases=2, out=out)
    torch.cuda.synchronize()
    lib.mmsim_debug_counters(buf, 16, 1)
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    knn_raw(q, g, 100, phases=2, out=out)
    t.record()
    torch.cuda.synchronize()
    lib.mmsim_debug_counters(buf, 16, 1)
    b = list(buf)
    tiles = max(b[9], 1)
    print(f"D={D} flags={flags} ({'product' if flags == '0' else 'every row closed: no candidates'}): sweep {s.elapsed_time(t):.2f} ms, "
          f"{tiles:,d} tiles over 148 CTAs")
    print(f"  MMA warp     : {b[0] / tiles:7.1f} cycles per tile = {b[1] / tiles:6.1f} waiting for a drained accumulator "
          f"+ {b[2] / tiles:6.1f} waiting for shared-memory stages + {(b[0] - b[1] - b[2]) / tiles:6.1f} issuing / other")
    print(f"  epilogue warp: {b[8] / tiles:7.1f} cycles per tile = {b[3] / tiles:6.1f} waiting for a ready accumulator "
          f"+ {b[4] / tiles:6.1f} ready -> released (TMEM loads) + {(b[8] - b[3] - b[4]) / tiles:6.1f} scan / candidates / other")
