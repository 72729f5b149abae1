"""Where the kNN sweep's time goes: the experiment counters of a MMSIM_DEBUG_BUILD=1 library (knn_tc.cu, g_dbg)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200 import _lib
from multimodal_similarity_b200.retrieval import knn_raw

dev = torch.device("cuda")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
from bench import synth_pair_torch
g, q, ql = synth_pair_torch(1_000_000, 100_000, D, 1000, 12345, dev, return_labels=True)
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * 8)()
names = {5: "8-column groups handed to the candidate path", 6: "epilogue warp cycles (sum)", 7: "32-column warp chunks with a hit"}
for name, qq in (("random order", q), ("sorted by cluster", q[torch.argsort(ql)].contiguous())):
    out = knn_raw(qq, g, 100)
    torch.cuda.synchronize()
    lib.mmsim_debug_counters(buf, 8, 1)
    knn_raw(qq, g, 100, phases=2, out=out)
    torch.cuda.synchronize()
    lib.mmsim_debug_counters(buf, 8, 1)
    chunks = 782 * 3907 * 16 * 2
    print(f"{name}: " + "; ".join(f"{n} {buf[i]:,d}" for i, n in names.items()) + f"; hit share of chunks {buf[7] / chunks:.3f}; "
          f"logged per query {int(out[2][0])} uncertified; log entries/query {float(buf[5]) / 1e5:.0f} groups")
