"""Candidates logged per query by the sweep (plain single-GPU path vs the reduced sharded protocol with simulated shards)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from bench import synth_torch
from multimodal_similarity_b200 import _lib
from multimodal_similarity_b200._util import _ws_cache
from multimodal_similarity_b200.retrieval import knn_raw
from multimodal_similarity_b200.sharded import ReducedShard, merge_pivots_into, reduced_kp, shard_bounds

lib = _lib.load()
dev = torch.device("cuda")
gfull = synth_torch(1_000_000, 128, 1000, 12345, dev)
q = synth_torch(100_000, 128, 1000, 12346, dev, centroid_seed=12345)   # same mixture as the gallery


def counts(ws, nq, ng):
    out = (ctypes.c_int64 * 14)()
    lib.mmsim_knn_plan(nq, ng, 128, 100, 148, out, 14)
    S, rows, off = out[4], out[2] * 128, out[12]
    c = ws[off: off + rows * S * 4].view(torch.int32).view(rows, S)[:nq].float()
    return S, c.sum(1).mean().item(), c.sum(1).max().item()


for G in (1_000_000, 500_000, 125_000):
    g = gfull[:G].contiguous()
    knn_raw(q, g, 100)
    torch.cuda.synchronize()
    ws = [v for k_, v in _ws_cache.items() if k_[0] == "knn"][0]
    print(f"plain G={G}: splits, mean, max candidates per query =", counts(ws, 100_000, G))

for R in (2, 8):
    kp = reduced_kp(R, 100)
    shards, packed, pivs = [], [], []
    for r in range(R):
        lo, hi = shard_bounds(1_000_000, R, r)
        shards.append(ReducedShard(gfull[lo:hi].contiguous(), lo))
        packed.append(torch.empty(ReducedShard.packed_elems(100_000, kp), dtype=torch.int32, device=dev))
        pivs.append(shards[r].stage1(q, 100, kp, packed[r]))
    allpiv = torch.stack(pivs)
    for r in range(R):
        merge_pivots_into(allpiv, pivs[r])
        shards[r].stage2(q, 100, kp, False, 0, packed[r])
    torch.cuda.synchronize()
    print(f"reduced R={R}: shard 0 splits, mean, max candidates per query =", counts(shards[0].ws, 100_000, shards[0].shard.shape[0]))
    del shards, packed, pivs, allpiv
