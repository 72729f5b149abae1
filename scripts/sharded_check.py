"""Multi-rank check of ShardedGallery under NCCL (run with torchrun on 2/4/8 GPUs): every decomposition and protocol gives the
single-GPU result on every rank, including the per-query repair path (MMSIM_KNN_FORCE_FALLBACK makes every m-th query fail
its certificate) and the end-to-end retrieve_host slices.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 scripts/sharded_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import multimodal_similarity_b200 as mm
from multimodal_similarity_b200.sharded import ShardedGallery
from conftest import clustered

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rs = np.random.RandomState(5)
x, _ = clustered(rs, 60000, 64, 50)
qn = clustered(rs, 4203, 64, 50)[0]
g, q = torch.from_numpy(x).to(dev), torch.from_numpy(qn).to(dev)
ok = True


def check(name, got, ref):
    global ok
    same = torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
    ok &= same
    print(f"[rank {rank}] {name}: {'ok' if same else 'MISMATCH'}", flush=True)


for k, excl in ((20, False), (100, True)):
    qq = g[:4203].contiguous() if excl else q
    ref = mm.retrieve(qq, g, k, exclude_self=excl)
    groups = sorted({1, world} | ({world // 2} if world >= 4 else set()))
    for R in groups:
        sg = ShardedGallery(g, query_groups=R)
        for proto in ("auto", "exact-shards"):
            out = sg.retrieve(qq, k, exclude_self=excl, protocol=proto)
            check(f"k={k} excl={excl} query_groups={R} protocol={proto} -> {sg.last_protocol}", out, ref)
        if sg.parts > 1:
            os.environ["MMSIM_KNN_FORCE_FALLBACK"] = "97"          # every 97th query fails its certificate -> repaired one by one
            out = sg.retrieve(qq, k, exclude_self=excl)
            del os.environ["MMSIM_KNN_FORCE_FALLBACK"]
            check(f"k={k} excl={excl} query_groups={R} forced repair ({sg.last_protocol}, repaired {sg.last_repaired}, "
                  f"uncertified {int(sg.last_uncertified)})", out, ref)
            ok &= sg.last_protocol == "reduced" and sg.last_repaired > 0
    if world > 1:
        sg = ShardedGallery(g, query_groups=1)
        qh = qq.cpu().pin_memory()
        gh = sg.shard.cpu().pin_memory()
        d, i, (lo, hi) = sg.retrieve_host(qh, k, gallery_host=gh, exclude_self=excl)
        check(f"k={k} excl={excl} retrieve_host slice [{lo}, {hi})", (d.to(dev), i.to(dev)), (ref[0][lo:hi], ref[1][lo:hi]))
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
if rank == 0:
    print("SHARDED CHECK", "PASSED" if int(flag) == 0 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(int(flag) != 0)
