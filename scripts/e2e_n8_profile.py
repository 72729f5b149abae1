"""Where the sharded end-to-end step goes at N ranks (run under torchrun): concurrent host->device bandwidth per rank, then
ShardedGallery.retrieve_host with (a) only the queries coming from the host, (b) queries + gallery shard from the host, with
different numbers of gallery splits in the copy/sweep pipeline, against the device-resident retrieve."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from bench import SEED, WORKLOAD, synth_torch
from multimodal_similarity_b200.sharded import ShardedGallery, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Q, G, D, k = WORKLOAD["queries"], WORKLOAD["gallery"], WORKLOAD["dim"], WORKLOAD["k"]
lo, hi = shard_bounds(G, world, rank)
full = synth_torch(G, D, WORKLOAD["clusters"], SEED, dev)
shard = full[lo:hi].clone()
del full
queries = synth_torch(Q, D, WORKLOAD["clusters"], SEED + 1, dev, centroid_seed=SEED)
q_host, g_host = queries.cpu().pin_memory(), shard.cpu().pin_memory()


def timed(fn, steps=6, warm=2):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    t.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(t) / steps], device=dev)
    lo_ = ms.clone()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    return float(ms), float(lo_)


def say(*a):
    if rank == 0:
        print(*a, flush=True)


stage = torch.empty_like(shard)
mx, mn = timed(lambda: stage.copy_(g_host, non_blocking=True))
mb = g_host.numel() * 4 / 1e6
say(f"H2D of the {mb:.0f} MB shard, all {world} ranks at once: {mx:.2f} ms slowest ({mb / mx:.1f} GB/s), {mn:.2f} ms fastest ({mb / mn:.1f} GB/s)")
res = torch.empty((Q // world, k), dtype=torch.float32).pin_memory()
dsrc = torch.empty((Q // world, k), dtype=torch.float32, device=dev)
mx, mn = timed(lambda: res.copy_(dsrc, non_blocking=True))
say(f"D2H of a {res.numel() * 4 / 1e6:.1f} MB result slice, all ranks at once: {mx:.3f} ms slowest")

sg = ShardedGallery(shard, presharded=True, row_offset=lo, total_rows=G)
mx, _ = timed(lambda: sg.retrieve(queries, k, check=False))
say(f"device-resident retrieve (1 x {world}): {mx:.2f} ms")
sg_e = ShardedGallery(torch.empty_like(shard), presharded=True, row_offset=lo, total_rows=G)
sg_q = ShardedGallery(shard, presharded=True, row_offset=lo, total_rows=G)
mx, _ = timed(lambda: sg_q.retrieve_host(q_host, k))
say(f"retrieve_host, queries from the host, shard resident: {mx:.2f} ms")
for splits in ("default", "1", "2", "3", "4"):
    if splits == "default":
        os.environ.pop("MMSIM_KNN_SPLITS", None)
    else:
        os.environ["MMSIM_KNN_SPLITS"] = splits
    sg_e._reduced_host = None
    mx, _ = timed(lambda: sg_e.retrieve_host(q_host, k, gallery_host=g_host))
    say(f"retrieve_host, queries + shard from the host, gallery splits = {splits}: {mx:.2f} ms")
os.environ.pop("MMSIM_KNN_SPLITS", None)
# host-side time of one call (is the launch path the bottleneck at this step size?)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    sg_e.retrieve_host(q_host, k, gallery_host=g_host)
host_ms = (time.perf_counter() - t0) / 5 * 1e3
torch.cuda.synchronize()
say(f"host wall time per retrieve_host call (includes its final sync): {host_ms:.2f} ms")
dist.destroy_process_group()
