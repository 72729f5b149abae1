"""GPU timeline (CUPTI via torch.profiler) of ONE ShardedGallery.retrieve_host step on rank 0: every kernel / memcpy with its
stream, start and duration -- where does the end-to-end step wait?  Run under torchrun (2+ ranks)."""
import os, sys
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
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
sg = ShardedGallery(torch.empty_like(shard), presharded=True, row_offset=lo, total_rows=G)
for _ in range(3):
    sg.retrieve_host(q_host, k, gallery_host=g_host)
dist.barrier(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    sg.retrieve_host(q_host, k, gallery_host=g_host)
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    print(f"{'start us':>9} {'dur us':>8} {'stream':>6}  name")
    for e in evs:
        print(f"{e.time_range.start - t0:9.0f} {e.time_range.end - e.time_range.start:8.0f} {getattr(e, 'device_index', 0):>6}  {e.name[:90]}")
    print("total span us", evs[-1].time_range.end - t0)
dist.destroy_process_group()
