"""Where the end-to-end step of the sharded call goes at N > 1: host->device copies, the sharded retrieve, the
device->host copy of the result, timed with CUDA events per rank -- first as launched, then again with the process
bound to the CPUs next to its GPU (NVML's ideal affinity) and the pinned buffers re-allocated from there.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/e2e_sharded_profile.py"""
import os
import subprocess
import sys

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
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)


def profile(tag):
    q_host, g_host = queries.cpu().pin_memory(), shard.cpu().pin_memory()
    res_d = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    res_i = torch.empty((Q, k), dtype=torch.int64).pin_memory()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    acc = [0.0, 0.0, 0.0]
    reps, warm = 5, 2
    for it in range(warm + reps):
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        qd = q_host.to(dev, non_blocking=True)
        gd = g_host.to(dev, non_blocking=True)
        ev[1].record()
        d_, i_ = ShardedGallery(gd, presharded=True, row_offset=lo, total_rows=G).retrieve(qd, k, check=False)
        ev[2].record()
        res_d.copy_(d_, non_blocking=True)
        res_i.copy_(i_, non_blocking=True)
        ev[3].record()
        torch.cuda.synchronize()
        if it >= warm:
            for j in range(3):
                acc[j] += ev[j].elapsed_time(ev[j + 1]) / reps
    h2d_mb = (q_host.numel() + g_host.numel()) * 4 / 1e6
    d2h_mb = (res_d.numel() * 4 + res_i.numel() * 8) / 1e6
    print(f"[{tag}] rank {rank}: cpus {len(os.sched_getaffinity(0))}  h2d {acc[0]:.2f} ms ({h2d_mb / acc[0]:.1f} GB/s)  "
          f"retrieve {acc[1]:.2f} ms  d2h {acc[2]:.2f} ms ({d2h_mb / acc[2]:.1f} GB/s)  total {sum(acc):.2f} ms", flush=True)


def back_to_back(steps=10):
    """The step exactly as bench.py times it (no synchronisation between steps), plus where the HOST spends its time."""
    import time
    q_host, g_host = queries.cpu().pin_memory(), shard.cpu().pin_memory()
    res_d = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    res_i = torch.empty((Q, k), dtype=torch.int64).pin_memory()
    host = [0.0, 0.0, 0.0]

    def step():
        t0 = time.perf_counter()
        qd = q_host.to(dev, non_blocking=True)
        gd = g_host.to(dev, non_blocking=True)
        t1 = time.perf_counter()
        d_, i_ = ShardedGallery(gd, presharded=True, row_offset=lo, total_rows=G).retrieve(qd, k, check=False)
        t2 = time.perf_counter()
        res_d.copy_(d_, non_blocking=True)
        res_i.copy_(i_, non_blocking=True)
        t3 = time.perf_counter()
        for j, v in enumerate((t1 - t0, t2 - t1, t3 - t2)):
            host[j] += v * 1e3 / steps

    for _ in range(2):
        step()
    host[:] = [0.0, 0.0, 0.0]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    a.record()
    for _ in range(steps):
        step()
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - w0) * 1e3 / steps
    print(f"[back to back] rank {rank}: {a.elapsed_time(b) / steps:.2f} ms/step by events, {wall:.2f} ms/step wall; host time per "
          f"step: h2d enqueue {host[0]:.2f}, retrieve call {host[1]:.2f}, d2h enqueue {host[2]:.2f} ms", flush=True)


profile("as launched")
back_to_back()
try:
    if os.environ.get("SKIP_AFFINITY"):
        raise RuntimeError("skipped (SKIP_AFFINITY)")
    import pynvml
    pynvml.nvmlInit()
    uuid = str(torch.cuda.get_device_properties(local).uuid)
    handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
    pynvml.nvmlDeviceSetCpuAffinity(handle)
    profile("bound to the GPU's CPUs")
except Exception as e:  # noqa: BLE001
    print(f"rank {rank}: affinity not set: {e!r}", flush=True)
dist.destroy_process_group()
