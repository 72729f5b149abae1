"""BASELINE config 5 at its upper size: 100,000 queries x 10,000,000 gallery rows, 256-d, top-100 on one GPU.
Checks a few queries against a brute-force fp32 computation in torch and prints the timing."""
import sys, time
import torch
sys.path.insert(0, ".")
from bench import synth_torch
import multimodal_similarity_b200 as mm

dev = torch.device("cuda")
G, Q, D = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000, 100_000, 256
g = torch.empty((G, D), device=dev)
for lo in range(0, G, 1_000_000):                      # generate in slabs: bounded temporaries
    hi = min(G, lo + 1_000_000)
    g[lo:hi] = synth_torch(hi - lo, D, 1000, 12345 + lo // 1_000_000, dev, centroid_seed=12345)
q = synth_torch(Q, D, 1000, 999, dev, centroid_seed=12345)
torch.cuda.synchronize()
t0 = time.perf_counter()
d, i = mm.retrieve(q, g, 100)
torch.cuda.synchronize()
t1 = time.perf_counter()
d, i = mm.retrieve(q, g, 100)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"G={G} Q={Q} D={D}: first call {1e3 * (t1 - t0):.0f} ms, second {1e3 * (t2 - t1):.0f} ms -> {Q / (t2 - t1):.0f} queries/s, "
      f"{2.0 * Q * G * D / (t2 - t1) / 1e12:.0f} TFLOP/s end to end; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
for qi in (0, 12345, Q - 1):                           # brute force: same candidates, same order
    dd = torch.cdist(q[qi:qi + 1].double(), g.double() if G <= 2_000_000 else g[i[qi].long()].double()).squeeze(0)
    if G <= 2_000_000:
        ref = torch.topk(dd, 100, largest=False).indices.sort().values
        assert torch.equal(ref, i[qi].sort().values), qi
    else:
        assert torch.allclose(dd.float(), d[qi], rtol=1e-5), qi
        assert bool((d[qi][1:] >= d[qi][:-1]).all())
print("ok")
