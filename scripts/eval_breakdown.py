"""Where the time of the leave-one-out evaluation goes (cfg 3: 5,924 x 128; cfg 4: 20,000 x 256): device time of the
kernels alone (CUDA events around the C-ABI call, inputs resident) against the whole `evaluate` call (host arrays in,
reference tuple out)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import multimodal_similarity_b200 as mm
from multimodal_similarity_b200 import retrieval as R
from conftest import clustered

rs = np.random.RandomState(12345)


def dev_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def wall_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for name, n, d, ncls, first in (("cfg3 5924x128/100 classes", 5924, 128, 100, 101), ("cfg4 20000x256/7 classes", 20000, 256, 7, 0),
                                ("12000x128/200 classes", 12000, 128, 200, 1)):
    x, lab = clustered(rs, n, d, ncls, first_label=first)
    xd = torch.from_numpy(x).cuda()
    reps = 5 if n < 10000 else 2
    k = dev_ms(lambda: R._loo_device(xd, lab, 0.5, False) if hasattr(R, "_loo_device") else R._loo_records(xd, lab, False, False, 0.5, False), reps)
    w = wall_ms(lambda: mm.evaluate(x, lab), reps)
    wd = wall_ms(lambda: mm.evaluate(xd, lab), reps)
    ws = wall_ms(lambda: mm.evaluate_simple(x, lab), reps)
    print(f"{name}: records (device events) {k:.2f} ms | evaluate(host arrays) {w:.2f} ms | evaluate(resident) {wd:.2f} ms | evaluate_simple {ws:.2f} ms", flush=True)
