#!/bin/bash
# 1 -> 8 GPU scaling of the headline bench on ONE box (run under `gpurun --gpus 8`).
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 --no-extras > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/scale_n$n.json") if l.startswith("{")][-1])
    print("N=$n", round(d["value"]), "q/s", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["value"]), " kernel_ms", round(d["roofline"]["kernel_ms"], 2),
          "frac", round(d["roofline"]["frac"], 3), d["roofline"]["other_kernels_ms"], "fallback", d["exact_fallback_queries"], d.get("protocol"))
except Exception as e:
    print("N=$n FAILED", e)
    print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
