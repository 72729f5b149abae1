#!/bin/bash
# 2-D decomposition against 1-D gallery sharding at N = 4 and 8 on ONE box (run under `gpurun --gpus 8`).
mkdir -p gpurun_out
for n in ${NS:-4 8}; do
  for qg in 1 auto; do
    out=gpurun_out/scaleg_n${n}_qg${qg}
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 --no-extras --query-groups $qg > $out.json 2> $out.err
    python - <<PY
import json
try:
    d = json.loads([l for l in open("$out.json") if l.startswith("{")][-1])
    print("N=$n query_groups=$qg ->", d["config"]["query_groups"], "x", d["config"]["gallery_parts"], ":", round(d["value"]), "q/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), "ms  sweep", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 3), {k: round(v, 3) for k, v in d["roofline"]["other_kernels_ms"].items()}, "unc", d["exact_fallback_queries"], d.get("protocol"), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("N=$n qg=$qg FAILED", e)
    print(open("$out.err").read()[-1500:])
PY
  done
done
