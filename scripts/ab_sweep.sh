#!/bin/bash
# A/B the sweep kernel under different experiment flags in ONE gpurun call (same box, interleaved).
for f in ${1:-0 2 4}; do
  MMSIM_SWEEP_FLAGS=$f python bench.py --steps 3 --warmup 2 --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('flags=$f', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],2), d['clocks'], d['exact_fallback_queries'])"
done
