#!/bin/bash
# A/B the sweep kernel under different experiment flags in ONE gpurun call (same box, interleaved).
for rep in 1 2; do
for f in 0 1; do
  MMSIM_SWEEP_FLAGS=$f python bench.py --steps 4 --warmup 2 --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('flags=$f rep=$rep', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],2), d['clocks'])"
done; done
