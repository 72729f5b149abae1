#!/bin/bash
# Run every GPU parity test file in its own process (a CUDA trap in one must not poison the others).
# usage (on the GPU box, via gpurun): bash scripts/gpu_tests.sh [pattern]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc_all=0
for f in tests/test_gpu_${1:-*}.py; do
  name=$(basename $f .py)
  timeout -k 10 600 python -m pytest $f -x -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $name rc=$rc: $(tail -1 gpurun_out/$name.log)"
  [ $rc -ne 0 ] && rc_all=1 && grep -E "^(E  |FAILED|mmsim:)" gpurun_out/$name.log | head -30
done
exit $rc_all
