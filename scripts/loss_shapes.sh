#!/bin/bash
for s in 0 1 2; do echo "shape $s"; MMSIM_LOSS_SHAPE=$s python scripts/loss_trace.py 2>&1 | tail -2; done
