#!/bin/bash
# A/B builds of libvp_b200.so (vision-processor_b200/lib/ab/*.so): the pipeline (3 lanes, automatic groups) and the per-kernel times
# (1 lane, a ring of 384 frames so that the inputs come from HBM like in bench.py) of each
for so in "$@"; do
  echo "== $so"
  VPB200_LIB=$PWD/$so python tools/prof_step.py --batch 384 --lanes 3 --steps 10 --warmup 3 --times 2>&1 | grep -E "us/frame" | head -1
  VPB200_LIB=$PWD/$so python tools/prof_step.py --batch 384 --group 128 --lanes 1 --steps 5 --warmup 2 --times 2>&1 | grep -E "grad_circ|reproject|us/frame, "
done
