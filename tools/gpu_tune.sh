#!/bin/bash
# schedule sweep of cg_step_kernel at the per-GPU slab shapes of the strong-scaling runs.   bash tools/gpu_tune.sh <tag>
TAG=${1:-t01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for shape in 4096x512 4096x1024 8192x1024 4096x2048; do
  for g in 1 1.5 2.5; do
    GLB_CGSTEP_GFAC=$g timeout 300 python tools/tune_cgstep.py $shape 508433 516433 532433 2>> $OUT/tune.err | sed "s/^{/{\"gfac\": $g, /" | grep -v two-kernel >> $OUT/tune.jsonl
  done
  timeout 300 python tools/tune_cgstep.py $shape 433 16433 32433 64433 2>> $OUT/tune.err >> $OUT/tune.jsonl
done
cut -c1-230 $OUT/tune.jsonl; tail -3 $OUT/tune.err
