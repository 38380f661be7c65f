#!/bin/bash
# kernel-shape / schedule sweep of cg_step_kernel + per-CTA traces.   bash tools/gpu_tune.sh <tag>
TAG=${1:-t01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m pytest tests/test_cgstep_gpu.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for g in 1.5; do
  GLB_CGSTEP_GFAC=$g timeout 300 python tools/tune_cgstep.py 4096 508433 508532 508333 508632 508334 508443 504433 512433 2>> $OUT/tune.err | sed "s/^{/{\"gfac\": $g, /" >> $OUT/tune.jsonl
  GLB_CGSTEP_PERSIST=0 GLB_CGSTEP_GFAC=$g timeout 300 python tools/tune_cgstep.py 4096 508433 508532 2>> $OUT/tune.err | sed "s/^{/{\"persist\": 0, \"gfac\": $g, /" | grep -v two-kernel >> $OUT/tune.jsonl
done
for g in 1.25 2; do
  GLB_CGSTEP_GFAC=$g timeout 300 python tools/tune_cgstep.py 4096 508433 508532 2>> $OUT/tune.err | sed "s/^{/{\"gfac\": $g, /" | grep -v two-kernel >> $OUT/tune.jsonl
done
echo "tune rc=$?"; cat $OUT/tune.jsonl; tail -3 $OUT/tune.err
GLB_CGSTEP_GFAC=1.5 GLB_CGSTEP_TRACE=$OUT/trace_persist.txt timeout 200 python tools/tune_cgstep.py 4096 508433 > $OUT/trace.log 2>&1
ls $OUT
