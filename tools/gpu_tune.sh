#!/bin/bash
TAG=${1:-t01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m pytest tests/test_cgstep_gpu.py tests/test_baseline_sizes_gpu.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for shape in 4096 4096x512 8192x1024 256; do
  timeout 300 python tools/tune_cgstep.py $shape 508433 2>> $OUT/tune.err >> $OUT/tune.jsonl
done
cut -c1-200 $OUT/tune.jsonl; tail -3 $OUT/tune.err
GLB_CGSTEP_TRACE=$OUT/trace_persist.txt timeout 200 python tools/tune_cgstep.py 4096 508433 > $OUT/trace.log 2>&1
