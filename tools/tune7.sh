#!/bin/bash
# A/B of the cp.async ring layout (0 private slots / 1 line-contiguous) x row-loop unrolling, then the new bench line
TAG=${1:-t07}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_apply_gpu.py -m gpu -q -x --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/summary.txt
tail -3 $OUT/pytest.log | tee -a $OUT/summary.txt
for rep in 1 2; do
  for v in "GLB_NORMAL_LAYOUT=0 GLB_NORMAL_UNROLL=1" "GLB_NORMAL_LAYOUT=0 GLB_NORMAL_UNROLL=3" "GLB_NORMAL_LAYOUT=1 GLB_NORMAL_UNROLL=1" \
           "GLB_NORMAL_LAYOUT=1 GLB_NORMAL_UNROLL=3" "GLB_NORMAL_LAYOUT=1 GLB_NORMAL_UNROLL=3 GLB_NORMAL_STAGES=3" "GLB_NORMAL_LAYOUT=0 GLB_NORMAL_UNROLL=3 GLB_NORMAL_STAGES=3"; do
    env $v timeout 300 python tools/tune_variant.py normal 2>&1 | tail -1 | tee -a $OUT/summary.txt
  done
done
timeout 600 python bench.py --steps 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
tail -c 3000 $OUT/bench.json | tee -a $OUT/summary.txt; tail -5 $OUT/bench.err | tee -a $OUT/summary.txt
