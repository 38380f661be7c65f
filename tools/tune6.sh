#!/bin/bash
# ring-layout / unrolled normal kernel and cp.async-ring coarse kernel: parity first, then variants
TAG=${1:-t06}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_solvers_gpu.py -m gpu -q -x --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/summary.txt
tail -5 $OUT/pytest.log | tee -a $OUT/summary.txt
for rep in 1 2; do
  for v in "GLB_NORMAL_UNROLL=3 GLB_NORMAL_STAGES=4" "GLB_NORMAL_UNROLL=1 GLB_NORMAL_STAGES=4" "GLB_NORMAL_UNROLL=3 GLB_NORMAL_STAGES=3" \
           "GLB_NORMAL_UNROLL=3 GLB_NORMAL_STAGES=4 GLB_NORMAL_STAGES_PLAIN=4" "GLB_NORMAL_UNROLL=3 GLB_NORMAL_STAGES=4 GLB_NORMAL_STAGES_PLAIN=3"; do
    env $v timeout 300 python tools/tune_variant.py normal 2>&1 | tail -1 | tee -a $OUT/summary.txt
  done
done
for v in "GLB_COARSE_RING=0" "GLB_COARSE_RING=1 GLB_COARSE_STAGES=3" "GLB_COARSE_RING=1 GLB_COARSE_STAGES=4" "GLB_COARSE_RING=1 GLB_COARSE_STAGES=6"; do
  env $v timeout 300 python tools/tune_variant.py coarse 2>&1 | grep stencil | tee -a $OUT/summary.txt
done
