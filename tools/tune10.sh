#!/bin/bash
# self-scheduling one-pass kernel (normal_ws.cu): parity under the variant, then timing; pending parity tests
TAG=${1:-t10}; OUT=gpurun_out/$TAG; mkdir -p $OUT
GLB_NORMAL_WS=2 timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_solvers_gpu.py -m gpu -q -x --timeout 600 > $OUT/pytest_ws.log 2>&1; echo "pytest (WS=2) rc=$?" | tee $OUT/summary.txt
tail -6 $OUT/pytest_ws.log | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests/test_family_gpu.py tests/test_eo_gpu.py -m gpu -q --timeout 600 > $OUT/pytest_new.log 2>&1; echo "pytest family/eo rc=$?" | tee -a $OUT/summary.txt
tail -6 $OUT/pytest_new.log | tee -a $OUT/summary.txt
for rep in 1 2; do
  for v in "GLB_NORMAL_WS=0" "GLB_NORMAL_WS=1" "GLB_NORMAL_WS=2 GLB_NORMAL_SPT1=0"; do
    env $v timeout 300 python tools/tune_variant.py normal 2>&1 | tail -1 | tee -a $OUT/summary.txt
  done
done
