#!/bin/bash
# one-site-per-thread shape of the one-pass D^dag D kernel (normal1.cu): parity under the variant, then timing
TAG=${1:-t08}; OUT=gpurun_out/$TAG; mkdir -p $OUT
GLB_NORMAL_SPT1=44 timeout 900 python -m pytest tests/test_apply_gpu.py tests/test_solvers_gpu.py -m gpu -q -x --timeout 600 > $OUT/pytest_spt1.log 2>&1; echo "pytest (SPT1=44) rc=$?" | tee $OUT/summary.txt
tail -4 $OUT/pytest_spt1.log | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests/test_family_gpu.py tests/test_eo_gpu.py -m gpu -q --timeout 600 > $OUT/pytest_family.log 2>&1; echo "pytest family/eo rc=$?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_family.log | tee -a $OUT/summary.txt
for rep in 1 2; do
  for v in "GLB_NORMAL_SPT1=0" "GLB_NORMAL_SPT1=44" "GLB_NORMAL_SPT1=34" "GLB_NORMAL_SPT1=45" "GLB_NORMAL_SPT1=35" "GLB_NORMAL_SPT1=36"; do
    env $v timeout 300 python tools/tune_variant.py normal 2>&1 | tail -1 | tee -a $OUT/summary.txt
  done
done
