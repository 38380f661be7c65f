#!/bin/bash
TAG=${1:-t09}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_family_gpu.py tests/test_apply_gpu.py -m gpu -q --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/summary.txt
tail -4 $OUT/pytest.log | tee -a $OUT/summary.txt
for v in "GLB_NORMAL_SPT1=0" "GLB_DEFAULT=1" "GLB_NORMAL_SPT1=34" "GLB_NORMAL_SPT1=134" "GLB_NORMAL_SPT1=144"; do
  env $v timeout 300 python tools/tune_variant.py normal 2>&1 | tail -1 | tee -a $OUT/summary.txt
done
GLB_NORMAL_SPT1=34 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"normal1_kernel" -s 12 -c 3 \
   -o $OUT/prof_spt1 python bench.py --steps 1 --warmup 3 --no-cpu --apply-reps 5 > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/summary.txt
