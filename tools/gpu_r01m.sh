#!/bin/bash
# 2-GPU session r01m: multigrid set-up + solve on y-slabs (slab_check), and the single-GPU set-up tests on the changed kernels
TAG=${1:-r01m}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest tests/test_mg_setup_gpu.py tests/test_mg_gpu.py (one GPU)" | tee $OUT/summary.txt
timeout 300 python -m pytest tests/test_mg_setup_gpu.py tests/test_mg_gpu.py -m gpu -q --timeout 200 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest.log | cut -c1-200 | tee -a $OUT/summary.txt
echo "== slab check x2, L=64" | tee -a $OUT/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
   tools/slab_check.py 64 > $OUT/slab_check_64.log 2>&1; echo "slab rc=$?" | tee -a $OUT/summary.txt
grep -E " ok | FAIL|SLAB|Error|error|Traceback" $OUT/slab_check_64.log | cut -c1-200 | tee -a $OUT/summary.txt
tail -15 $OUT/slab_check_64.log | cut -c1-300 >> $OUT/summary.txt
