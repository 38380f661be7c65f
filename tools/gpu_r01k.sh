#!/bin/bash
# GPU session r01k: the whole suite on the final code (new: BLOCK_CORNER, free-field vectors, argument errors)
TAG=${1:-r01k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest tests/test_mg_setup_gpu.py -s" | tee $OUT/summary.txt
timeout 600 python -m pytest tests/test_mg_setup_gpu.py -m gpu -q -s --timeout 300 > $OUT/pytest_setup.log 2>&1; echo "pytest setup rc=$?" | tee -a $OUT/summary.txt
grep -E "^\.*setup L=|passed|failed|^E  |FAILED|Segmentation" $OUT/pytest_setup.log | tail -30 | cut -c1-250 | tee -a $OUT/summary.txt
echo "== pytest -m gpu (all)" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -12 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
echo "== smoke" | tee -a $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
