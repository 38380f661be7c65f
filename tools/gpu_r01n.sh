#!/bin/bash
# GPU session r01n: the rest of operators.h + the whole suite on the final code
TAG=${1:-r01n}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest tests/test_operators_rest_gpu.py" | tee $OUT/summary.txt
timeout 300 python -m pytest tests/test_operators_rest_gpu.py -m gpu -q --timeout 200 > $OUT/pytest_ops.log 2>&1; echo "rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest_ops.log | cut -c1-220 | tee -a $OUT/summary.txt
echo "== pytest -m gpu (all)" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -12 $OUT/pytest_gpu.log | cut -c1-220 | tee -a $OUT/summary.txt
