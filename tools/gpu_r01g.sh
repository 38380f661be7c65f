#!/bin/bash
# GPU session r01g: the whole GPU suite (new: preconditioned stencil paths, SOR / MinRes, preconditioned null-vector
# solves in the set-up), per-config measurements.   bash tools/gpu_r01g.sh [tag]
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest -m gpu (new modules first, verbose)" | tee $OUT/summary.txt
timeout 600 python -m pytest tests/test_stencil_prec_gpu.py tests/test_mg_setup_gpu.py tests/test_family_gpu.py tests/test_mg_gpu.py -m gpu -q -s --timeout 300 > $OUT/pytest_new.log 2>&1; echo "pytest new rc=$?" | tee -a $OUT/summary.txt
grep -E "^setup L=|passed|failed|Error|^E  |FAILED" $OUT/pytest_new.log | tail -40 | tee -a $OUT/summary.txt
echo "== pytest -m gpu (all)" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -12 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
echo "== smoke" | tee -a $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
echo "== per-config measurements" | tee -a $OUT/summary.txt
timeout 600 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/configs.err | tee -a $OUT/summary.txt
wc -l $OUT/configs.jsonl | tee -a $OUT/summary.txt
ls -la $OUT | tee -a $OUT/summary.txt
