#!/bin/bash
# GPU session r01j: wide loads in the partial-apply kernel -- parity, timings; then the whole suite + bench on the final code
TAG=${1:-r01j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== prof_setup timings" | tee $OUT/summary.txt
timeout 300 python tools/prof_setup.py 2048 > $OUT/prof_setup.jsonl 2> $OUT/prof_setup.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/prof_setup.jsonl | tee -a $OUT/summary.txt; tail -3 $OUT/prof_setup.err | tee -a $OUT/summary.txt
echo "== pytest -m gpu (all)" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -12 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
echo "== bench (default)" | tee -a $OUT/summary.txt
timeout 420 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cut -c1-300 $OUT/bench.json | tee -a $OUT/summary.txt; tail -3 $OUT/bench.err | tee -a $OUT/summary.txt
