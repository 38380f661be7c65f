#!/bin/bash
# MG parity on the GPU + config 5 end to end.   bash tools/gpu_mg.sh <tag> [L] [masses...]
TAG=${1:-g01}; L=${2:-2048}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_mg_gpu.py -m gpu -q --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/summary.txt
tail -12 $OUT/pytest.log | tee -a $OUT/summary.txt
for m in "$@"; do
  timeout 1500 python tools/bench_mg.py $L $m > $OUT/mg_${L}_m$m.jsonl 2> $OUT/mg_${L}_m$m.err; echo "mg $L m=$m rc=$?" | tee -a $OUT/summary.txt
  cat $OUT/mg_${L}_m$m.jsonl | cut -c1-700 | tee -a $OUT/summary.txt; tail -3 $OUT/mg_${L}_m$m.err | tee -a $OUT/summary.txt
done
