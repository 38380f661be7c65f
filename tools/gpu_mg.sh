#!/bin/bash
TAG=${1:-g01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_mg_gpu.py -m gpu -q --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/summary.txt
tail -12 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 900 python tools/bench_mg.py 512 > $OUT/mg_512.jsonl 2> $OUT/mg_512.err; echo "mg 512 rc=$?" | tee -a $OUT/summary.txt
cat $OUT/mg_512.jsonl | cut -c1-700 | tee -a $OUT/summary.txt; tail -3 $OUT/mg_512.err | tee -a $OUT/summary.txt
timeout 1500 python tools/bench_mg.py 2048 > $OUT/mg_2048.jsonl 2> $OUT/mg_2048.err; echo "mg 2048 rc=$?" | tee -a $OUT/summary.txt
cat $OUT/mg_2048.jsonl | cut -c1-700 | tee -a $OUT/summary.txt; tail -3 $OUT/mg_2048.err | tee -a $OUT/summary.txt
