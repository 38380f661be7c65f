#!/bin/bash
# GPU session r01o: final check of the round -- whole GPU suite, smoke, default bench
TAG=${1:-r01o}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu (all)" | tee $OUT/summary.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | cut -c1-220 | tee -a $OUT/summary.txt
echo "== smoke" | tee -a $OUT/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/smoke.log | tee -a $OUT/summary.txt
echo "== bench" | tee -a $OUT/summary.txt
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cut -c1-260 $OUT/bench.json | tee -a $OUT/summary.txt
