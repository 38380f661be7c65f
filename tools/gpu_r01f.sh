#!/bin/bash
# GPU session r01f (short budget): the multigrid set-up tests, config 5 with the hierarchy set up on the
# device, then the whole GPU suite, the bench (both arms) and the ncu launch list.   bash tools/gpu_r01f.sh [tag]
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest tests/test_mg_setup_gpu.py" | tee $OUT/summary.txt
timeout 420 python -m pytest tests/test_mg_setup_gpu.py -m gpu -q -s --timeout 300 > $OUT/pytest_mg_setup.log 2>&1; echo "pytest mg_setup rc=$?" | tee -a $OUT/summary.txt
grep -E "bit-identical|^setup L=|passed|failed|Error|assert" $OUT/pytest_mg_setup.log | tail -40 | tee -a $OUT/summary.txt
echo "== config 5, set-up on the device (2048^2, m = 0.1)" | tee -a $OUT/summary.txt
timeout 420 python tools/bench_mg.py 2048 0.1 --numpy-setup > $OUT/mg_2048_m0.1.jsonl 2> $OUT/mg_2048_m0.1.err; echo "bench_mg rc=$?" | tee -a $OUT/summary.txt
cut -c1-600 $OUT/mg_2048_m0.1.jsonl | tee -a $OUT/summary.txt; tail -3 $OUT/mg_2048_m0.1.err | tee -a $OUT/summary.txt
echo "== pytest -m gpu (everything else)" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_mg_setup_gpu.py > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
echo "== smoke" | tee -a $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
echo "== bench (default)" | tee -a $OUT/summary.txt
timeout 420 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cut -c1-900 $OUT/bench.json | tee -a $OUT/summary.txt; tail -3 $OUT/bench.err | tee -a $OUT/summary.txt
echo "== bench --impl reference" | tee -a $OUT/summary.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; echo "bench ref rc=$?" | tee -a $OUT/summary.txt
cut -c1-400 $OUT/bench_ref.json | tee -a $OUT/summary.txt
echo "== ncu launch list" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --apply-reps 5 > $OUT/ncu_bench.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/summary.txt
ls -la $OUT | tee -a $OUT/summary.txt
