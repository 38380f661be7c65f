#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (both arms), per-config measurements, ncu launch list +
# full capture of the step's kernels.   Usage (repo root, on the GPU box):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== smoke" | tee $OUT/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
echo "== bench (default)" | tee -a $OUT/summary.txt
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt; tail -5 $OUT/bench.err | tee -a $OUT/summary.txt
echo "== bench --impl reference" | tee -a $OUT/summary.txt
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err ) 2>&1 | grep real | tee -a $OUT/summary.txt
cat $OUT/bench_ref.json | tee -a $OUT/summary.txt
echo "== per-config measurements" | tee -a $OUT/summary.txt
timeout 1200 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/configs.err | tee -a $OUT/summary.txt
echo "== ncu launch list" | tee -a $OUT/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --apply-reps 5 > $OUT/ncu_bench.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/summary.txt
echo "== ncu full (normal_kernel, cg_update, stag_kernel)" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"normal_kernel|normal1_kernel|cg_update_kernel|stag_kernel" -s 10 -c 16 \
   -o $OUT/prof_top python bench.py --steps 1 --warmup 3 --no-cpu --apply-reps 5 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/summary.txt
ls -la $OUT | tee -a $OUT/summary.txt
