#!/bin/bash
# one 1-GPU session on the B200 box: parity tests, bench, ncu launch list + full capture of the step kernel.
#   bash tools/gpu_session.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-s01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
if [ "${2:-tests}" = "tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest.log
fi
timeout 900 python bench.py --steps 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 7000 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -c 1500 $OUT/bench_ref.json
if [ "${3:-ncu}" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
     python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-configs --apply-reps 5 > $OUT/ncu_bench.log 2>&1; echo "ncu list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cg_step_kernel" -s 2 -c 1 \
     -o $OUT/prof_cgstep python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-configs --apply-reps 5 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls -la $OUT
