#!/bin/bash
TAG=${1:-p02}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --apply-reps 5 > $OUT/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"normal_kernel|cg_update_kernel" -s 20 -c 4 \
   -o $OUT/prof_top python bench.py --steps 1 --warmup 3 --no-cpu --apply-reps 5 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.json
ls -la $OUT
