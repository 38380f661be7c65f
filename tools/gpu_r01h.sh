#!/bin/bash
# GPU session r01h: timings and one ncu --set full capture of the set-up kernels and the preconditioned stencil passes
TAG=${1:-r01h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== prof_setup timings (not under a profiler)" | tee $OUT/summary.txt
timeout 300 python tools/prof_setup.py 2048 > $OUT/prof_setup.jsonl 2> $OUT/prof_setup.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/prof_setup.jsonl | tee -a $OUT/summary.txt; tail -3 $OUT/prof_setup.err | tee -a $OUT/summary.txt
echo "== ncu full: set-up kernels + partial-apply kernel" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mg_block|mg_galerkin|mg_partition|mg_interleave|coarse_part|coarse_sign" -c 24 \
   -o $OUT/prof_setup python tools/prof_setup.py 2048 > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/summary.txt
ncu -i $OUT/prof_setup.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size > $OUT/prof_setup_raw.csv 2>> $OUT/ncu_full.log
head -40 $OUT/prof_setup_raw.csv | cut -c1-400 | tee -a $OUT/summary.txt
ls -la $OUT | tee -a $OUT/summary.txt
