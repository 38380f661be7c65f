#!/bin/bash
# GPU session r01i: paired thread mapping of the partial-apply kernel -- parity, then timings
TAG=${1:-r01i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (partial applies, preconditioned stencil paths, set-up with preconditioned null solves)" | tee $OUT/summary.txt
timeout 600 python -m pytest tests/test_eo_gpu.py tests/test_stencil_prec_gpu.py tests/test_mg_setup_gpu.py -m gpu -q --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -6 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== prof_setup timings" | tee -a $OUT/summary.txt
timeout 300 python tools/prof_setup.py 2048 > $OUT/prof_setup.jsonl 2> $OUT/prof_setup.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/prof_setup.jsonl | tee -a $OUT/summary.txt; tail -3 $OUT/prof_setup.err | tee -a $OUT/summary.txt
echo "== ncu: partial-apply kernel" | tee -a $OUT/summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"coarse_part" -c 6 \
   -o $OUT/prof_part python tools/prof_setup.py 2048 > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/summary.txt
ncu -i $OUT/prof_part.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size > $OUT/prof_part_raw.csv 2>> $OUT/ncu_full.log
cut -c1-400 $OUT/prof_part_raw.csv | tee -a $OUT/summary.txt
