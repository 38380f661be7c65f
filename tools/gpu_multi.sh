#!/bin/bash
# multi-GPU session: slab parity, per-step wait trace, then bench.py (weak scaling line + strong block) at N ranks.
#   bash tools/gpu_multi.sh <tag> <ngpus> [bench extra args...]
TAG=${1:-m01}; NG=${2:-2}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
echo "== slab check x$NG" | tee $OUT/summary.txt
timeout 600 $TR --master-port 29511 tools/slab_check.py 64 > $OUT/slab_check.log 2>&1; echo "slab rc=$?" | tee -a $OUT/summary.txt
grep -E "FAIL|SLAB|single-kernel|solve CGNE" $OUT/slab_check.log | tee -a $OUT/summary.txt; tail -5 $OUT/slab_check.log >> $OUT/summary.txt
echo "== per-step wait traces" | tee -a $OUT/summary.txt
nvidia-smi nvlink -gt d -i 0 > $OUT/nvlink_before.txt 2>&1
for shape in "4096 $((4096*NG))" "4096 4096"; do
  timeout 300 $TR --master-port 29512 tools/slab_trace.py $shape $OUT/traces 2>/dev/null | grep "^rank" | tee -a $OUT/summary.txt
done
nvidia-smi nvlink -gt d -i 0 > $OUT/nvlink_after.txt 2>&1
echo "== bench N=$NG" | tee -a $OUT/summary.txt
timeout 900 $TR --master-port 29521 bench.py --gpus $NG --steps 5 "$@" > $OUT/bench_n$NG.json 2> $OUT/bench_n$NG.err
echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/bench_n$NG.json | cut -c1-7000 | tee -a $OUT/summary.txt; grep -v "^W\|^\[W\|^\*\|OMP_NUM" $OUT/bench_n$NG.err | tail -5 | tee -a $OUT/summary.txt
rm -f $OUT/traces/*rank[1-9]*.txt   # keep rank 0's per-CTA record and everybody's per-step stamps
