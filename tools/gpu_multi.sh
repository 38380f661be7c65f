#!/bin/bash
# multi-GPU session: slab parity, then bench.py (weak scaling line + strong block) at N ranks.
#   bash tools/gpu_multi.sh <tag> <ngpus> [bench extra args...]
TAG=${1:-m01}; NG=${2:-2}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== slab check x$NG" | tee $OUT/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
   tools/slab_check.py 64 > $OUT/slab_check.log 2>&1; echo "slab rc=$?" | tee -a $OUT/summary.txt
grep -E "FAIL|SLAB|single-kernel|solve CGNE" $OUT/slab_check.log | tee -a $OUT/summary.txt; tail -5 $OUT/slab_check.log >> $OUT/summary.txt
echo "== bench N=$NG" | tee -a $OUT/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 \
   bench.py --gpus $NG --steps 5 "$@" > $OUT/bench_n$NG.json 2> $OUT/bench_n$NG.err
echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/bench_n$NG.json | cut -c1-6000 | tee -a $OUT/summary.txt; tail -5 $OUT/bench_n$NG.err | tee -a $OUT/summary.txt
