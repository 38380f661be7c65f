#!/bin/bash
# multi-GPU session: slab parity + scaling bench.   bash tools/gpu_multi.sh <tag> <ngpus>
TAG=${1:-m01}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
echo "== slab check x$NG" | tee $OUT/summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
   tools/slab_check.py 64 > $OUT/slab_check.log 2>&1; echo "slab rc=$?" | tee -a $OUT/summary.txt
grep -E "ok|FAIL|SLAB" $OUT/slab_check.log | tee -a $OUT/summary.txt; tail -5 $OUT/slab_check.log >> $OUT/summary.txt
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    echo "== bench N=$n" | tee -a $OUT/summary.txt
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py --gpus 1 --steps 3 --no-cpu > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
        bench.py --gpus $n --steps 3 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
    fi
    echo "rc=$?" | tee -a $OUT/summary.txt
    tail -1 $OUT/bench_n$n.json | tee -a $OUT/summary.txt; tail -3 $OUT/bench_n$n.err | tee -a $OUT/summary.txt
  fi
done
