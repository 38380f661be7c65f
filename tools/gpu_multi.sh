#!/bin/bash
# multi-GPU session: slab parity + weak scaling (4096^2 per GPU) + strong scaling (config 4: 8192^2 global).
#   bash tools/gpu_multi.sh <tag> <ngpus> [strong-too: 1|0]
TAG=${1:-m01}; NG=${2:-2}; STRONG=${3:-1}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
echo "== slab check x$NG" | tee $OUT/summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
   tools/slab_check.py 64 > $OUT/slab_check.log 2>&1; echo "slab rc=$?" | tee -a $OUT/summary.txt
grep -E "ok|FAIL|SLAB" $OUT/slab_check.log | tee -a $OUT/summary.txt; tail -5 $OUT/slab_check.log >> $OUT/summary.txt
run_bench () {  # name n extra-args
  local name=$1 n=$2; shift 2
  echo "== bench $name N=$n" | tee -a $OUT/summary.txt
  if [ $n -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 3 --no-cpu "$@" > $OUT/${name}_n$n.json 2> $OUT/${name}_n$n.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
      bench.py --gpus $n --steps 3 "$@" > $OUT/${name}_n$n.json 2> $OUT/${name}_n$n.err
  fi
  echo "rc=$?" | tee -a $OUT/summary.txt
  tail -1 $OUT/${name}_n$n.json | cut -c1-1200 | tee -a $OUT/summary.txt; tail -3 $OUT/${name}_n$n.err | tee -a $OUT/summary.txt
}
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then run_bench bench $n; fi
done
if [ "$STRONG" = "1" ]; then
  for n in 1 2 4 8; do
    if [ $n -le $NG ]; then run_bench strong8192 $n --L 8192 --Y 8192 --apply-reps 10; fi
  done
fi
