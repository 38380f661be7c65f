#!/bin/bash
OUT=gpurun_out/${1:-d01}; mkdir -p $OUT
for bind in 1 0; do
  if [ $bind = 0 ]; then export BENCH_NO_BIND=1; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$bind \
     bench.py --gpus 2 --steps 1 --L 1024 --no-strong > $OUT/b$bind.json 2> $OUT/b$bind.err
  echo "bind=$bind rc=$?"; grep -v "^W\|^\[W\|^\*\|OMP_NUM" $OUT/b$bind.err | head -8; tail -c 1500 $OUT/b$bind.json
done
