#!/bin/bash
# 2-GPU session r01l: slab parity including the stencil-side preconditioned paths, even and odd slab offsets
TAG=${1:-r01l}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for L in 64 66; do
  echo "== slab check x2, L=$L" | tee -a $OUT/summary.txt
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((L % 10)) \
     tools/slab_check.py $L > $OUT/slab_check_$L.log 2>&1; echo "slab rc=$?" | tee -a $OUT/summary.txt
  grep -E " ok | FAIL|SLAB|Error|error" $OUT/slab_check_$L.log | cut -c1-160 | tee -a $OUT/summary.txt
done
