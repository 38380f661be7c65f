#!/bin/bash
TAG=${1:-p01}; NG=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 tools/p2p_bench.py 80 2>/dev/null | grep -v "^W\|^\[W\|^\*\|OMP" | tee $OUT/p2p_bench.txt
nvidia-smi nvlink -s -i 0 2>&1 | head -8 > $OUT/nvlink_status.txt
