#!/bin/bash
# 2-GPU box: MG parity on one GPU, slab parity + weak/strong bench at N=2 with the warp-parallel peer allreduce
TAG=${1:-c01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_mg_gpu.py tests/test_solvers_gpu.py tests/test_apply_gpu.py -m gpu -q -x --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/summary.txt
tail -15 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   tools/slab_check.py 64 > $OUT/slab_check.log 2>&1; echo "slab rc=$?" | tee -a $OUT/summary.txt
grep -E "FAIL|SLAB" $OUT/slab_check.log | tee -a $OUT/summary.txt; tail -3 $OUT/slab_check.log >> $OUT/summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 \
   bench.py --gpus 2 --steps 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/bench_n2.json | cut -c1-1500 | tee -a $OUT/summary.txt; tail -3 $OUT/bench_n2.err | tee -a $OUT/summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 \
   bench.py --gpus 2 --steps 3 --L 8192 --Y 2048 --apply-reps 10 > $OUT/strong_n2.json 2> $OUT/strong_n2.err; echo "strong-like (1024 rows/GPU) rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/strong_n2.json | cut -c1-1500 | tee -a $OUT/summary.txt; tail -3 $OUT/strong_n2.err | tee -a $OUT/summary.txt
