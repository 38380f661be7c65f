#!/bin/bash
# 1-GPU session: parity of the Gram-Schmidt kernels (multi_dot / lincomb) and config 5 with the per-class profile.
#   bash tools/gpu_gs.sh <tag> [masses...]
TAG=${1:-g01}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_blas_gpu.py tests/test_solvers_gpu.py tests/test_family_gpu.py tests/test_mg_gpu.py \
   tests/test_mg_setup_gpu.py tests/test_krylov_gpu.py -m gpu -q --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest.log
for m in "$@"; do
  timeout 900 python tools/bench_mg.py 2048 $m > $OUT/mg_2048_m$m.jsonl 2> $OUT/mg_m$m.err; echo "mg m=$m rc=$?"
  grep -E '"kind": "solve' $OUT/mg_2048_m$m.jsonl | cut -c1-2600; tail -3 $OUT/mg_m$m.err
done
