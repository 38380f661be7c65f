#!/bin/bash
# 1-GPU session: parity of the fused BiCGStab inputs / the CR dot fold, then the fused-vs-unfused sweep.  bash tools/gpu_fuse.sh <tag> [sizes...]
TAG=${1:-f01}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_krylov_gpu.py tests/test_solvers_gpu.py tests/test_family_gpu.py tests/test_apply_gpu.py \
   tests/test_baseline_sizes_gpu.py tests/test_mg_setup_gpu.py -m gpu -q --timeout 600 -s > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "bit-identical|fused ==|passed|failed|Error|assert" $OUT/pytest.log | tail -30
timeout 600 python tools/tune_krylov.py "$@" > $OUT/tune_krylov.jsonl 2> $OUT/tune.err; echo "tune rc=$?"; cat $OUT/tune_krylov.jsonl; tail -3 $OUT/tune.err
