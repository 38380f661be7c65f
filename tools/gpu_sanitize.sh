#!/bin/bash
# compute-sanitizer over the kernels added this round (small cases only).   bash tools/gpu_sanitize.sh <tag>
TAG=${1:-c01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SEL="tests/test_blas_gpu.py::test_multi_vector_ops tests/test_krylov_gpu.py::test_graph_and_direct_batches_agree tests/test_krylov_gpu.py::test_iteration_cap_and_initial_guess tests/test_krylov_gpu.py::test_real_laplace_and_coarse_stencil"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SEL -m gpu -q -x --timeout 450 > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|error" $OUT/memcheck.log | tail -8
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_blas_gpu.py::test_multi_vector_ops -m gpu -q -x --timeout 280 > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" $OUT/racecheck.log | tail -6
