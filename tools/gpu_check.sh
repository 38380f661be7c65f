#!/bin/bash
# smoke() + compute-sanitizer memcheck of the single-kernel CG on small lattices.   bash tools/gpu_check.sh <tag>
TAG=${1:-c01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_cgstep_gpu.py -x -q -k "first_iterations or max_iter" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|=========" $OUT/memcheck.log | tail -8
