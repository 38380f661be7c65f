#!/bin/bash
TAG=${1:-t04}; OUT=gpurun_out/$TAG; mkdir -p $OUT
GLB_CGSTEP_GFAC=1.5 GLB_CGSTEP_TRACE=$OUT/trace_persist.txt timeout 200 python tools/tune_cgstep.py 4096 508433 > $OUT/trace.log 2>&1
tail -2 $OUT/trace.log | cut -c1-200
