#!/bin/bash
# kernel-variant sweep on one GPU: bash tools/tune.sh <tag>
TAG=${1:-t01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --steps 5 --apply-reps 30 > $OUT/$name.json 2> $OUT/$name.err
  python - "$name" "$OUT/$name.json" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-28s solve %.2f ms  %4d it  value %.0f GB/s (%.1f%%)  apply %.4f ms %.0f GB/s (%.1f%%)  e2e %.0f" % (sys.argv[1], j["ms_per_step"], j["config"]["iterations"], j["value"], 100*j["frac_of_hbm_peak"], j["roofline"]["ms_per_launch"], j["roofline"]["achieved"], 100*j["roofline"]["frac"], j["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run fused_pf1   GLB_NORMAL_FUSED=1 GLB_STAG_PF=1 | tee -a $OUT/summary.txt
run fused_pf2   GLB_NORMAL_FUSED=1 GLB_STAG_PF=2 | tee -a $OUT/summary.txt
run twopass_pf1 GLB_NORMAL_FUSED=0 GLB_STAG_PF=1 | tee -a $OUT/summary.txt
run twopass_pf2 GLB_NORMAL_FUSED=0 GLB_STAG_PF=2 | tee -a $OUT/summary.txt
run twopass_spt1 GLB_NORMAL_FUSED=0 GLB_STAG_PF=1 GLB_STAG_SPT=1 | tee -a $OUT/summary.txt
run twopass_spt1_pf2 GLB_NORMAL_FUSED=0 GLB_STAG_PF=2 GLB_STAG_SPT=1 | tee -a $OUT/summary.txt
