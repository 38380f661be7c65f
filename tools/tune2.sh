#!/bin/bash
TAG=${1:-t02}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { name=$1; shift
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
for pf in 0 2 4 8 16; do run pfl2_$pf GLB_PF_L2=$pf | tee -a $OUT/summary.txt; done
run twopass_pfl2_4 GLB_NORMAL_FUSED=0 GLB_PF_L2=4 | tee -a $OUT/summary.txt
run twopass_pfl2_8 GLB_NORMAL_FUSED=0 GLB_PF_L2=8 | tee -a $OUT/summary.txt
timeout 600 python -m pytest tests/test_apply_gpu.py tests/test_solvers_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -3 | tee -a $OUT/summary.txt
