#!/bin/bash
TAG=${1:-t04}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --steps 5 --apply-reps 30 > $OUT/$name.json 2> $OUT/$name.err
  python - "$name" "$OUT/$name.json" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-22s solve %.2f ms %4d it value %.0f (%.1f%%) apply %.4f ms (%.1f%%) normal %.4f ms (%.1f%%) e2e %.0f actual %.3f" % (sys.argv[1], j["ms_per_step"], j["config"]["iterations"], j["value"], 100*j["frac_of_hbm_peak"], j["roofline"]["ms_per_launch"], 100*j["roofline"]["frac"], j["roofline"]["other_kernels"][0]["ms_per_launch"], 100*j["roofline"]["other_kernels"][0]["frac"], j["e2e"]["value"], j["config"]["actual_traffic_frac_of_peak"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for st in 0 3 4 6 4 0; do run stages_$st GLB_NORMAL_STAGES=$st | tee -a $OUT/summary.txt; done
