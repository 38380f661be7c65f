#!/bin/bash
# 1-GPU session: parity tests, bench configs (device-resident BiCGStab / CR), config-5 profile.  bash tools/gpu_krylov.sh <tag> [full]
TAG=${1:-k01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest.log
if [ "${2:-quick}" = "full" ]; then
  timeout 900 python bench.py --steps 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
else
  timeout 600 python bench.py --steps 3 --no-cpu --no-strong > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
fi
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d.get("value"), "ms", d.get("ms_per_step"), "frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("value"))
for c in d.get("configs", []):
    print(json.dumps(c))
PY
tail -3 $OUT/bench.err
timeout 600 python tools/bench_mg.py 2048 0.1 > $OUT/mg_2048_m0.1.jsonl 2> $OUT/mg.err; echo "mg rc=$?"
cut -c1-2500 $OUT/mg_2048_m0.1.jsonl; tail -3 $OUT/mg.err
