#!/bin/bash
# final 1-GPU session of the round: the whole -m gpu suite, smoke, both bench arms, ncu launch list of the bench and
# a full capture of one Gram-Schmidt kernel launch inside the config-5 solve.   bash tools/gpu_final.sh <tag> [ncu|noncu]
TAG=${1:-z01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python bench.py --steps 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d.get("value"), "ms", d.get("ms_per_step"), "frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("value"),
      "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
for c in d.get("configs", []):
    print(json.dumps(c)[:400])
print(json.dumps(d.get("strong"))[:600])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -c 600 $OUT/bench_ref.json
if [ "${2:-ncu}" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
     python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --apply-reps 5 > $OUT/ncu_bench.log 2>&1; echo "ncu list rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"multi_dot_kernel" -s 400 -c 1 \
     -o $OUT/prof_multi_dot python tools/bench_mg.py 2048 0.1 > $OUT/ncu_md.log 2>&1; echo "ncu multi_dot rc=$?"
fi
ls -la $OUT
