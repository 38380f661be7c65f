#!/bin/bash
# short multi-GPU check: slab parity, then bench.py at N ranks.   bash tools/gpu_multi_quick.sh <tag> <ngpus>
TAG=${1:-q01}; NG=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/slab_check.py 64 > $OUT/slab_check.log 2>&1; echo "slab rc=$?"
grep -E "FAIL|SLAB|solve |MG solve" $OUT/slab_check.log | tail -20
timeout 600 $TR --master-port 29521 bench.py --gpus $NG --steps 3 > $OUT/bench_n$NG.json 2> $OUT/bench_n$NG.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/bench_n$NG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "slab_parity", d["config"].get("slab_parity", d.get("run", {}).get("slab_parity")))
print(json.dumps(d.get("strong"))[:900])
PY
grep -v "^W\|^\[W\|^\*\|OMP_NUM" $OUT/bench_n$NG.err | tail -3
