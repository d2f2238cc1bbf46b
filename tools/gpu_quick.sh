#!/bin/bash
# Quick single-GPU check: parity tests + default bench (no ncu).  Usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -15 $OUT/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" | tee -a $OUT/rc.txt
tail -5 $OUT/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench.json") if l.startswith("{")][-1])
print("ms/step", d["ms_per_step"], d["phase_ms"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_call"], d["parity"], d["config"].get("ladder"))
PY
