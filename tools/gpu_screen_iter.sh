#!/bin/bash
# One screen-kernel iteration on the GPU box: small-case parity for every mode, timeline, kernel timings, short bench.
# usage: bash tools/gpu_screen_iter.sh <tag>
TAG=${1:-it}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 200 python tools/screen_check.py screen --no-big > $OUT/check.log 2>&1; echo "check rc=$? bad=$(grep -c '^BAD\|Traceback' $OUT/check.log) ok=$(grep -c '^ok' $OUT/check.log)"
timeout 120 python tools/screen_trace.py 30000 80 0,3 > $OUT/trace.log 2>&1; tail -1 $OUT/trace.log
timeout 200 python tools/profile_kernels.py > $OUT/prof.json 2>&1; grep -E "screen_ms" $OUT/prof.json
timeout 300 python bench.py --steps 5 --warmup 3 --skip-extras > $OUT/bench.json 2> $OUT/bench.err; echo bench rc=$?
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d.get("phase_ms"), "frac", d["roofline"]["frac"], d["parity"]["matches_reference"])
PY
