#!/bin/bash
# parity tests + bench + launch list (one GPU).  Usage: bash tools/gpu_quick2.sh <tag>
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -12 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" | tee -a $OUT/rc.txt
tail -5 $OUT/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench.json") if l.startswith("{")][-1])
print("ms/step", d["ms_per_step"], d["phase_ms"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_call"], d["parity"], d["config"].get("ladder"))
PY
CMD="python bench.py --steps 2 --warmup 3 --skip-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
for r in rows[-9:]:
    print(r[ik][:60], r[iv])
PY
