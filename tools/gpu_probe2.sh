#!/bin/bash
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python tools/umma_probe.py > $OUT/umma.log 2>&1; tail -30 $OUT/umma.log
CMD="python bench.py --steps 2 --warmup 3 --skip-extras"
$CMD > $OUT/plain.json 2> $OUT/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
for r in rows[-40:]:
    print(r[ik][:60], r[iv])
PY
