#!/bin/bash
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --skip-extras --variant tf32"
$CMD > $OUT/plain.json 2> $OUT/plain.err && \
ncu --set full --clock-control none --import-source on -k regex:rmsd_tf32_kernel -s 3 -c 1 -o $OUT/prof_tf32 $CMD > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?" | tee -a $OUT/rc.txt
tail -c 600 $OUT/plain.json
