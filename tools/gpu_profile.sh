#!/bin/bash
# ncu evidence for the hot kernel (1 GPU).  Plain run first (must exit 0), then the launch list,
# then one --set full capture of the all-pairs screen.
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --skip-extras"
$CMD > $OUT/plain.json 2> $OUT/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee -a $OUT/rc.txt
$CMD > $OUT/plain2.json 2> $OUT/plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:rmsd_sim_kernel -s 3 -c 1 -o $OUT/prof_sim $CMD > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?" | tee -a $OUT/rc.txt
ls -la $OUT
