#!/bin/bash
# ncu evidence for the hot kernel (1 GPU).  Plain run first (must exit 0), then the launch list,
# then one --set full capture of the all-pairs screen.  Usage: bash tools/gpu_profile.sh <tag> [variant]
TAG=${1:-r01}
VAR=${2:-tf32}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KRE=rmsd_tf32_kernel
[ "$VAR" != "tf32" ] && KRE=rmsd_sim_kernel
CMD="python bench.py --steps 2 --warmup 3 --skip-extras --variant $VAR"
$CMD > $OUT/plain.json 2> $OUT/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee -a $OUT/rc.txt
$CMD > $OUT/plain2.json 2> $OUT/plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -o $OUT/prof_$VAR $CMD > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?" | tee -a $OUT/rc.txt
ls -la $OUT
