#!/bin/bash
# One gpurun call with everything a round's evidence needs (one GPU): smoke -> parity tests -> bench + reference arm ->
# launch list -> ncu full capture of the screen inside bench.py's step and of every hot kernel (profile_kernels.py) ->
# timeline and shape probes.  Every ncu pass runs only after the same command has exited 0 without ncu.
# Usage (repo root on the GPU box): bash tools/gpu_round.sh <tag>
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?" | tee -a $OUT/rc.txt
tail -2 $OUT/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -3 $OUT/pytest_gpu.log
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err ; echo "bench ref rc=$?" | tee -a $OUT/rc.txt
echo "== bench" ; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" | tee -a $OUT/rc.txt
tail -c 300 $OUT/bench.json
CMD="python bench.py --steps 2 --warmup 3 --skip-extras"
$CMD > $OUT/bench_plain.json 2> $OUT/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee -a $OUT/rc.txt
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rmsd_screen_kernel -c 2 -o $OUT/screen_c3 $CMD > $OUT/ncu_screen.log 2>&1
echo "screen capture rc=$?" | tee -a $OUT/rc.txt
K="python tools/profile_kernels.py"
$K > $OUT/profile_kernels_plain.json 2> $OUT/profile_kernels.err && \
ncu --set full --clock-control none --import-source on -k regex:"rmsd_screen_kernel|rmsd_verify_list|elim_fused|embed_clash_kernel|rotcorr_scan" -c 14 -o $OUT/prof_kernels $K > $OUT/ncu_kernels.log 2>&1
echo "kernel captures rc=$?" | tee -a $OUT/rc.txt
echo "== probes"
timeout 120 python tools/screen_trace.py 30000 80 0,1,2 > $OUT/screen_trace.log 2>&1 ; grep steady $OUT/screen_trace.log
timeout 300 python tools/screen_check.py screen > $OUT/screen_check.log 2>&1 ; grep -E "small cases|screen mode" $OUT/screen_check.log | cut -c1-140
timeout 120 python tools/clash_run.py > $OUT/clash_run.log 2>&1 ; tail -2 $OUT/clash_run.log
# gpurun brings back at most 64 MiB: summaries are made here, and the big report is dropped if the total would not fit
python tools/ncu_summary.py $OUT/screen_c3.ncu-rep > $OUT/ncu_screen_c3.md 2>/dev/null
python tools/ncu_summary.py $OUT/prof_kernels.ncu-rep > $OUT/ncu_kernels.md 2>/dev/null
[ $(du -sm gpurun_out | cut -f1) -gt 60 ] && rm -f $OUT/prof_kernels.ncu-rep
ls -la $OUT
