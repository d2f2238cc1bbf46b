#!/bin/bash
# One gpurun call: smoke -> parity tests -> bench (default variant) -> launch list -> ncu full capture
# of the default screen kernel.  Usage (repo root on the GPU box): bash tools/gpu_round.sh <tag>
TAG=${1:-r01t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?" | tee -a $OUT/rc.txt
tail -3 $OUT/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -5 $OUT/pytest_gpu.log
echo "== bench (default)" ; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" | tee -a $OUT/rc.txt
tail -c 400 $OUT/bench.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ; echo "bench ref rc=$?" | tee -a $OUT/rc.txt
CMD="python bench.py --steps 2 --warmup 3 --skip-extras"
$CMD > $OUT/plain.json 2> $OUT/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee -a $OUT/rc.txt
$CMD > $OUT/plain2.json 2> $OUT/plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:rmsd_ts_kernel -s 3 -c 1 -o $OUT/prof_screen $CMD > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?" | tee -a $OUT/rc.txt
ls -la $OUT
echo "== probes" ; timeout 120 python tools/umma_probe.py > $OUT/umma_probe.log 2>&1 ; timeout 120 python tools/trace_probe.py > $OUT/trace_probe.log 2>&1 ; timeout 120 python tools/clash_run.py > $OUT/clash_run.log 2>&1; tail -3 $OUT/clash_run.log
timeout 120 python tools/aniso_probe.py 50000 0 80 isotropic,elongated > $OUT/aniso_probe.log 2>&1 ; grep f16 $OUT/aniso_probe.log
timeout 120 python tools/e2e_timeline.py > $OUT/e2e_timeline.log 2>&1 ; tail -12 $OUT/e2e_timeline.log
