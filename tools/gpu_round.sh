#!/bin/bash
# One gpurun call: smoke -> parity tests -> bench -> launch list.  Logs land in gpurun_out/.
# Usage (from repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?" | tee -a $OUT/rc.txt
tail -3 $OUT/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" | tee -a $OUT/rc.txt
tail -5 $OUT/pytest_gpu.log
echo "== bench tf32 (default)" ; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_tf32.json 2> $OUT/bench_tf32.err ; echo "bench tf32 rc=$?" | tee -a $OUT/rc.txt
tail -c 600 $OUT/bench_tf32.json
echo "== bench dmma" ; timeout 900 python bench.py --steps 3 --warmup 3 --variant dmma --no-cpu > $OUT/bench_dmma.json 2> $OUT/bench_dmma.err ; echo "bench dmma rc=$?" | tee -a $OUT/rc.txt
tail -c 300 $OUT/bench_dmma.json
echo "== bench fma" ; timeout 900 python bench.py --steps 3 --warmup 3 --variant fma --no-cpu > $OUT/bench_fma.json 2> $OUT/bench_fma.err ; echo "bench fma rc=$?" | tee -a $OUT/rc.txt
tail -c 300 $OUT/bench_fma.json
