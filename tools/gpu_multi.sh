#!/bin/bash
# Multi-GPU parity + bench (one box, N GPUs).  Usage: bash tools/gpu_multi.sh <N> [tag]
N=${1:-2}
TAG=${2:-r02m$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR tools/mgpu_check.py > $OUT/mgpu_check.log 2>&1 ; echo "mgpu_check rc=$?" | tee -a $OUT/rc.txt
grep -E "world=|MISMATCH|Error" $OUT/mgpu_check.log | tail -20
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_$N.json 2> $OUT/bench_$N.err ; echo "bench rc=$?" | tee -a $OUT/rc.txt
tail -c 2500 $OUT/bench_$N.json
