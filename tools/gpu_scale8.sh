#!/bin/bash
# 8-GPU parity + bench + C5 only (reduced form of gpu_scale.sh).
TAG=${1:-r02s8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29518 tools/mgpu_check.py > $OUT/mgpu_check_8.log 2>&1
echo "mgpu_check 8 rc=$?" | tee -a $OUT/rc.txt
grep -cE " OK " $OUT/mgpu_check_8.log; grep -E "MISMATCH|Error" $OUT/mgpu_check_8.log | head -3
timeout 300 $TR --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/bench_8.json 2> $OUT/bench_8.err
echo "bench 8 rc=$?" | tee -a $OUT/rc.txt
C5_CHECK_PRUNE=0 timeout 300 $TR --nproc-per-node 8 --master-port 29538 tools/bench_c5.py > $OUT/c5_8.json 2> $OUT/c5_8.err
echo "c5 8 rc=$?" | tee -a $OUT/rc.txt
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_8.json") if l.startswith("{")][-1])
print(8, "ms/step %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["phase_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_call"], 2), d["parity"]["matches_reference"])
d = json.loads([l for l in open("$OUT/c5_8.json") if l.startswith("{")][-1])
print("C5 8 ms %.2f" % d["ms_end_to_end_incl_h2d"], d["phase_ms_rank0"], d["clash_digest"], d["prune_digest"])
PY
