#!/bin/bash
# 1/2/4/8-GPU scaling of bench.py + multi-rank parity (one box with 8 GPUs).
TAG=${1:-r01s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N tools/mgpu_check.py > $OUT/mgpu_check_$N.log 2>&1
  echo "mgpu_check $N rc=$?" | tee -a $OUT/rc.txt
  grep -cE " OK " $OUT/mgpu_check_$N.log; grep -E "MISMATCH|Error" $OUT/mgpu_check_$N.log | head -3
done
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $OUT/bench_1.json 2> $OUT/bench_1.err; echo "bench 1 rc=$?" | tee -a $OUT/rc.txt
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_$N.json 2> $OUT/bench_$N.err
  echo "bench $N rc=$?" | tee -a $OUT/rc.txt
done
TAG=$TAG python - <<'PY'
import json, os
tag = os.environ["TAG"]
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{tag}/bench_{n}.json") if l.startswith("{")][-1])
        print(n, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "phases", {k: round(v, 3) for k, v in d["phase_ms"].items()},
              "e2e ms", round(d["e2e"]["ms_per_call"], 2) if d.get("e2e") else None, d["parity"]["matches_reference"])
    except Exception as e:
        print(n, "ERR", e)
PY
