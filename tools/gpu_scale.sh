#!/bin/bash
# Multi-GPU parity + scaling of bench.py on one box with 8 GPUs.  Usage: bash tools/gpu_scale.sh <tag> [list of N, default "8 4 2"]
TAG=${1:-scale}
NS=${2:-"8 4 2"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29518 tools/mgpu_check.py > $OUT/mgpu_check_8.log 2>&1
echo "mgpu_check 8 rc=$?" | tee -a $OUT/rc.txt
grep -cE " OK" $OUT/mgpu_check_8.log; grep -E "MISMATCH|Error" $OUT/mgpu_check_8.log | head -3
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --skip-extras > $OUT/bench_1.json 2> $OUT/bench_1.err; echo "bench 1 rc=$?" | tee -a $OUT/rc.txt
for N in $NS; do
  timeout 600 $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_$N.json 2> $OUT/bench_$N.err
  echo "bench $N rc=$?" | tee -a $OUT/rc.txt
done
TAG=$TAG python - <<'PY'
import json, os
tag = os.environ["TAG"]
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{tag}/bench_{n}.json") if l.startswith("{")][-1])
        ws = d.get("weak_scaling") or {}
        print(n, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "phases", {k: round(v, 3) for k, v in d["phase_ms"].items()},
              "e2e ms", round(d["e2e"]["ms_per_call"], 2) if d.get("e2e") else None, d["parity"]["matches_reference"],
              "weak ms", ws.get("ms_per_step"), ws.get("matches_c_oracle_digest"),
              "clash", (d.get("clash") or {}).get("value"), "C5 ms", ((d.get("configs") or {}).get("C5_embed_pipeline") or {}).get("ms_total"))
    except Exception as e:
        print(n, "ERR", e)
PY
