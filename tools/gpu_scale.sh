#!/bin/bash
# 1/2/4/8-GPU scaling of bench.py + multi-rank parity + C5 end-to-end (one box with 8 GPUs).
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29518 tools/mgpu_check.py > $OUT/mgpu_check_8.log 2>&1
echo "mgpu_check 8 rc=$?" | tee -a $OUT/rc.txt
grep -cE " OK " $OUT/mgpu_check_8.log; grep -E "MISMATCH|Error" $OUT/mgpu_check_8.log | head -3
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $OUT/bench_1.json 2> $OUT/bench_1.err; echo "bench 1 rc=$?" | tee -a $OUT/rc.txt
for N in 2 4 8; do
  timeout 600 $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_$N.json 2> $OUT/bench_$N.err
  echo "bench $N rc=$?" | tee -a $OUT/rc.txt
done
C5_CHECK_PRUNE=1 timeout 600 python tools/bench_c5.py > $OUT/c5_1.json 2> $OUT/c5_1.err; echo "c5 1 rc=$?" | tee -a $OUT/rc.txt
for N in 2 4 8; do
  C5_CHECK_PRUNE=0 timeout 600 $TR --nproc-per-node $N --master-port 2953$N tools/bench_c5.py > $OUT/c5_$N.json 2> $OUT/c5_$N.err
  echo "c5 $N rc=$?" | tee -a $OUT/rc.txt
done
TAG=$TAG python - <<'PY'
import json, os
tag = os.environ["TAG"]
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{tag}/bench_{n}.json") if l.startswith("{")][-1])
        print(n, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "phases", {k: round(v, 3) for k, v in d["phase_ms"].items()},
              "e2e ms", round(d["e2e"]["ms_per_call"], 2) if d.get("e2e") else None, d["parity"]["matches_reference"], d["config"].get("ladder"))
    except Exception as e:
        print(n, "ERR", e)
    try:
        d = json.loads([l for l in open(f"gpurun_out/{tag}/c5_{n}.json") if l.startswith("{")][-1])
        print("  C5", n, "ms %.2f" % d["ms_end_to_end_incl_h2d"], d["phase_ms_rank0"], d["clash_digest"], d["prune_digest"], d.get("prune_mask_matches_oracle"))
    except Exception as e:
        print("  C5", n, "ERR", e)
PY
