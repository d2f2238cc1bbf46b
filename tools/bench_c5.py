#!/usr/bin/env python
"""BASELINE configs[4]: 1M generated trimolecular poses -> rotation transforms + clash screen + RMSD pruning,
on 1..8 GPUs (run under torchrun for N > 1).  Prints one JSON line (rank 0).  The clash verdicts of a sample
and the prune mask are checked against the oracle (rank 0, outside the timed region)."""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tscode_b200.embeds import screen_and_prune  # noqa: E402
from tscode_b200.synth import gen_poses, mask_digest  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
P = int(os.environ.get("C5_P", "1000000"))
frags, conf, R, t = gen_poses(2, P, (50, 50, 50))
Rp, tp = torch.from_numpy(R).pin_memory(), torch.from_numpy(t).pin_memory()
cp = torch.from_numpy(conf.astype(np.int32)).pin_memory()
atomnos = np.full(150, 6)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(2):
    res = screen_and_prune(frags, cp, Rp, tp, atomnos, 1.5, 0, 0.5, rank=rank, world=world)
ts = []
for _ in range(5):
    barrier(); t0 = time.perf_counter()
    res = screen_and_prune(frags, cp, Rp, tp, atomnos, 1.5, 0, 0.5, rank=rank, world=world)
    barrier(); ts.append(time.perf_counter() - t0)
dt = torch.tensor([min(ts)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
v, m = res["verdict"].cpu().numpy(), res["mask"].cpu().numpy()
if rank == 0:
    from oracle import oracle_c
    sel = np.arange(0, P, max(P // 5000, 1))
    ok_clash = bool(np.array_equal(v[sel], oracle_c.embed_clash_batch(frags, conf[sel], R[sel], t[sel], 1.5, 0)))
    line = {"config": f"C5 end-to-end {P} trimolecular poses (3 x 50 atoms): transform + clash -> gather -> RMSD prune",
            "n_gpus": world, "ms_end_to_end_incl_h2d": float(dt[0]) * 1e3, "poses_per_s": P / float(dt[0]),
            "phase_ms_rank0": res["ms"], "clash_pass": int(v.sum()), "clash_digest": mask_digest(v),
            "clash_sample_matches_oracle": ok_clash, "prune_survivors": int(m.sum()), "prune_digest": mask_digest(m),
            "h2d_bytes_per_rank": int((R.nbytes + t.nbytes + conf.size * 4) / world)}
    if os.environ.get("C5_CHECK_PRUNE", "1") == "1":
        mref, _, _ = oracle_c.prune_heavy(res["poses"].cpu().numpy(), 0.5)
        line["prune_mask_matches_oracle"] = bool(np.array_equal(m, mref))
    print(json.dumps(line), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
