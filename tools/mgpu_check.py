#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-sharded prune with an
NCCL all-gather per elimination round must give the single-GPU / live-reference mask."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
gold = json.load(open("tests/golden/prune_masks.json"))["rows"] + json.load(open("tests/golden/prune_masks_big.json"))["rows"]
ok = True
for r in gold:
    if r.get("mixed_h") or r["N"] < 64:
        continue
    S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"])
    pr = RmsdPruner(S, np.full(r["M"], 6), r["thr"], rank=rank, world=world)
    mask = pr.run().cpu().numpy()
    good = mask_digest(mask) == r["digest"]
    ok &= good
    if rank == 0:
        print(f"world={world} N={r['N']} M={r['M']} survivors={int(mask.sum())} digest={mask_digest(mask)} "
              f"{'OK' if good else 'MISMATCH'} rounds={pr.rounds}", flush=True)
t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t[0]) == 1 else 1)
