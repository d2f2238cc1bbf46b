#!/usr/bin/env python
"""Peer-written pair lists against the NCCL all-gather (run under torchrun, one rank per GPU): same masks, and the time
of the ladder phase with either exchange.  python -m torch.distributed.run --nproc-per-node N tools/p2p_check.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200 import rmsd_pruning as rp  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
S = gen_ensemble(3, 50000, 80, 5000)
Sd = torch.from_numpy(S).cuda()
atomnos = np.full(80, 6)
out = {}
for mode in ("p2p", "nccl"):
    rp._PeerLists.enabled = mode == "p2p"
    rp._PeerLists._cache.clear()
    pr = rp.RmsdPruner(Sd, atomnos, 0.5, rank=rank, world=world, group=dist.group.WORLD)
    used = pr._peer is not None
    for _ in range(3):
        m = pr.run()
    ts, te = [], []
    for _ in range(10):
        pr.pack(); pr.similarity()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        m = pr.eliminate()
        torch.cuda.synchronize()
        te.append((time.perf_counter() - t0) * 1e3)
        dist.barrier(); t1 = time.perf_counter()
        m = pr.run(); torch.cuda.synchronize(); dist.barrier()
        ts.append((time.perf_counter() - t1) * 1e3)
    out[mode] = (mask_digest(m.cpu().numpy()), used, min(te), sorted(te)[5], min(ts), sorted(ts)[5], getattr(pr._peer, "err", None) if pr._peer else rp._PeerLists._cache and list(rp._PeerLists._cache.values())[0].err)
    if rank == 0:
        d, u, e0, e1, s0, s1, err = out[mode]
        print(f"world={world} {mode}: peer lists used={u} digest={d} {'OK' if d == '478bc29df1e239da' else 'MISMATCH'} "
              f"ladder phase min {e0:.3f} median {e1:.3f} ms; whole step (host clock) min {s0:.3f} median {s1:.3f} ms; err={err}", flush=True)
dist.barrier()
dist.destroy_process_group()
