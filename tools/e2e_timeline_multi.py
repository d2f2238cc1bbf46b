#!/usr/bin/env python
"""Host timeline of the drop-in with a process group (run under torchrun): where the end-to-end time of rank 0 goes.
python -m torch.distributed.run --nproc-per-node N tools/e2e_timeline_multi.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.rmsd_pruning import RmsdPruner, prune_conformers_rmsd  # noqa: E402
from tscode_b200.synth import gen_ensemble  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
S0 = gen_ensemble(3, 50000, 80, 5000)
atomnos = np.full(80, 6)
pin = torch.empty(S0.shape, dtype=torch.float64).pin_memory()
pin.copy_(torch.from_numpy(S0))
S = pin.numpy()
g = dist.group.WORLD
for _ in range(3):
    prune_conformers_rmsd(S, atomnos, 0.5, group=g)
rows = []
for rep in range(8):
    torch.cuda.synchronize(); dist.barrier()
    t = [time.perf_counter()]
    pr = RmsdPruner(S, atomnos, 0.5, rank=rank, world=world, group=g); t.append(time.perf_counter())
    pr.run_async(); t.append(time.perf_counter())
    torch.cuda.current_stream().synchronize(); t.append(time.perf_counter())
    lo, hi = pr.row_slice()
    out_buf = torch.empty((hi - lo,) + S.shape[1:], dtype=torch.float64, pin_memory=True); t.append(time.perf_counter())
    m = pr.finish(); t.append(time.perf_counter())
    idx = torch.nonzero(m[lo:hi]).squeeze(1); n = int(idx.numel()); t.append(time.perf_counter())
    mask_host = torch.empty(S.shape[0], dtype=torch.bool, pin_memory=True)
    mask_host.copy_(m, non_blocking=True)
    dev_rows = torch.index_select(pr.S[lo:hi], 0, idx)
    out_buf[:n].copy_(dev_rows, non_blocking=True)
    torch.cuda.current_stream().synchronize(); t.append(time.perf_counter())
    rows.append(np.diff(t) * 1e3)
names = ["RmsdPruner() incl. plan", "run_async enqueue", "GPU drains (upload, gather, pack, screen, verify, ladder)",
         "pinned output buffer", "finish (status readback)", "nonzero (sync)", "gather + D2H of the slice + mask"]
r = np.array(rows)[3:]
if rank == 0:
    for n_, v in zip(names, r.mean(0)):
        print(f"{n_:58s} {v:7.3f} ms")
    print(f"{'total':58s} {r.sum(1).mean():7.3f} ms   (world {world})")
# device-side pieces
pr = RmsdPruner(S, atomnos, 0.5, rank=rank, world=world, group=g)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
torch.cuda.synchronize(); dist.barrier()
ev[0].record(); pr._upload_sharded(pr._host); ev[1].record()
host = pr._host; pr._host = None
pr.pack(); ev[2].record(); pr.screen(); ev[3].record(); pr.verify(); ev[4].record(); mm = pr.eliminate(); ev[5].record()
torch.cuda.synchronize()
if rank == 0:
    print("device: upload 1/world + NVLink all-gather %.3f | pack %.3f | screen %.3f | verify %.3f | ladder %.3f ms" % tuple(
        ev[i].elapsed_time(ev[i + 1]) for i in range(5)))
ts = []
for _ in range(6):
    dist.barrier(); t0 = time.perf_counter(); prune_conformers_rmsd(S, atomnos, 0.5, group=g); ts.append((time.perf_counter() - t0) * 1e3)
if rank == 0:
    print("prune_conformers_rmsd(pinned, group) calls:", " ".join(f"{x:.2f}" for x in ts))
dist.barrier()
dist.destroy_process_group()
