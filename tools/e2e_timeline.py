#!/usr/bin/env python
"""Timeline of the numpy drop-in prune_conformers_rmsd on C3 (host stamps; the same steps as rmsd_pruning.py):
constructor (incl. the screen plan), enqueue of the pipelined upload / pack / screen / verify / ladder, output
allocation, wait for the mask, survivor gather on the GPU and D2H.  python tools/e2e_timeline.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200 import rmsd_pruning as rp  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner, prune_conformers_rmsd  # noqa: E402
from tscode_b200.synth import gen_ensemble  # noqa: E402

S0 = gen_ensemble(3, 50000, 80, 5000)
atomnos = np.full(80, 6)
pin = torch.empty(S0.shape, dtype=torch.float64).pin_memory()
pin.copy_(torch.from_numpy(S0))
S = pin.numpy()
for _ in range(3):
    prune_conformers_rmsd(S, atomnos, 0.5)
rows = []
for rep in range(8):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    pr = RmsdPruner(S, atomnos, 0.5); t.append(time.perf_counter())
    pr.run_async(); t.append(time.perf_counter())
    out_buf = torch.empty(S.shape, dtype=torch.float64, pin_memory=True); t.append(time.perf_counter())
    m = pr.finish(); t.append(time.perf_counter())
    idx = torch.nonzero(m).squeeze(1); n = int(idx.numel()); t.append(time.perf_counter())
    mask_host = torch.empty(S.shape[0], dtype=torch.bool, pin_memory=True)
    mask_host.copy_(m, non_blocking=True)
    dev_rows = torch.index_select(pr.S, 0, idx)
    out_buf[:n].copy_(dev_rows, non_blocking=True)
    t.append(time.perf_counter())
    torch.cuda.current_stream().synchronize(); t.append(time.perf_counter())
    rows.append(np.diff(t) * 1e3)
names = ["RmsdPruner() incl. plan", "run_async enqueue", "pinned output buffer", "finish (GPU wait + status readback)",
         "nonzero (sync: count)", "enqueue gather + D2H", "wait for D2H"]
r = np.array(rows)[3:]
for n_, v in zip(names, r.mean(0)):
    print(f"{n_:38s} {v:7.3f} ms")
print(f"{'total':38s} {r.sum(1).mean():7.3f} ms")
ts = []
for _ in range(6):
    t0 = time.perf_counter(); prune_conformers_rmsd(S, atomnos, 0.5); ts.append((time.perf_counter() - t0) * 1e3)
print("prune_conformers_rmsd(pinned) calls:", " ".join(f"{x:.2f}" for x in ts))
# GPU-only time of the pipelined path, and of its pieces
pr = RmsdPruner(S, atomnos, 0.5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); pr.run_async(); e1.record(); torch.cuda.synchronize()
print("GPU time upload+pack+screen+verify+ladder (events on the main stream): %.3f ms" % e0.elapsed_time(e1))
x = torch.empty(S.shape, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); e0.record(); x.copy_(pin, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("plain H2D of the ensemble: %.3f ms" % e0.elapsed_time(e1))
y = torch.empty((48867,) + S.shape[1:], dtype=torch.float64, pin_memory=True)
torch.cuda.synchronize(); e0.record(); y.copy_(x[:48867], non_blocking=True); e1.record(); torch.cuda.synchronize()
print("plain D2H of the survivors: %.3f ms" % e0.elapsed_time(e1))
t0 = time.perf_counter(); pr2 = RmsdPruner(S, atomnos, 0.5); t1 = time.perf_counter()
print("constructor again: %.3f ms (sample_undecided %.4f, mode %d)" % ((t1 - t0) * 1e3, pr2.sample_undecided, pr2.screen_mode))
r2, m2 = prune_conformers_rmsd(S0, atomnos, 0.5)
print("result equals structures[mask]:", bool(np.array_equal(r2, S0[m2])), r2.shape)
print("threads", torch.get_num_threads(), "cpus", os.cpu_count())
for share in (0.0, 0.35, 0.45, 0.55, 0.65, 0.75, 0.85):
    rp.HOST_GATHER_SHARE = share
    prune_conformers_rmsd(S, atomnos, 0.5)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); r3, m3 = prune_conformers_rmsd(S, atomnos, 0.5); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"host gather share {share:.2f}: min {min(ts):.2f} median {sorted(ts)[4]:.2f} ms  equal {bool(np.array_equal(r3, S0[m3]))}")
