#!/usr/bin/env python
"""Timeline of the numpy drop-in prune_conformers_rmsd on C3 (host stamps; the same steps as rmsd_pruning.py)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.rmsd_pruning import RmsdPruner, prune_conformers_rmsd
from tscode_b200.synth import gen_ensemble
S0 = gen_ensemble(3, 50000, 80, 5000); atomnos = np.full(80, 6)
pin = torch.empty(S0.shape, dtype=torch.float64).pin_memory(); pin.copy_(torch.from_numpy(S0)); S = pin.numpy()
for _ in range(3):
    prune_conformers_rmsd(S, atomnos, 0.5)
rows = []
for rep in range(6):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    pr = RmsdPruner(S, atomnos, 0.5); t.append(time.perf_counter())
    pr.run_async(); t.append(time.perf_counter())
    out_buf = torch.empty(S.shape, dtype=torch.float64); t.append(time.perf_counter())
    out_buf.zero_(); t.append(time.perf_counter())
    m = pr.finish(); t.append(time.perf_counter())
    mask = m.cpu().numpy().astype(np.bool_); t.append(time.perf_counter())
    idx = torch.from_numpy(np.flatnonzero(mask)); out = out_buf[:idx.numel()]
    torch.index_select(torch.from_numpy(S), 0, idx, out=out); t.append(time.perf_counter())
    rows.append(np.diff(t) * 1e3)
names = ["RmsdPruner()", "run_async enqueue", "torch.empty out", "out.zero_ (first touch)", "finish (GPU wait + readback)", "mask D2H", "index_select"]
r = np.array(rows)[2:]
for n, v in zip(names, r.mean(0)):
    print(f"{n:32s} {v:7.3f} ms")
print(f"{'total':32s} {r.sum(1).mean():7.3f} ms")
ts = []
for _ in range(5):
    t0 = time.perf_counter(); prune_conformers_rmsd(S, atomnos, 0.5); ts.append((time.perf_counter() - t0) * 1e3)
print("prune_conformers_rmsd(pinned) calls:", " ".join(f"{x:.2f}" for x in ts))
# GPU-only time of the pipelined path
pr = RmsdPruner(S, atomnos, 0.5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); pr.run_async(); e1.record(); torch.cuda.synchronize()
print("GPU time upload+pack+screen+verify+ladder (events on the main stream): %.3f ms" % e0.elapsed_time(e1))
r2, m2 = prune_conformers_rmsd(S0, atomnos, 0.5)
print("result equals structures[mask]:", bool(np.array_equal(r2, S0[m2])), r2.shape)
print("threads", torch.get_num_threads(), "cpus", os.cpu_count())
