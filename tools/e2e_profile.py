#!/usr/bin/env python
"""Where does the end-to-end time of the drop-in prune_conformers_rmsd go?  (host-side breakdown)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.rmsd_pruning import RmsdPruner, prune_conformers_rmsd
from tscode_b200.synth import gen_ensemble
S = gen_ensemble(3, 50000, 80, 5000); atomnos = np.full(80, 6)
def T(label, fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"{label:55s} {min(ts)*1e3:8.2f} ms", flush=True); return r
T("prune_conformers_rmsd(numpy pageable) total", lambda: prune_conformers_rmsd(S, atomnos, 0.5))
pin = torch.empty(S.shape, dtype=torch.float64).pin_memory(); pin.copy_(torch.from_numpy(S)); Sp = pin.numpy()
T("prune_conformers_rmsd(numpy pinned) total", lambda: prune_conformers_rmsd(Sp, atomnos, 0.5))
T("H2D pageable 96 MB", lambda: torch.from_numpy(S).to("cuda"))
T("H2D pinned 96 MB", lambda: pin.to("cuda"))
Sd = pin.to("cuda")
pr = T("RmsdPruner.__init__ (device input)", lambda: RmsdPruner(Sd, atomnos, 0.5))
mask_d = T("pr.run()", pr.run)
mask = T("mask D2H", lambda: mask_d.cpu().numpy())
T("structures[mask] (numpy bool index)", lambda: S[mask])
idx = np.flatnonzero(mask)
T("np.take(S, idx, axis=0)", lambda: np.take(S, idx, axis=0))
out = np.empty((idx.size,) + S.shape[1:])
T("np.take(..., out=prealloc)", lambda: np.take(S, idx, axis=0, out=out))
T("np.compress", lambda: np.compress(mask, S, axis=0))
T("torch index_select host (threads=%d)" % torch.get_num_threads(), lambda: torch.from_numpy(S).index_select(0, torch.from_numpy(idx)).numpy())
T("GPU gather + D2H to pageable", lambda: Sd[mask_d].cpu().numpy())
pout = torch.empty((idx.size,) + S.shape[1:], dtype=torch.float64).pin_memory()
T("GPU gather + D2H to pinned", lambda: pout.copy_(Sd[mask_d]))
T("np.empty + fill (alloc 94MB)", lambda: np.empty((idx.size,) + S.shape[1:]).fill(0))
print("cpu count", os.cpu_count())
