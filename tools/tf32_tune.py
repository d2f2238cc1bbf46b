#!/usr/bin/env python
"""Time the tf32 screen under the alternative epilogue configurations (tuning aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.rmsd_pruning import RmsdPruner
from tscode_b200.synth import gen_ensemble, mask_digest
S = gen_ensemble(3, 50000, 80, 5000)
for variant, cfg, name in (("tf32", -1, "A in TMEM, 2 groups x8"), ("tf32", -2, "A in TMEM, 2 groups x4"), ("tf32", -3, "A in TMEM, 4 groups (column halves) x4"),
                           ("tf32ss", -3, "A in smem, 2 groups x4")):
    pr = RmsdPruner(S, np.full(80, 6), 0.5, variant=variant, grid_ctas=cfg)
    pr.pack()
    for _ in range(2):
        pr.screen()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pr.screen(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    pr.verify(); mask = pr.eliminate().cpu().numpy()
    print(f"cfg {name}: screen {min(ts):.3f} ms (median {sorted(ts)[2]:.3f})  digest {mask_digest(mask)} {pr.stats_dict()}", flush=True)
