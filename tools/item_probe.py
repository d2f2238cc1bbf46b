#!/usr/bin/env python
"""Probe: screen time on C3 as a function of the work-item length (j tiles per item) and configuration —
separates the per-item cost (pipeline drain + panel rows -> TMEM) from the per-tile cost."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200 import _host  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

S = gen_ensemble(3, 50000, 80, 5000)
cfgs = [int(c) for c in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["-2", "-8"])]
chunks = [int(c) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["64", "128", "256", "512"])]
for cfg in cfgs:
    pr = RmsdPruner(S, np.full(80, 6), 0.5, variant="f16", grid_ctas=cfg)
    pr.pack()
    for chunk in chunks:
        items = np.ascontiguousarray(_host.build_tf32_items(pr.N, pr.row_blocks_np, chunk=chunk))
        pr.items, pr.n_items = torch.from_numpy(items).to(pr.device), int(items.shape[0])
        for _ in range(2):
            pr.screen()
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            e0.record(); pr.screen(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        pr.verify()
        mask = pr.eliminate().cpu().numpy()
        print(f"cfg {cfg} chunk {chunk}: items {pr.n_items} screen {min(ts):.3f} ms (median {sorted(ts)[len(ts)//2]:.3f}) "
              f"digest {mask_digest(mask)} {pr.stats_dict()}", flush=True)
