#!/usr/bin/env python
"""Probe: screen time on an ANISOTROPIC ensemble (elongated molecule: base coordinates scaled (6, 2, 1)) next to
the isotropic C3-like one.  Samuelson's bound sqrt(3) ||S||_F only excludes pairs of near-isotropic covariances;
for elongated molecules the second-stage sign test decides nearly every pair."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cfgs = [int(c) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
M = int(sys.argv[3]) if len(sys.argv) > 3 else 80
kinds = sys.argv[4].split(",") if len(sys.argv) > 4 else ["isotropic", "elongated", "planar"]
for name, scale in (("isotropic", 3.0), ("elongated", np.array([6.0, 2.0, 1.0])), ("planar", np.array([4.0, 4.0, 0.5]))):
    if name not in kinds:
        continue
    S = gen_ensemble(3, N, M, N // 10, scale=scale)
    ref = None
    for variant, cfg in [("f16", c) for c in cfgs] + [("dmma", 0)]:
        pr = RmsdPruner(S, np.full(M, 6), 0.5, variant=variant, grid_ctas=cfg)
        pr.pack()
        pr.screen(); torch.cuda.synchronize()
        ts = []
        for _ in range(2):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(); pr.screen(); e1.record(); pr.verify(); e2.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        mask = pr.eliminate().cpu().numpy()
        d = mask_digest(mask)
        ref = ref or d
        print(f"{name} N={N} M={M} {variant} cfg {cfg}: screen {min(ts):.3f} ms verify {e1.elapsed_time(e2):.3f} ms "
              f"digest {d} same={d == ref} {pr.stats_dict()}", flush=True)
