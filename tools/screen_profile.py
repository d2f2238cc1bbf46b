#!/usr/bin/env python
"""One ensemble, a few launches of the default screen (for ncu): python tools/screen_profile.py N M mode(-1 = automatic) [scale]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble  # noqa: E402

N, M, pace = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
scale = np.array([float(x) for x in sys.argv[4].split(",")]) if len(sys.argv) > 4 else 3.0
variant = sys.argv[5] if len(sys.argv) > 5 else "screen"
S = gen_ensemble(3, N, M, N // 10, scale=scale)
pr = RmsdPruner(S, np.full(M, 6), 0.5, variant=variant, screen_mode=None if pace < 0 else pace)
pr.pack()
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pr.screen(); e1.record(); torch.cuda.synchronize()
    print("screen ms", e0.elapsed_time(e1), flush=True)
print(pr.cand_list[0, 0].item())
