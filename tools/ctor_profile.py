#!/usr/bin/env python
"""cProfile of the host side of one end-to-end prune (constructor + enqueue), C3, pinned input.
python tools/ctor_profile.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.rmsd_pruning import RmsdPruner, prune_conformers_rmsd  # noqa: E402
from tscode_b200.synth import gen_ensemble  # noqa: E402

S0 = gen_ensemble(3, 50000, 80, 5000)
atomnos = np.full(80, 6)
pin = torch.empty(S0.shape, dtype=torch.float64).pin_memory()
pin.copy_(torch.from_numpy(S0))
S = pin.numpy()
for _ in range(3):
    prune_conformers_rmsd(S, atomnos, 0.5)
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); pr = RmsdPruner(S, atomnos, 0.5); ts.append(time.perf_counter() - t0)
print("constructor ms:", [round(x * 1e3, 3) for x in ts])
pr_ = cProfile.Profile()
for _ in range(20):
    torch.cuda.synchronize()
    pr_.enable(); p = RmsdPruner(S, atomnos, 0.5); pr_.disable()
    p.run_async(); p.finish()
pstats.Stats(pr_).sort_stats("tottime").print_stats(22)
pr2 = cProfile.Profile()
for _ in range(20):
    p = RmsdPruner(S, atomnos, 0.5)
    torch.cuda.synchronize()
    pr2.enable(); p.run_async(); pr2.disable()
    p.finish()
pstats.Stats(pr2).sort_stats("tottime").print_stats(14)
