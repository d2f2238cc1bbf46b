#!/usr/bin/env python
"""C2 (100k two-fragment poses) / C5 (1M three-fragment poses) clash screen, timed with CUDA events (profiling aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.numba_functions import PoseBatch
from tscode_b200.synth import gen_poses, mask_digest
for name, seed, P, na in (("C2", 0, 100_000, (50, 50)), ("C5", 2, 1_000_000, (50, 50, 50))):
    frags, conf, R, t = gen_poses(seed, P, na)
    pb = PoseBatch(frags, conf, R, t)
    for _ in range(3):
        v = pb.clash(1.5, 0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); v = pb.clash(1.5, 0); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name}: {min(ts):.4f} ms  {P / min(ts) / 1e3:.1f} M poses/s  pass {int(v.sum())} digest {mask_digest(v.cpu().numpy())}", flush=True)
