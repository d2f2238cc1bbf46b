#!/usr/bin/env python
"""First-contact probe of the tcgen05 pre-screen (run under `timeout`): small cases against the
oracle (candidate bits must be a superset of the similar pairs; final mask must match), then timing."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_c  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

for seed, N, M, nc, noise, thr in ((0, 257, 40, 20, 0.05, 0.5), (0, 1000, 40, 100, 0.05, 0.5), (13, 650, 80, 30, 0.2, 0.5),
                                   (4, 2000, 80, 200, 0.08, 0.5), (5, 777, 29, 60, 0.05, 0.25), (14, 31, 1, 2, 0.05, 0.5)):
    S = gen_ensemble(seed, N, M, nc, sigma_noise=noise)
    sim = oracle_c.sim_rows(S, thr, 0, N).astype(bool)
    pr = RmsdPruner(S, np.full(M, 6), thr, variant="tf32")
    pr.sim_bits.fill_(-1)
    pr.pack(); pr.screen(); torch.cuda.synchronize()
    rows, cand = pr.sim_rows_dense()
    lost = int((sim & ~cand[:N]).sum())
    pr.verify(); torch.cuda.synchronize()
    rows, fin = pr.sim_rows_dense()
    mask = pr.eliminate().cpu().numpy()
    ref, _, _ = oracle_c.prune_heavy(S, thr)
    print(f"N={N} M={M}: similar={int(sim.sum())} tf32-candidates={int(cand[:N].sum())} lost={lost} "
          f"final_mismatch={int((fin[:N] != sim).sum())} mask_ok={bool(np.array_equal(mask, ref))} {pr.stats_dict()}", flush=True)
    assert lost == 0

S = gen_ensemble(3, 50000, 80, 5000)
for variant in ("tf32", "tf32ss", "dmma"):
    pr = RmsdPruner(S, np.full(80, 6), 0.5, variant=variant)
    pr.pack()
    for _ in range(2):
        pr.screen()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); pr.screen(); e1.record(); pr.verify(); e2.record(); torch.cuda.synchronize()
    mask = pr.eliminate().cpu().numpy()
    print(f"C3 {variant}: screen {e0.elapsed_time(e1):.2f} ms  verify {e1.elapsed_time(e2):.2f} ms  digest {mask_digest(mask)} "
          f"{pr.stats_dict()}", flush=True)
