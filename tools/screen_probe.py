#!/usr/bin/env python
"""First-contact probe of the tcgen05 pre-screens (run under `timeout`): small cases against the
oracle (candidate bits must be a superset of the similar pairs; final mask must match), then timing
of the epilogue configurations on C3."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_c  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

for seed, N, M, nc, noise, thr in ((0, 257, 40, 20, 0.05, 0.5), (0, 1000, 40, 100, 0.05, 0.5), (13, 650, 80, 30, 0.2, 0.5),
                                   (4, 2000, 80, 200, 0.08, 0.5), (5, 777, 29, 60, 0.05, 0.25), (14, 31, 1, 2, 0.05, 0.5),
                                   (6, 300, 150, 30, 0.05, 0.5)):
    S = gen_ensemble(seed, N, M, nc, sigma_noise=noise)
    sim = oracle_c.sim_rows(S, thr, 0, N).astype(bool)
    ref, _, _ = oracle_c.prune_heavy(S, thr)
    for vv, cfg in (("f16", 0), ("f16", -11), ("tf32", 0)):
        if vv == "tf32" and M > 120:
            continue
        pr = RmsdPruner(S, np.full(M, 6), thr, variant=vv, grid_ctas=cfg)
        pr.sim_bits.fill_(-1)
        pr.pack(); pr.screen(); torch.cuda.synchronize()
        rows, cand = pr.sim_rows_dense()
        lost = int((sim & ~cand[:N]).sum())
        pr.verify(); torch.cuda.synchronize()
        rows, fin = pr.sim_rows_dense()
        mask = pr.eliminate().cpu().numpy()
        print(f"{vv}{cfg} N={N} M={M}: similar={int(sim.sum())} candidates={int(cand[:N].sum())} lost={lost} "
              f"final_mismatch={int((fin[:N] != sim).sum())} mask_ok={bool(np.array_equal(mask, ref))} {pr.stats_dict()}", flush=True)
        assert lost == 0

S = gen_ensemble(3, 50000, 80, 5000)
for variant, cfg in (("f16", 0), ("f16", -11), ("tf32", 0)):
    pr = RmsdPruner(S, np.full(80, 6), 0.5, variant=variant, grid_ctas=cfg)
    pr.pack()
    for _ in range(2):
        pr.screen()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); pr.screen(); e1.record(); pr.verify(); e2.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    mask = pr.eliminate().cpu().numpy()
    print(f"C3 {variant} cfg {cfg}: screen {min(ts):.3f} ms  verify {e1.elapsed_time(e2):.2f} ms  digest {mask_digest(mask)} "
          f"{pr.stats_dict()}", flush=True)
