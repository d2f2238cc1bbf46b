#!/usr/bin/env python
"""Probe: screen time on C3 (or N given) as a function of how the (panel, j tile) pairs are cut into work items and
dealt to the CTAs — separates the per-item cost, the tail imbalance and the effect of the order."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200 import _host  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

cfgs = [int(c) for c in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["-10"])]
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["chunk128", "linear", "linear128", "even", "evensnake"]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
S = gen_ensemble(3, N, 80, N // 10)


def build(mode, pr):
    rb = pr.row_blocks_np
    if mode.startswith("chunk"):
        return _host.build_tf32_items(pr.N, rb, chunk=int(mode[5:]))
    if mode == "linear":
        return _host.build_tf32_items_balanced(pr.N, rb, 148)
    if mode.startswith("linear"):
        return _host.build_tf32_items_balanced(pr.N, rb, 148, max_item=int(mode[6:]))
    if mode == "even":
        return _host.build_tf32_items_even(pr.N, rb, 148, snake=False)
    if mode == "evensnake":
        return _host.build_tf32_items_even(pr.N, rb, 148, snake=True)
    raise ValueError(mode)


for cfg in cfgs:
    pr = RmsdPruner(S, np.full(80, 6), 0.5, variant="f16", grid_ctas=cfg)
    pr.pack()
    for mode in modes:
        items = np.ascontiguousarray(build(mode, pr))
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
        print(f"cfg {cfg} {mode}: items {pr.n_items} screen {min(ts):.3f} ms (median {sorted(ts)[len(ts)//2]:.3f}) "
              f"digest {mask_digest(mask)} {pr.stats_dict()}", flush=True)
