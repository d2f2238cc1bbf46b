#!/usr/bin/env python
"""Probe of the default screen (rmsd_screen.cu; run under `timeout`): on small ensembles of several shapes the
candidate bits must be a superset of the oracle's similar pairs and the final bits / mask must match; then timing on
BASELINE configs[2] and on its elongated / planar variants, next to the FP64 tensor-core variant ("dmma")."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_c  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["screen"]
big = "--no-big" not in sys.argv
def _arg(name, default):
    return [int(x) if x != "auto" else None for x in sys.argv[sys.argv.index(name) + 1].split(",")] if name in sys.argv else default


modes = _arg("--modes", [None])                      # screen forms to run on the isotropic ensembles (None = automatic)
modes_aniso = _arg("--modes-aniso", [None])          # ... on the anisotropic ones (mode 0 there means billions of candidates)
modes_small = _arg("--modes-small", [None, 0, 1, 2, 3])
small = "--no-small" not in sys.argv
cases = ((0, 257, 40, 20, 0.05, 0.5, 3.0), (0, 1000, 40, 100, 0.05, 0.5, 3.0), (13, 650, 80, 30, 0.2, 0.5, 3.0),
         (4, 2000, 80, 200, 0.08, 0.5, 3.0), (5, 777, 29, 60, 0.05, 0.25, 3.0), (14, 31, 1, 2, 0.05, 0.5, 3.0),
         (6, 300, 150, 30, 0.05, 0.5, 3.0), (7, 400, 192, 30, 0.05, 0.5, 3.0), (8, 500, 100, 30, 0.05, 0.5, 3.0),
         (31, 900, 40, 60, 0.05, 0.5, [6.0, 2.0, 1.0]), (32, 700, 80, 40, 0.08, 0.5, [4.0, 4.0, 0.5]),
         (33, 600, 17, 40, 0.05, 0.3, [8.0, 1.0, 1.0]), (34, 600, 30, 40, 0.05, 0.5, [5.0, 5.0, 0.02]))
bad = 0
for seed, N, M, nc, noise, thr, scale in (cases if small else ()):
    S = gen_ensemble(seed, N, M, nc, sigma_noise=noise, scale=np.array(scale) if isinstance(scale, list) else scale)
    sim = oracle_c.sim_rows(S, thr, 0, N).astype(bool)
    ref, _, _ = oracle_c.prune_heavy(S, thr)
    for vv, md in [(v, m) for v in variants for m in (modes_small if v == "screen" else [None])]:
        pr = RmsdPruner(S, np.full(M, 6), thr, variant=vv, screen_mode=md)
        pr.sim_bits.fill_(-1)
        pr.pack(); pr.screen(); torch.cuda.synchronize()
        rows, cand = pr.sim_rows_dense()
        lost = int((sim & ~cand[:N]).sum())
        ncl = int(pr.cand_list[0, 0].item())
        pr.verify(); torch.cuda.synchronize()
        rows, fin = pr.sim_rows_dense()
        mask = pr.eliminate().cpu().numpy()
        ok = lost == 0 and int((fin[:N] != sim).sum()) == 0 and bool(np.array_equal(mask, ref))
        bad += not ok
        print(f"{'ok ' if ok else 'BAD'} {vv} mode={md}->{pr.screen_mode if vv == 'screen' else '-'} N={N} M={M} scale={scale}: similar={int(sim.sum())} candidates={int(cand[:N].sum())} "
              f"list={ncl} lost={lost} final_mismatch={int((fin[:N] != sim).sum())} mask_ok={bool(np.array_equal(mask, ref))} "
              f"{pr.stats_dict()}", flush=True)
print("small cases:", "ALL OK" if bad == 0 else f"{bad} BAD", flush=True)

if big:
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for name, scale, N, M in (("isotropic", 3.0, 50000, 80), ("elongated", np.array([6.0, 2.0, 1.0]), 50000, 80),
                              ("planar", np.array([4.0, 4.0, 0.5]), 20000, 80), ("isotropic", 3.0, 20000, 40),
                              ("isotropic", 3.0, 20000, 150)):
        S = gen_ensemble(3, N, M, N // 10, scale=scale)
        ref = None
        for vv, md in [(v, m) for v in variants for m in ((modes if name == "isotropic" else modes_aniso) if v == "screen" else [None])] + [("dmma", None)]:
            if vv == "dmma" and "--no-dmma" in sys.argv:
                continue
            pr = RmsdPruner(S, np.full(M, 6), 0.5, variant=vv, screen_mode=md)
            pr.pack()
            for _ in range(2):
                pr.screen()
            torch.cuda.synchronize()
            ts = []
            for _ in range(4):
                flush.fill_(1)
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record(); pr.screen(); e1.record(); pr.verify(); e2.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            mask = pr.eliminate().cpu().numpy()
            d = mask_digest(mask)
            ref = ref or d
            print(f"{name} N={N} M={M} {vv} mode={md}->{pr.screen_mode if vv == 'screen' else '-'}: screen {min(ts):.3f} ms (mean {sum(ts) / len(ts):.3f}) verify {e1.elapsed_time(e2):.3f} ms "
                  f"digest {d} same={d == ref} {pr.stats_dict()}", flush=True)
            del pr
