#!/usr/bin/env python
"""One short program that launches every hot kernel once or twice on representative inputs, for ncu:
  1. prune step on the isotropic BASELINE-like ensemble (30 000 x 80): screen form 0, verify (list), fused ladder
  2. the same on an elongated molecule (20 000 x 80): screen form 2
  3. fused transform + clash screen on 100 000 two-fragment poses that ALL pass (no early exit: the full pair count)
     and on the BASELINE configs[1] set (87 % clash early)
  4. rot_corr forward scan, 5 000 structures x 63 atoms
python tools/profile_kernels.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rotor_molecules as rm  # noqa: E402
from tscode_b200.numba_functions import PoseBatch  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, gen_poses  # noqa: E402
from tscode_b200.torsion_module import RotCorrPruner, TorsionInfo  # noqa: E402


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), r


out = {}
for name, scale, N in (("isotropic", 3.0, 30000), ("elongated", np.array([6.0, 2.0, 1.0]), 20000)):
    S = gen_ensemble(3, N, 80, N // 10, scale=scale)
    pr = RmsdPruner(S, np.full(80, 6), 0.5)
    for rep in range(2):
        pr.pack()
        t_s, _ = timed(pr.screen)
        t_v, _ = timed(pr.verify)
        t_e, m = timed(pr.eliminate)
    out[name] = {"N": N, "screen_mode": pr.screen_mode, "screen_ms": t_s, "verify_ms": t_v, "eliminate_ms": t_e,
                 "survivors": int(m.sum().item()), **pr.stats_dict()}
    del pr
for name, kw in (("all_pass", dict(dmin=14.0, dmax=20.0)), ("configs1", dict())):
    frags, conf, R, t = gen_poses(0, 100000, (50, 50), **kw)
    pb = PoseBatch(frags, conf, R, t)
    for rep in range(2):
        t_c, v = timed(lambda: pb.clash(1.5, 0))
    out["clash_" + name] = {"poses": 100000, "kernel_ms": t_c, "passes": int(v.sum().item()),
                            "fp64_tflops_algorithmic": 21800 * 100000 / (t_c * 1e-3) / 1e12}
f = json.load(open(os.path.join(ROOT, "tests/golden/rotcorr.json")))["fixtures"]["tritbu63_s7"]
g = np.load(os.path.join(ROOT, "tests/golden/rotcorr_tritbu63_s7.npz"))
info = TorsionInfo([tuple(t) for t in f["torsions"]], [tuple(a) for a in f["angles"]], g["rot_masks"].astype(bool),
                   g["node_lists"].astype(bool))
S5, at = rm.ensemble_tritbu63(13, 5000)
prc = RotCorrPruner(S5 - S5.mean(axis=1, keepdims=True), at, info, 0.25, want_codes=False)
for rep in range(2):
    t_r, (fh, lk) = timed(prc.scan)
out["rotcorr_scan"] = {"N": 5000, "scan_ms_incl_compaction_and_d2h": t_r, "pairs_evaluated": prc.pairs_evaluated}
print(json.dumps(out, indent=1))
