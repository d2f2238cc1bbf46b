#!/usr/bin/env python
"""Parity + timing of the BASELINE.json configs that are not bench.py's headline line
(configs[0] C1, configs[3] C4 rot_corr, configs[4] C5 end-to-end) on ONE GPU.
Prints one JSON object per config (also written to gpurun_out/<tag>/configs.jsonl)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from oracle import oracle_c, oracle_np  # noqa: E402  (checker / CPU baseline only)
import rotor_molecules as rm  # noqa: E402
from tscode_b200.numba_functions import PoseBatch  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner, prune_conformers_rmsd  # noqa: E402
from tscode_b200.synth import gen_ensemble, gen_poses, mask_digest  # noqa: E402
from tscode_b200.torsion_module import (RotCorrPruner, TorsionInfo, ladder_replay, ladder_replay_scan,  # noqa: E402
                                        prune_conformers_rmsd_rot_corr)

out_path = sys.argv[1] if len(sys.argv) > 1 else None
lines = []


def emit(d):
    print(json.dumps(d), flush=True)
    lines.append(d)


def cuda_ms(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), r


# ---- C1: prune_conformers_rmsd 1000 x 40 ---------------------------------------------------------
S = gen_ensemble(0, 1000, 40, 100); atomnos = np.full(40, 6)
prune_conformers_rmsd(S, atomnos, 0.5)
t0 = time.perf_counter(); out, mask = prune_conformers_rmsd(S, atomnos, 0.5); t_api = time.perf_counter() - t0
pr = RmsdPruner(S, atomnos, 0.5)
t_dev, _ = cuda_ms(pr.run)
t0 = time.perf_counter(); m_ref, ne, _ = oracle_c.prune_heavy(S, 0.5); t_cpu = time.perf_counter() - t0
emit({"config": "C1 prune_conformers_rmsd 1000x40 thr 0.5", "survivors": int(mask.sum()), "digest": mask_digest(mask),
      "matches_reference_digest": mask_digest(mask) == "94e7a6f4441ed28b", "api_ms": t_api * 1e3, "device_ms": t_dev,
      "cpu_port_lazy_ms": t_cpu * 1e3, "cpu_pairs_evaluated": int(ne), "cpu_threads": oracle_c.num_threads(),
      "reference_numba_ms_survey": 590})

# ---- C4: rot_corr -------------------------------------------------------------------------------------
fx = json.load(open(os.path.join(ROOT, "tests/golden/rotcorr.json")))["fixtures"]["tritbu63_s7"]
g = np.load(os.path.join(ROOT, "tests/golden/rotcorr_tritbu63_s7.npz"))
info = TorsionInfo([tuple(t) for t in fx["torsions"]], [tuple(a) for a in fx["angles"]], g["rot_masks"].astype(bool),
                   g["node_lists"].astype(bool))
S300, atomnos63 = rm.ensemble_tritbu63(fx["seed"], fx["N"])
prune_conformers_rmsd_rot_corr(S300, atomnos63, None, fx["thr"], torsion_info=info)
t0 = time.perf_counter(); out, mask = prune_conformers_rmsd_rot_corr(S300, atomnos63, None, fx["thr"], torsion_info=info)
t300 = time.perf_counter() - t0
emit({"config": "C4a rot_corr 300 x 63 atoms, 6 rotors (fixture, reference ran it in %.1f s)" % fx["wall_s"],
      "survivors": int(mask.sum()), "matches_reference": bool(np.array_equal(mask, g["mask"])),
      "returned_structures_maxdev": float(np.abs(out - g["out"]).max()), "api_ms": t300 * 1e3})
S750, _ = rm.ensemble_tritbu63(11, 750)
t0 = time.perf_counter(); out, mask = prune_conformers_rmsd_rot_corr(S750, atomnos63, None, 0.25, torsion_info=info)
t750 = time.perf_counter() - t0
emit({"config": "C4b rot_corr 750 x 63 (the reference's own size limit, torsion_module.py:1056)",
      "survivors": int(mask.sum()), "api_ms": t750 * 1e3, "pairs": 750 * 749 // 2,
      "reference_per_pair_ms_survey": 0.57})
N4 = int(os.environ.get("C4_N", "20000"))
S20, _ = rm.ensemble_tritbu63(13, N4)
Sc = S20 - S20.mean(axis=1, keepdims=True)
pr4 = RotCorrPruner(Sc, atomnos63, info, 0.25, want_codes=False)
pr4.scan(); torch.cuda.synchronize()
t0 = time.perf_counter(); first_hit, lookup = pr4.scan(); t_scan = time.perf_counter() - t0
t0 = time.perf_counter(); mask, state = ladder_replay_scan(first_hit, N4, lookup); t_lad = time.perf_counter() - t0
t0 = time.perf_counter(); out4, mask4 = prune_conformers_rmsd_rot_corr(S20, atomnos63, None, 0.25, torsion_info=info,
                                                                      max_structures=None); t_api = time.perf_counter() - t0
rng = np.random.default_rng(0)
ii = rng.integers(0, N4 - 1, 300)
bad = 0
for a in ii:
    b = int(min(first_hit[a], N4 - 1))
    for bb in {b, min(a + 1, N4 - 1)}:
        if a < bb:
            r, _, _ = oracle_np.rotationally_corrected_rmsd(Sc[a], Sc[bb], atomnos63 != 1, info.torsions, info.angles,
                                                            info.rot_masks, [np.flatnonzero(n) for n in info.node_masks])
            bad += int((r < 0.25) != (bb == first_hit[a])) if bb <= first_hit[a] else 0
npairs = N4 * (N4 - 1) // 2
emit({"config": f"C4c rot_corr {N4} x 63 atoms, 6 rotors, guard lifted (beyond the reference's 750 limit), forward scan",
      "pairs_total": npairs, "pairs_evaluated": int(pr4.pairs_evaluated), "scan_ms_incl_code_d2h": t_scan * 1e3,
      "host_ladder_ms": t_lad * 1e3, "api_total_ms": t_api * 1e3, "survivors": int(mask.sum()),
      "api_mask_equal": bool(np.array_equal(mask, mask4)),
      "near_threshold_pairs": int(pr4.near.item()), "sampled_pairs_vs_numpy_oracle_mismatches": bad,
      "local_rmsd_evals_per_pair": int(sum(len(a) for a in info.angles)) + 1})
if N4 <= 6000:      # all-pairs cross-check (O(N^2) whatever the data)
    pr5 = RotCorrPruner(Sc, atomnos63, info, 0.25, want_codes=True)
    t_pairs, _ = cuda_ms(pr5.similarity, reps=2)
    m5, _ = ladder_replay(pr5.similar_matrix(), N4, pr5.best_angles())
    emit({"config": f"C4c' all-pairs evaluation of the same {N4} structures", "gpu_all_pairs_ms": t_pairs,
          "pairs_per_s": npairs / (t_pairs * 1e-3), "mask_equal_to_scan": bool(np.array_equal(m5, mask))})
    del pr5
del pr4
torch.cuda.empty_cache()

# ---- C5: 1M trimolecular poses -> clash -> gather -> prune (one GPU) ----------------------------------------
P5 = int(os.environ.get("C5_P", "1000000"))
frags, conf, R, t = gen_poses(2, P5, (50, 50, 50))
Rp, tp, cp = torch.from_numpy(R).pin_memory(), torch.from_numpy(t).pin_memory(), torch.from_numpy(conf.astype(np.int32)).pin_memory()
atom150 = np.full(150, 6)


def c5():
    pb = PoseBatch(frags, cp, Rp, tp)
    v = pb.clash(1.5, 0)
    keep = v.nonzero().squeeze(1)
    poses = pb.gather(keep)
    pr = RmsdPruner(poses, atom150, 0.5)
    m = pr.run()
    return v, keep, poses, m


c5(); torch.cuda.synchronize()
t0 = time.perf_counter(); v, keep, poses, m = c5(); torch.cuda.synchronize(); t_c5 = time.perf_counter() - t0
pb = PoseBatch(frags, cp, Rp, tp)
t_clash, _ = cuda_ms(lambda: pb.clash(1.5, 0))
sel = np.arange(0, P5, max(P5 // 5000, 1))
ref = oracle_c.embed_clash_batch(frags, conf[sel], R[sel], t[sel], 1.5, 0)
poses_h = poses.cpu().numpy()
t0 = time.perf_counter(); mref, ne, _ = oracle_c.prune_heavy(poses_h, 0.5); t_cpu_prune = time.perf_counter() - t0
emit({"config": f"C5 end-to-end {P5} trimolecular poses (3 x 50 atoms): transform + clash -> gather -> RMSD prune, 1 GPU",
      "poses": P5, "clash_pass": int(v.sum().item()), "clash_kernel_ms": t_clash,
      "clash_poses_per_s": P5 / (t_clash * 1e-3), "clash_sample_matches_oracle": bool(np.array_equal(v.cpu().numpy()[sel], ref)),
      "prune_survivors": int(m.sum().item()), "prune_mask_matches_oracle": bool(np.array_equal(m.cpu().numpy(), mref)),
      "end_to_end_ms_incl_h2d": t_c5 * 1e3, "h2d_bytes": int(R.nbytes + t.nbytes + conf.size * 4),
      "cpu_port_prune_ms": t_cpu_prune * 1e3})

if out_path:
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        for d in lines:
            f.write(json.dumps(d) + "\n")
