#!/usr/bin/env python
"""Freeze the LIVE reference's results for BASELINE configs[4] (end-to-end trimolecular cyclical embed) — build
container only.  TEST INFRASTRUCTURE.     python oracle/gen_golden_c5.py [--groups 4630] [--small 60]

The pose space comes from tscode_b200.synth.gen_cyclical_groups (three fragments on a triangle, 6^3 angle
combinations per group; 4 630 groups = 1 000 080 poses).  What runs here is the body of the reference's generator loop
(embeds.py:657-718) and the pruning that follows it (embedder.py:1363), with the reference's OWN functions, unmodified:

    A = align_vec_pair([end - start, directions[i]], [pivot, mol_direction])            algebra.py:258   (:682)
    S = rot_mat_from_pointer(A @ axis, angle)                                            algebra.py:325   (:696)
    mol.rotation = S @ A;  mol.position = c - S @ c + (mean(vec_pair) - A @ meanpoint)   (:703-707)
    pose = get_embed(mols, conf_ids)                                                     embeds.py:961    (:713)
    if compenetration_check(pose, ids=ids, thresh=1.5):                                  numba_functions.py:59  (:714)
        if not _rmsd_similarity(pose, angular_poses, rmsd_thr=1): keep                   rmsd_pruning.py:208    (:715)
    structures, mask = prune_conformers_rmsd(np.array(poses), atomnos, rmsd_thr=0.5)     rmsd_pruning.py:164

(align_vec_pair is evaluated once per (group, molecule) instead of once per pose: it is a pure function of
arguments that do not depend on the angles.)  Stored: counts and digests of the clash verdicts, of the kept-after-
de-duplication flags (both over ALL poses, generation order) and of the final prune mask (over the kept poses); the
small case also stores the bit strings.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from tscode_b200.synth import gen_cyclical_groups, mask_digest  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def run(seed, n_groups, store_bits):
    from tscode.algebra import align_vec_pair, rot_mat_from_pointer
    from tscode.embeds import get_embed
    from tscode.numba_functions import compenetration_check
    from tscode.rmsd_pruning import _rmsd_similarity, prune_conformers_rmsd
    d = gen_cyclical_groups(seed, n_groups)
    frags, ang = d["frags"], d["systematic_angles"]
    F, C = len(frags), ang.shape[0]
    ids = np.array([f.shape[1] for f in frags])

    class Mol:
        pass
    mols = [Mol() for _ in range(F)]
    for k, m in enumerate(mols):
        m.atomcoords = frags[k]
    P = n_groups * C
    verdict = np.zeros(P, np.uint8)
    kept = np.zeros(P, np.uint8)
    poses = []
    t0 = time.perf_counter()
    for g in range(n_groups):
        conf_ids = d["group_conf"][g]
        A = [align_vec_pair(d["ref2"][g, i], d["tgt2"][g, i]) for i in range(F)]
        axis = [A[i] @ d["axis_src"][g, i] for i in range(F)]
        cor = [A[i] @ d["atomic_pivot_mean"][g, i] for i in range(F)]
        pos = [d["vec_mean"][g, i] - A[i] @ d["pivot_mean"][g, i] for i in range(F)]
        angular_poses = []
        for c in range(C):
            for i in range(F):
                step = rot_mat_from_pointer(axis[i], float(ang[c, i]))
                mols[i].rotation = step @ A[i]
                mols[i].position = cor[i] - step @ cor[i] + pos[i]
            pose = get_embed(mols, conf_ids)
            p = g * C + c
            if compenetration_check(pose, ids=ids, thresh=1.5):
                verdict[p] = 1
                if not _rmsd_similarity(pose, angular_poses, rmsd_thr=1):
                    kept[p] = 1
                    poses.append(pose)
                    angular_poses.append(pose)
        if g % 500 == 0:
            print(f"  group {g}/{n_groups}  {time.perf_counter() - t0:.0f} s  kept {len(poses)}", flush=True)
    t_gen = time.perf_counter() - t0
    S = np.array(poses)
    atomnos = np.full(int(ids.sum()), 6)
    t0 = time.perf_counter()
    out, mask = prune_conformers_rmsd(S, atomnos, rmsd_thr=0.5)
    t_prune = time.perf_counter() - t0
    row = dict(seed=seed, n_groups=n_groups, poses=P, clash_pass=int(verdict.sum()), clash_digest=mask_digest(verdict),
               kept=int(kept.sum()), kept_digest=mask_digest(kept), survivors=int(mask.sum()), prune_digest=mask_digest(mask),
               wall_s_generate_clash_dedup=round(t_gen, 1), wall_s_prune=round(t_prune, 1))
    if store_bits:
        row.update(verdict_hex=np.packbits(verdict).tobytes().hex(), kept_hex=np.packbits(kept).tobytes().hex(),
                   mask_hex=np.packbits(mask.astype(np.uint8)).tobytes().hex())
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=4630)
    ap.add_argument("--small", type=int, default=60)
    args = ap.parse_args()
    ref_harness.install(full=True)
    import numba
    path = os.path.join(GOLD, "embed_pipeline.json")
    res = json.load(open(path)) if os.path.exists(path) else {"rows": {}}
    res["meta"] = {"numba": numba.__version__, "numpy": np.__version__, "threads": numba.get_num_threads()}
    for name, n, bits in (("small", args.small, True), ("c5", args.groups, False)):
        if n <= 0:
            continue
        res["rows"][name] = run(5, n, bits)
        print(name, {k: v for k, v in res["rows"][name].items() if not k.endswith("_hex")}, flush=True)
        json.dump(res, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
