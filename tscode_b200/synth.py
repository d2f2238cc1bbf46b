"""Synthetic workloads for the conformer-ensemble hot path.

These are the generators SURVEY.md Appendix A.1 / A.3 define (they are *ours*, not the
reference's: the reference ships no benchmark inputs).  bench.py, the tests, the golden
generator (oracle/gen_golden.py) and __graft_entry__.smoke() all draw inputs from here so
that every number in the repo is quoted on the same data.

Everything is numpy Generator(PCG64) based, so a (seed, shape) tuple pins the bytes.
"""
from __future__ import annotations

import hashlib

import numpy as np

__all__ = ["gen_ensemble", "gen_poses", "gen_pose_groups", "gen_cyclical_groups", "materialise_poses", "mask_digest", "CONFIGS"]


def gen_ensemble(seed, N, M, n_clusters, sigma_cluster=1.0, sigma_noise=0.05, scale=3.0):
    """Clustered 'conformer ensemble': a Gaussian blob molecule of M heavy atoms,
    n_clusters cluster centres (base + N(0, sigma_cluster)), members = centre +
    N(0, sigma_noise), randomly permuted.  Returns (N, M, 3) float64, C-contiguous."""
    rng = np.random.default_rng(seed)
    base = rng.normal(size=(M, 3)) * scale
    base -= base.mean(axis=0)
    centers = base[None] + rng.normal(size=(n_clusters, M, 3)) * sigma_cluster
    labels = rng.integers(0, n_clusters, size=N)
    S = centers[labels] + rng.normal(size=(N, M, 3)) * sigma_noise
    return np.ascontiguousarray(S[rng.permutation(N)])


def gen_poses(seed, P, n_atoms=(50, 50), n_conf=4, blob=2.0, dmin=3.0, dmax=9.0):
    """Fragment library + per-pose rigid transforms.

    Returns (frags, conf, R, t): frags[k] is (n_conf, n_atoms[k], 3); conf is (P, F) int64;
    R is (P, F, 3, 3); t is (P, F, 3).  Fragment 0 keeps R = I, t = 0 like the reference's
    first molecule (hypermolecule_class.py:171-172)."""
    rng = np.random.default_rng(seed)
    frags = [rng.normal(size=(n_conf, n, 3)) * blob for n in n_atoms]
    F = len(n_atoms)
    conf = rng.integers(0, n_conf, size=(P, F))
    q = rng.normal(size=(P, F, 4))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                  2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
                 axis=-1).reshape(P, F, 3, 3)
    d = rng.normal(size=(P, F, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    t = d * rng.uniform(dmin, dmax, size=(P, F, 1))
    R[:, 0] = np.eye(3)
    t[:, 0] = 0.0
    return frags, conf, np.ascontiguousarray(R), np.ascontiguousarray(t)


def gen_pose_groups(seed, n_groups, steps=12, n_atoms=(30, 30), n_conf=3, blob=1.6, dist=(4.0, 7.0)):
    """Poses as a cyclical embed generates them (embeds.py:657-718): per group (one combination of conformers,
    pairing and orientation) the second fragment is stepped through `steps` angles about an axis through its own
    reactive centre while the first stays put, so neighbouring angles of a group can be similar.
    Returns (frags, conf (P, 2), R (P, 2, 3, 3), t (P, 2, 3), group_id (P,)) with P = n_groups * steps."""
    rng = np.random.default_rng(seed)
    frags = [rng.normal(size=(n_conf, n, 3)) * blob for n in n_atoms]
    # make the stepped fragment elongated along its own z axis: rotations about an axis near z move it little
    frags[1][..., :2] *= 0.35
    P = n_groups * steps
    conf = np.repeat(rng.integers(0, n_conf, size=(n_groups, 2)), steps, axis=0)
    gid = np.repeat(np.arange(n_groups), steps)
    R = np.zeros((P, 2, 3, 3)); t = np.zeros((P, 2, 3))
    R[:, 0] = np.eye(3)
    for g in range(n_groups):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        x, y, z, w = q
        R0 = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        pos = d * rng.uniform(*dist)
        tilt = rng.normal(size=3) * 0.15
        axis = R0 @ (np.array([0.0, 0.0, 1.0]) + tilt); axis /= np.linalg.norm(axis)
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        for k in range(steps):
            a = 2 * np.pi * k / steps
            Rs = np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)
            R[g * steps + k, 1] = Rs @ R0
            t[g * steps + k, 1] = pos
    return frags, conf, np.ascontiguousarray(R), np.ascontiguousarray(t), gid


def gen_cyclical_groups(seed, n_groups, steps=6, n_atoms=(50, 50, 50), n_conf=3, side=(3.0, 4.2), dup_every=3):
    """A trimolecular cyclical embed as the generator loop of embeds.py:657-718 sees it (BASELINE configs[4]): three
    fragments on the sides of a triangle, every fragment stepped through `steps` angles about the line through its
    two reactive atoms; group = one combination of conformers and triangle, poses per group = steps^3 in the order of
    embedder.systematic_angles (cartesian product, last molecule fastest).

    Returns a dict with the per-group descriptors tscode_b200.embeds.cyclical_embed_poses takes — ref2 (G, 3, 2, 3) =
    [end - start, directions[i]], tgt2 = [pivots[i].pivot, mol_direction], axis_src = reactive_coords[0] -
    reactive_coords[1], atomic_pivot_mean, vec_mean = mean(vec_pair), pivot_mean = pivots[i].meanpoint, group_conf —
    plus frags [(n_conf, n_k, 3)] and systematic_angles (steps^3, 3).

    The numbers are chosen so that every stage of the pipeline has work to do: fragments are elongated away from
    their reactive atoms, so stepping them towards the inside of the triangle makes them clash; conformer 0 of the
    last fragment is thin around its rotation axis, so its angular images are near-duplicates (the group-local
    de-duplication, rmsd_thr 1.0, removes them); and every `dup_every`-th group repeats the previous one with a
    1e-3 A jitter, so the final prune_conformers_rmsd (rmsd_thr 0.5) meets the same pose twice."""
    rng = np.random.default_rng(seed)
    F = len(n_atoms)
    frags = []
    for k, n in enumerate(n_atoms):
        X = rng.normal(size=(n_conf, n, 3)) * np.array([1.0, 1.0, 2.2]) + np.array([0.0, 0.0, 1.5])
        if k == F - 1:                                   # a rod along its own rotation axis (the line y = 0, z = -3)
            X[0, :, 0] = rng.normal(size=n) * 1.6
            X[0, :, 1:] = rng.normal(size=(n, 2)) * 0.15 + np.array([0.0, -3.0])
        X[:, 0] = np.array([-0.7, 0.0, -3.0]); X[:, 1] = np.array([0.7, 0.0, -3.0])     # the two reactive atoms
        frags.append(np.ascontiguousarray(X))
    angles = np.arange(steps) * (360.0 / steps)
    sys_angles = np.array(np.meshgrid(*([angles] * F), indexing="ij")).reshape(F, -1).T.copy()
    G = n_groups
    ref2 = np.zeros((G, F, 2, 3)); tgt2 = np.zeros((G, F, 2, 3)); axis_src = np.zeros((G, F, 3))
    apm = np.zeros((G, F, 3)); vmean = np.zeros((G, F, 3)); pmean = np.zeros((G, F, 3))
    gconf = np.zeros((G, F), dtype=np.int32)
    for g in range(G):
        if dup_every and g % dup_every == dup_every - 1 and g > 0:
            j = 1e-3
            ref2[g] = ref2[g - 1]; tgt2[g] = tgt2[g - 1]; axis_src[g] = axis_src[g - 1]; apm[g] = apm[g - 1]
            vmean[g] = vmean[g - 1] + rng.normal(size=(F, 3)) * j; pmean[g] = pmean[g - 1]; gconf[g] = gconf[g - 1]
            continue
        gconf[g] = rng.integers(0, n_conf, size=F)
        L = rng.uniform(*side)
        verts = np.array([[L / np.sqrt(3) * np.cos(a), L / np.sqrt(3) * np.sin(a), 0.0]
                          for a in np.deg2rad([90.0, 210.0, 330.0])])
        verts += rng.normal(size=(3, 3)) * 0.15
        centre = verts.mean(axis=0)
        for i in range(F):
            start, end = verts[i], verts[(i + 1) % F]
            mid = 0.5 * (start + end)
            X = frags[i][gconf[g, i]]
            r0, r1 = X[0], X[1]
            ref2[g, i, 0] = end - start
            ref2[g, i, 1] = centre - mid + rng.normal(size=3) * 0.05          # the molecule faces the centre
            tgt2[g, i, 0] = (r1 - r0) * rng.uniform(0.9, 1.1)                  # pivot
            apm[g, i] = 0.5 * (r0 + r1)
            pmean[g, i] = apm[g, i] + np.array([0.0, 0.0, -0.8])
            tgt2[g, i, 1] = pmean[g, i] - apm[g, i]                            # mol_direction
            axis_src[g, i] = r0 - r1
            vmean[g, i] = mid
    return dict(frags=frags, group_conf=gconf, ref2=ref2, tgt2=tgt2, axis_src=axis_src, atomic_pivot_mean=apm,
                vec_mean=vmean, pivot_mean=pmean, systematic_angles=sys_angles)


def materialise_poses(frags, conf, R, t, sel=None):
    """Host-side (numpy) materialisation of poses: what embeds.get_embed does per pose,
    vectorised.  Only for building inputs of the 'already materialised' clash entry point
    and for tests; the product path does this on the GPU."""
    P, F = conf.shape
    idx = np.arange(P) if sel is None else np.asarray(sel)
    parts = []
    for k in range(F):
        X = frags[k][conf[idx, k]]                      # (p, n_k, 3)
        parts.append(np.einsum("pij,pnj->pni", R[idx, k], X) + t[idx, k][:, None, :])
    return np.ascontiguousarray(np.concatenate(parts, axis=1))


def mask_digest(mask) -> str:
    """First 16 hex chars of sha256 over the bit-packed mask."""
    return hashlib.sha256(np.packbits(np.asarray(mask).astype(np.uint8)).tobytes()).hexdigest()[:16]


# BASELINE.json configs, as generator arguments.
CONFIGS = {
    "C1": dict(kind="prune", seed=0, N=1000, M=40, n_clusters=100, thr=0.5),
    "C2": dict(kind="clash", seed=0, P=100_000, n_atoms=(50, 50), thresh=1.5, max_clashes=0),
    "C3": dict(kind="prune", seed=3, N=50_000, M=80, n_clusters=5000, thr=0.5),
    "C5_clash": dict(kind="clash", seed=2, P=1_000_000, n_atoms=(50, 50, 50), thresh=1.5, max_clashes=0),
}
