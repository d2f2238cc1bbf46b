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

__all__ = ["gen_ensemble", "gen_poses", "materialise_poses", "mask_digest", "CONFIGS"]


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
