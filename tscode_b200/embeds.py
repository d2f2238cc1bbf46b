"""Batched form of the embed post-processing the reference runs pose by pose (tscode/embeds.py):

    for every generated pose:   get_embed -> compenetration_check           (embeds.py:116-118, 713-714, 841-842)
    afterwards, on the poses kept: prune_conformers_rmsd                    (embedder.py:1356-1382)

`screen_and_prune` does both for P poses given as (conformer id, R, t) per fragment, on one GPU or
sharded over the ranks of a torch.distributed group ("End-to-end C5"):

    phase 1  poses are split into contiguous ranges, one per rank; fused transform + clash screen
             (tsc_embed_clash), survivors materialised locally (tsc_embed_gather);
    exchange survivor counts and coordinates all-gathered over NCCL (variable sizes: counts first),
             verdict bytes all-gathered so that every rank holds the full clash mask;
    phase 2  row-sharded all-pairs prune of the survivors (RmsdPruner: screen / verify on the owned
             rows, ONE all-gather of the confirmed-pair lists, fused ladder on every rank).

Everything returned is identical on all ranks and identical to the single-GPU result.
"""
from __future__ import annotations

import numpy as np

from ._lib import require_cuda
from .numba_functions import PoseBatch
from .rmsd_pruning import RmsdPruner


def pose_range(P: int, rank: int, world: int):
    """Contiguous pose range of a rank (poses are independent units: no balancing needed)."""
    per = (P + world - 1) // world
    lo = min(rank * per, P)
    return lo, min(lo + per, P)


def string_embed_poses(frags, centers, vecs, angles):
    """The pose space of a string embed (embeds.py:91-114) generated ON THE DEVICE.

    frags   : [atomcoords of mol1 (n_conf1, n_1, 3), atomcoords of mol2 (n_conf2, n_2, 3)]
    centers : [ra1.center per conformer (n_conf1, n_c1, 3), ra2.center (n_conf2, n_c2, 3)]   (mol.get_r_atoms(c)[0].center)
    vecs    : [ra1.orb_vecs per conformer (n_conf1, n_c1, 3), ra2.orb_vecs (n_conf2, n_c2, 3)]
    angles  : embedder.systematic_angles (degrees)
    Returns a PoseBatch of P = n_conf1*n_conf2*n_c1*n_c2*len(angles) poses in the reference's loop order
    (conformer pairs, then centre pairs, then angles); `.clash()` screens them, `.gather(keep)` materialises
    survivors.  Only the small centre / vector tables cross PCIe."""
    torch = require_cuda()
    from ._lib import check, lib, ptr, stream_ptr
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    c1, c2 = (np.ascontiguousarray(c, dtype=np.float64) for c in centers)
    v1, v2 = (np.ascontiguousarray(v, dtype=np.float64) for v in vecs)
    n_conf1, n_c1 = c1.shape[:2]
    n_conf2, n_c2 = c2.shape[:2]
    if v1.shape != c1.shape or v2.shape != c2.shape or frags[0].shape[0] != n_conf1 or frags[1].shape[0] != n_conf2:
        raise ValueError("centers / vecs must be (n_conf, n_centres, 3) per molecule")
    ang = np.asarray(angles, dtype=np.float64).reshape(-1)
    half = ang * np.pi / 180 / 2                                    # `angle *= np.pi/180`, then angle/2 (algebra.py:337-341)
    # rot_mat_from_pointer([0, 0, 1], 180) exactly as the reference evaluates it (utils.py:203-205)
    q = np.array([0.0, 0.0, np.sin(np.pi / 2) * 1.0, np.cos(np.pi / 2)])
    q0, q1, q2, q3 = q[3], q[0], q[1], q[2]
    flip = np.array([2 * (q0 * q0 + q1 * q1) - 1, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2),
                     2 * (q1 * q2 + q0 * q3), 2 * (q0 * q0 + q2 * q2) - 1, 2 * (q2 * q3 - q0 * q1),
                     2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 2 * (q0 * q0 + q3 * q3) - 1])
    P = n_conf1 * n_conf2 * n_c1 * n_c2 * ang.size
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_c1, d_v1, d_c2, d_v2 = up(c1), up(v1), up(c2), up(v2)
    d_sin, d_cos, d_nz, d_flip = up(np.sin(half)), up(np.cos(half)), up((ang != 0).astype(np.uint8)), up(flip)
    conf = torch.empty((max(P, 1), 2), dtype=torch.int32, device=dev)
    R = torch.empty((max(P, 1), 2, 3, 3), dtype=torch.float64, device=dev)
    t = torch.empty((max(P, 1), 2, 3), dtype=torch.float64, device=dev)
    check(lib().tsc_string_embed_params(ptr(d_c1), ptr(d_v1), ptr(d_c2), ptr(d_v2), n_conf1, n_conf2, n_c1, n_c2,
                                        ptr(d_sin), ptr(d_cos), ptr(d_nz), int(ang.size), ptr(d_flip), ptr(conf), ptr(R),
                                        ptr(t), stream_ptr()), "tsc_string_embed_params")
    return PoseBatch(frags, conf[:P], R[:P], t[:P])


def cyclical_embed_poses(frags, group_conf, ref2, tgt2, axis_src, atomic_pivot_mean, vec_mean, pivot_mean,
                         systematic_angles):
    """The pose space of a cyclical embed (embeds.py:657-718) generated ON THE DEVICE.

    Per group g (one combination of conformers, pivots and polygon orientation) and molecule i the host supplies
    what the reference's loop derives before touching the angles:
      group_conf (G, F) conformer ids;  ref2 (G, F, 2, 3) = [end - start, directions[i]];
      tgt2 (G, F, 2, 3) = [pivots[i].pivot, mol_direction];  axis_src (G, F, 3) = reactive_coords[0] -
      reactive_coords[1] (two reactive atoms) or pivots[i].pivot;  atomic_pivot_mean (G, F, 3);
      vec_mean (G, F, 3) = np.mean(vec_pair, axis=0);  pivot_mean (G, F, 3) = pivots[i].meanpoint;
    systematic_angles: (C, F) degrees (embedder.systematic_angles).
    Returns (PoseBatch of G*C poses in the reference's order — groups outermost, angles innermost —, group_id (G*C,)
    device tensor for dedup_groups)."""
    torch = require_cuda()
    from ._lib import check, lib, ptr, stream_ptr
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    group_conf = np.ascontiguousarray(group_conf, dtype=np.int32)
    G, F = group_conf.shape
    ang = np.asarray(systematic_angles, dtype=np.float64).reshape(-1, F)
    C = ang.shape[0]
    table, inv = np.unique(ang, return_inverse=True)
    combos = np.ascontiguousarray(inv.reshape(C, F).astype(np.int32))
    half = table * np.pi / 180 / 2                                  # algebra.py:337-341
    up = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    d_ref, d_tgt = up(np.reshape(ref2, (G, F, 2, 3))), up(np.reshape(tgt2, (G, F, 2, 3)))
    d_axis, d_apm, d_vm, d_pm = (up(np.reshape(a, (G, F, 3))) for a in (axis_src, atomic_pivot_mean, vec_mean, pivot_mean))
    d_gconf, d_combos = up(group_conf, np.int32), up(combos, np.int32)
    d_sin, d_cos = up(np.sin(half)), up(np.cos(half))
    P = G * C
    scratch = torch.empty(max(G * F * 18, 1), dtype=torch.float64, device=dev)
    conf = torch.empty((max(P, 1), F), dtype=torch.int32, device=dev)
    R = torch.empty((max(P, 1), F, 3, 3), dtype=torch.float64, device=dev)
    t = torch.empty((max(P, 1), F, 3), dtype=torch.float64, device=dev)
    check(lib().tsc_cyclical_embed_params(ptr(d_ref), ptr(d_tgt), ptr(d_axis), ptr(d_apm), ptr(d_vm), ptr(d_pm),
                                          ptr(d_gconf), G, F, ptr(d_combos), C, ptr(d_sin), ptr(d_cos), ptr(scratch),
                                          ptr(conf), ptr(R), ptr(t), stream_ptr()), "tsc_cyclical_embed_params")
    gid = torch.arange(G, device=dev).repeat_interleave(C)
    return PoseBatch(frags, conf[:P], R[:P], t[:P]), gid


def dedup_groups(poses, group_id, passed=None, rmsd_thr=1.0):
    """Group-local de-duplication of generated poses — the `_rmsd_similarity(pose, angular_poses, rmsd_thr=1)`
    step of the cyclical embeds (embeds.py:714-718, 842-846), batched.

    poses    : (P, A, 3) float64, numpy or device tensor, in generation order
    group_id : (P,) ints, non-decreasing: poses of one (conformers, pairing, orientation) combination share an id
    passed   : (P,) bool clash verdicts (compenetration_check); None = all passed
    Returns keep (P,) bool device tensor: pose p is kept iff it passed and is not similar (all atoms, rmsd < thr and
    max deviation < 2 thr) to any pose of its group kept before it — exactly the reference's sequential logic."""
    torch = require_cuda()
    from ._lib import check, lib, ptr, stream_ptr
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    S = poses.to(dev, dtype=torch.float64).contiguous() if torch.is_tensor(poses) else \
        torch.as_tensor(np.ascontiguousarray(poses, dtype=np.float64)).to(dev)
    P, A = int(S.shape[0]), int(S.shape[1])
    gid = torch.as_tensor(np.asarray(group_id)).to(dev).to(torch.int64) if not torch.is_tensor(group_id) else group_id.to(dev).to(torch.int64)
    ok = torch.ones(P, dtype=torch.bool, device=dev) if passed is None else \
        (passed.to(dev) if torch.is_tensor(passed) else torch.as_tensor(np.asarray(passed)).to(dev)).to(torch.bool)
    keep = torch.zeros(max(P, 1), dtype=torch.uint8, device=dev)
    if P == 0:
        return keep[:0].bool()
    if P > 1 and bool((gid[1:] < gid[:-1]).any()):
        raise ValueError("group_id must be non-decreasing (poses in generation order)")
    order = ok.nonzero().squeeze(1)                                  # members, in generation order
    n = int(order.numel())
    if n == 0:
        return keep[:P].bool()
    g = gid[order]
    new = torch.ones(n, dtype=torch.bool, device=dev)
    new[1:] = g[1:] != g[:-1]
    starts = new.nonzero().squeeze(1)                                # first member of every group
    n_groups = int(starts.numel())
    g_begin = torch.cat([starts, torch.tensor([n], device=dev)]).to(torch.int32)
    gidx = torch.cumsum(new.to(torch.int64), 0) - 1                  # group index of every member
    rank = torch.arange(n, device=dev) - starts[gidx]                # earlier members of the same group
    pair_base = torch.cumsum(rank, 0) - rank
    n_pairs = int(rank.sum().item())
    sim = torch.zeros(max(n_pairs, 1), dtype=torch.uint8, device=dev)
    if n_pairs:
        m = torch.repeat_interleave(torch.arange(n, device=dev), rank)            # member of every pair ...
        r = torch.arange(n_pairs, device=dev) - pair_base[m]                       # ... and rank of its earlier partner
        pi = order[m].to(torch.int32)
        pj = order[starts[gidx[m]] + r].to(torch.int32)
        check(lib().tsc_rmsd_pairs_idx(ptr(S), ptr(pi), ptr(pj), n_pairs, A, float(rmsd_thr), ptr(sim), stream_ptr()),
              "tsc_rmsd_pairs_idx")
    order32 = order.to(torch.int32)
    check(lib().tsc_group_greedy(ptr(g_begin), n_groups, ptr(order32), ptr(pair_base), ptr(sim), ptr(keep), stream_ptr()),
          "tsc_group_greedy")
    return keep[:P].bool()


def gather_varlen(x, world: int, group=None):
    """All-gather of per-rank tensors whose first dimension differs: counts first (one host readback:
    the sizes must be known to size the receive buffer), then one padded all-gather, then the ranks'
    valid parts concatenated in rank order.  Works on any backend (NCCL on device, gloo in the tests)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return x
    cnt = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    cnts = torch.empty(world, dtype=torch.int64, device=x.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    counts = cnts.tolist()
    cmax = max(max(counts), 1)
    send = torch.zeros((cmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    send[:x.shape[0]] = x
    recv = torch.empty((world * cmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return torch.cat([recv[r * cmax:r * cmax + counts[r]] for r in range(world)])


def screen_and_prune(frags, conf, R, t, atomnos, thresh=1.5, max_clashes=0, rmsd_thr=0.5, *, rank=0, world=1,
                     group=None, variant="screen"):
    """Returns dict(verdict (P,) uint8 device tensor, keep (n_pass,) int64 global pose indices, poses
    (n_pass, A, 3) device tensor of the poses that pass the clash screen, mask (n_pass,) bool device tensor
    of the RMSD prune over them, timings of the phases in ms)."""
    torch = require_cuda()
    import torch.distributed as dist
    P = int(conf.shape[0])
    lo, hi = pose_range(P, rank, world)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    pb = PoseBatch(frags, conf[lo:hi], R[lo:hi], t[lo:hi])
    v_loc = pb.clash(thresh, max_clashes)
    keep_loc = v_loc.nonzero().squeeze(1)
    poses_loc = pb.gather(keep_loc)
    ev[1].record()
    dev = poses_loc.device
    if world > 1:
        per = (P + world - 1) // world
        vpad = torch.zeros(per, dtype=torch.uint8, device=dev)
        vpad[:hi - lo] = v_loc
        vall = torch.empty(world * per, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(vall, vpad, group=group)
        verdict = vall[:P]
        poses = gather_varlen(poses_loc, world, group)
        keep = verdict.nonzero().squeeze(1)
    else:
        verdict, keep, poses = v_loc, keep_loc, poses_loc
    ev[2].record()
    pr = RmsdPruner(poses, atomnos, rmsd_thr, variant=variant, rank=rank, world=world, group=group)
    mask = pr.run() if poses.shape[0] else torch.zeros(0, dtype=torch.bool, device=dev)
    ev[3].record()
    torch.cuda.synchronize()
    return {"verdict": verdict, "keep": keep, "poses": poses, "mask": mask, "pruner": pr,
            "ms": {"clash_gather": ev[0].elapsed_time(ev[1]), "exchange": ev[1].elapsed_time(ev[2]),
                   "prune": ev[2].elapsed_time(ev[3])}}


def group_range(G: int, rank: int, world: int):
    """Contiguous range of groups of a rank (groups are independent units for clash screen and de-duplication)."""
    per = (G + world - 1) // world
    lo = min(rank * per, G)
    return lo, min(lo + per, G)


def cyclical_embed_pipeline(desc, atomnos, thresh=1.5, max_clashes=0, dedup_thr=1.0, rmsd_thr=0.5, *, rank=0, world=1,
                            group=None):
    """The whole post-generation path of a cyclical embed (BASELINE configs[4]), batched and sharded:

        for every group, for every angle combination (embeds.py:657-709):   pose parameters on the device
            get_embed -> compenetration_check (embeds.py:713-714)            fused transform + clash screen
            _rmsd_similarity(pose, angular_poses, rmsd_thr=1) (embeds.py:715) group-local greedy de-duplication
        prune_conformers_rmsd(poses, atomnos, rmsd_thr) (embedder.py:1363)   row-sharded all-pairs prune

    desc: the per-group descriptors cyclical_embed_poses takes plus `frags` (e.g. synth.gen_cyclical_groups).
    With several ranks the GROUPS are dealt in contiguous ranges (no data-path collective until the kept poses are
    all-gathered for the prune).  Returns a dict: verdict (P,) uint8 and kept (P,) bool over all poses in generation
    order, poses (n_kept, A, 3) device tensor, mask (n_kept,) bool device tensor of the prune, ms per phase.
    Everything returned is identical on all ranks and to the single-GPU result."""
    torch = require_cuda()
    import torch.distributed as dist
    G = int(desc["group_conf"].shape[0])
    C = int(np.asarray(desc["systematic_angles"]).reshape(-1, len(desc["frags"])).shape[0])
    lo, hi = group_range(G, rank, world)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    sub = {k: desc[k][lo:hi] for k in ("group_conf", "ref2", "tgt2", "axis_src", "atomic_pivot_mean", "vec_mean", "pivot_mean")}
    if hi > lo:
        pb, gid = cyclical_embed_poses(desc["frags"], sub["group_conf"], sub["ref2"], sub["tgt2"], sub["axis_src"],
                                       sub["atomic_pivot_mean"], sub["vec_mean"], sub["pivot_mean"], desc["systematic_angles"])
        dev = pb.verdict.device
        v_loc = pb.clash(thresh, max_clashes)
        ev[1].record()
        idx = v_loc.nonzero().squeeze(1)
        poses_pass = pb.gather(idx)
        keep_pass = dedup_groups(poses_pass, gid[idx], None, dedup_thr)
        kept_loc = torch.zeros(pb.P, dtype=torch.bool, device=dev)
        kept_loc[idx[keep_pass]] = True
        poses_loc = poses_pass[keep_pass]
    else:
        dev = torch.device(f"cuda:{torch.cuda.current_device()}")
        A = int(sum(f.shape[1] for f in desc["frags"]))
        v_loc = torch.zeros(0, dtype=torch.uint8, device=dev)
        ev[1].record()
        kept_loc = torch.zeros(0, dtype=torch.bool, device=dev)
        poses_loc = torch.zeros((0, A, 3), dtype=torch.float64, device=dev)
    ev[2].record()
    if world > 1:
        per = (G + world - 1) // world * C
        pad = torch.zeros((2, per), dtype=torch.uint8, device=dev)
        pad[0, :v_loc.numel()] = v_loc
        pad[1, :kept_loc.numel()] = kept_loc.to(torch.uint8)
        allp = torch.empty((world, 2, per), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allp.view(-1), pad.view(-1), group=group)
        sizes = [(group_range(G, r, world)[1] - group_range(G, r, world)[0]) * C for r in range(world)]
        verdict = torch.cat([allp[r, 0, :sizes[r]] for r in range(world)])
        kept = torch.cat([allp[r, 1, :sizes[r]] for r in range(world)]).to(torch.bool)
        poses = gather_varlen(poses_loc, world, group)
    else:
        verdict, kept, poses = v_loc, kept_loc, poses_loc
    ev[3].record()
    if poses.shape[0]:
        pr = RmsdPruner(poses, atomnos, rmsd_thr, rank=rank, world=world, group=group)
        mask = pr.run()
    else:
        pr, mask = None, torch.zeros(0, dtype=torch.bool, device=dev)
    ev[4].record()
    torch.cuda.synchronize()
    return {"verdict": verdict, "kept": kept, "poses": poses, "mask": mask, "pruner": pr, "n_poses": G * C,
            "ms": {"params_clash": ev[0].elapsed_time(ev[1]), "gather_dedup": ev[1].elapsed_time(ev[2]),
                   "exchange": ev[2].elapsed_time(ev[3]), "prune": ev[3].elapsed_time(ev[4])}}

