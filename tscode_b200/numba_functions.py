"""B200-native drop-in for the clash screen of tscode/numba_functions.py and the pose
materialisation of tscode/embeds.py.

    compenetration_check(coords, ids=None, thresh=1.5, max_clashes=0) -> int        # numba_functions.py:59-105
    get_embed(mols, conf_ids) -> ndarray                                             # embeds.py:961-969

and the batched forms the generators' inner loops collapse into:

    compenetration_check_batch(structures, ids, thresh, max_clashes) -> uint8 (P,)   # embedder.py:1245-1248
    PoseBatch(frags, conf, R, t).clash(thresh, max_clashes) -> uint8 (P,)            # embeds.py:116-118, 713-714
    PoseBatch.gather(keep_idx) -> (n_keep, A, 3)

All compute is in clash.cu behind the C-ABI; no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _host
from ._lib import check, lib, ptr, require_cuda, stream_ptr


def _dev(torch):
    return torch.device(f"cuda:{torch.cuda.current_device()}")


def _to_dev(torch, a, dtype, np_dtype):
    if torch.is_tensor(a):
        return a.to(_dev(torch), dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np_dtype)).to(_dev(torch))


def compenetration_check_batch(structures, ids=None, thresh=1.5, max_clashes=0, *, report_near=False,
                               return_tensor=False):
    """compenetration_check over P already-materialised structures (P, A, 3).

    ids None  -> intramolecular: count of 0 < d < 0.5 over the full symmetric matrix (:49-56, :71-72)
    len 2     -> bimolecular (only ids[0] is read, :74-81);   len 3 -> trimolecular (:85-105)
    Returns uint8 verdicts (1 = passes) — numpy unless return_tensor.  With report_near=True also
    returns the number of atom pairs within 1e-9 A of `thresh` (parity diagnostics)."""
    torch = require_cuda()
    S = _to_dev(torch, structures, torch.float64, np.float64)
    if S.dim() != 3 or S.shape[2] != 3:
        raise ValueError("structures must be (P, A, 3)")
    P, A = int(S.shape[0]), int(S.shape[1])
    verdict = torch.empty(max(P, 1), dtype=torch.uint8, device=S.device)
    near = torch.zeros(1, dtype=torch.int64, device=S.device) if report_near else None
    if ids is None:
        F, ids_t = 0, None
    else:
        ids_np = np.asarray(ids).astype(np.int32).ravel()
        F = int(ids_np.size)
        if F not in (2, 3):
            raise ValueError("ids must hold 2 or 3 fragment sizes (or be None)")
        ids_t = torch.from_numpy(ids_np).to(S.device)
    if P:
        check(lib().tsc_clash_structs(ptr(S), P, A, ptr(ids_t), F, _host.sqrt_threshold_image(thresh), float(thresh),
                                      _host.sqrt_threshold_image(0.5), int(max_clashes), ptr(verdict), ptr(near),
                                      stream_ptr()), "tsc_clash_structs")
    out = verdict[:P] if return_tensor else verdict[:P].cpu().numpy()
    return (out, int(near.item())) if report_near else out


def compenetration_check(coords, ids=None, thresh=1.5, max_clashes=0) -> int:
    """Drop-in for tscode.numba_functions.compenetration_check: one structure, returns int 0/1.
    (One kernel launch per call — the batched forms are what the GPU is for.)"""
    return int(compenetration_check_batch(np.asarray(coords, dtype=np.float64)[None], ids, thresh, max_clashes)[0])


class PoseBatch:
    """P rigid-body poses of F fragments, resident in HBM as (conformer id, R, t) per fragment
    — the poses themselves are never materialised unless `gather` is asked for survivors.

    frags : list of F arrays (n_conf_k, n_k, 3)   (mol.atomcoords of each Hypermolecule)
    conf  : (P, F) conformer index per fragment;  R : (P, F, 3, 3);  t : (P, F, 3)
    Pose p is concatenate_k[(R[p,k] @ frags[k][conf[p,k]].T).T + t[p,k]]   (embeds.py:969)."""

    def __init__(self, frags, conf, R, t):
        torch = require_cuda()
        self.torch = torch
        dev = _dev(torch)
        self.F = len(frags)
        if self.F not in (2, 3):
            raise ValueError("PoseBatch supports 2 or 3 fragments")
        n_atoms = np.array([f.shape[1] for f in frags], dtype=np.int32)
        sizes = np.array([int(np.prod(f.shape)) for f in frags], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        if torch.is_tensor(frags[0]):
            self.frag_lib = torch.cat([f.to(dev, dtype=torch.float64).reshape(-1) for f in frags])
        else:
            self.frag_lib = torch.as_tensor(np.concatenate([np.ascontiguousarray(f, dtype=np.float64).ravel()
                                                            for f in frags])).to(dev)
        self.n_conf = [int(f.shape[0]) for f in frags]
        self.n_atoms_np = n_atoms
        self.A = int(n_atoms.sum())
        self.frag_off = torch.from_numpy(off).to(dev)
        self.n_atoms = torch.from_numpy(n_atoms).to(dev)
        self.conf = _to_dev(torch, conf, torch.int32, np.int32)
        self.R = _to_dev(torch, R, torch.float64, np.float64)
        self.t = _to_dev(torch, t, torch.float64, np.float64)
        self.P = int(self.conf.shape[0])
        if tuple(self.conf.shape) != (self.P, self.F) or tuple(self.R.shape) != (self.P, self.F, 3, 3) \
                or tuple(self.t.shape) != (self.P, self.F, 3):
            raise ValueError("conf (P,F), R (P,F,3,3), t (P,F,3) expected")
        self.verdict = torch.empty(max(self.P, 1), dtype=torch.uint8, device=dev)

    def clash(self, thresh=1.5, max_clashes=0, *, report_near=False, lo=0, hi=None):
        """Fused transform + compenetration_check for poses [lo, hi).  Returns the uint8 verdict
        tensor (device) — 1 = pose passes."""
        torch = self.torch
        hi = self.P if hi is None else hi
        n = hi - lo
        near = torch.zeros(1, dtype=torch.int64, device=self.verdict.device) if report_near else None
        if n > 0:
            check(lib().tsc_embed_clash(ptr(self.frag_lib), ptr(self.frag_off), ptr(self.n_atoms), self.F, self.A,
                                        ptr(self.conf[lo:]), ptr(self.R[lo:]), ptr(self.t[lo:]), n,
                                        _host.sqrt_threshold_image(thresh), float(thresh), int(max_clashes),
                                        ptr(self.verdict[lo:]), ptr(near), stream_ptr()), "tsc_embed_clash")
        v = self.verdict[lo:hi]
        return (v, int(near.item())) if report_near else v

    def gather(self, keep_idx=None, n=None):
        """Materialise poses keep_idx (device int64 tensor / array; None = first n poses) as
        (n_keep, A, 3) on the device — get_embed for the survivors only."""
        torch = self.torch
        if keep_idx is None:
            n_keep, kp = (self.P if n is None else int(n)), None
        else:
            kp = _to_dev(torch, keep_idx, torch.int64, np.int64)
            n_keep = int(kp.shape[0])
        out = torch.empty((n_keep, self.A, 3), dtype=torch.float64, device=self.verdict.device)
        if n_keep:
            check(lib().tsc_embed_gather(ptr(self.frag_lib), ptr(self.frag_off), ptr(self.n_atoms), self.F, self.A,
                                         ptr(self.conf), ptr(self.R), ptr(self.t), ptr(kp), n_keep, ptr(out),
                                         stream_ptr()), "tsc_embed_gather")
        return out


def get_embed(mols, conf_ids):
    """Drop-in for tscode.embeds.get_embed (embeds.py:961-969): `mols[k]` needs `.rotation`
    (3,3), `.position` (3,) and `.atomcoords` (n_conf, n_k, 3).  Returns a numpy (sum n_k, 3)."""
    torch = require_cuda()
    F = len(mols)
    if F in (2, 3):
        pb = PoseBatch([np.asarray(m.atomcoords) for m in mols], np.asarray(conf_ids, dtype=np.int32)[None],
                       np.stack([np.asarray(m.rotation, dtype=np.float64) for m in mols])[None],
                       np.stack([np.asarray(m.position, dtype=np.float64) for m in mols])[None])
        return pb.gather(None, 1)[0].cpu().numpy()
    # any other fragment count: compose from single-fragment gathers padded with an empty partner
    parts = []
    for m, c in zip(mols, conf_ids):
        X = np.asarray(m.atomcoords)
        pb = PoseBatch([X, np.zeros((1, 0, 3))], np.array([[c, 0]], dtype=np.int32),
                       np.stack([np.asarray(m.rotation, float), np.eye(3)])[None],
                       np.stack([np.asarray(m.position, float), np.zeros(3)])[None])
        parts.append(pb.gather(None, 1)[0].cpu().numpy())
    return np.concatenate(parts)


def prune_conformers_tfd(structures, quadruplets, thresh=10, verbose=False):
    """Drop-in for tscode.numba_functions.prune_conformers_tfd (numba_functions.py:142-231): removes
    structures whose torsion fingerprints (float32 dihedrals of `quadruplets`) differ from an earlier
    one's by less than `thresh` degrees in total, with the reference's k-ladder, cache and
    connected-component survivor choice.  Returns (structures[mask], mask).

    Fingerprints and the pair test run on the GPU; the grouping loop stops at the first similar later
    structure of every row and caches dissimilar pairs, so it only needs first_hit[i] (tsc_tfd_scan) and
    is replayed on the host exactly as written (torsion_module.ladder_replay_scan, shared with rot_corr —
    the two loops are the same code in the reference)."""
    torch = require_cuda()
    from .torsion_module import ladder_replay_scan
    structures = np.asarray(structures)
    N = structures.shape[0]
    quads = np.ascontiguousarray(np.asarray(quadruplets, dtype=np.int32).reshape(-1, 4))
    Q = quads.shape[0]
    if N == 0:
        return structures, np.ones(0, dtype=bool)
    dev = _dev(torch)
    S = torch.as_tensor(np.ascontiguousarray(structures, dtype=np.float64)).to(dev)
    first = torch.full((N,), N, dtype=torch.int32, device=dev)
    near = torch.zeros(1, dtype=torch.int64, device=dev)
    tf = torch.zeros((N, max(Q, 1)), dtype=torch.float32, device=dev)
    d_quads = torch.from_numpy(quads).to(dev)            # (kept alive until the end of the call: never pass temporaries)
    if Q:
        check(lib().tsc_tfd_fingerprints(ptr(S), N, int(S.shape[1]), ptr(d_quads), Q, ptr(tf), stream_ptr()),
              "tsc_tfd_fingerprints")
    # Q == 0: every sum is 0 < thresh, the scan handles it (row i hits i + 1)
    check(lib().tsc_tfd_scan(ptr(tf), N, Q, float(thresh), ptr(first), ptr(near), stream_ptr()), "tsc_tfd_scan")
    mask, _ = ladder_replay_scan(first.cpu().numpy().astype(np.int64), N, None, verbose=verbose)
    prune_conformers_tfd.last_near_threshold = int(near.item())
    return structures[mask], mask


def torsion_fingerprints(structures, quadruplets):
    """_get_tf_mat (numba_functions.py:233-239): (N, Q) float32 dihedrals in degrees, numpy."""
    torch = require_cuda()
    structures = np.ascontiguousarray(structures, dtype=np.float64)
    quads = np.ascontiguousarray(np.asarray(quadruplets, dtype=np.int32).reshape(-1, 4))
    N, Q = structures.shape[0], quads.shape[0]
    dev = _dev(torch)
    tf = torch.zeros((N, max(Q, 1)), dtype=torch.float32, device=dev)
    d_S, d_quads = torch.from_numpy(structures).to(dev), torch.from_numpy(quads).to(dev)
    if N and Q:
        check(lib().tsc_tfd_fingerprints(ptr(d_S), N, structures.shape[1], ptr(d_quads), Q, ptr(tf), stream_ptr()),
              "tsc_tfd_fingerprints")
    return tf[:, :Q].cpu().numpy()


def _score_embed_poses(structures, constrained_indices, constrained_distances):
    """Drop-in for tscode.numba_functions._score_embed_poses (:273-288): float32 score per structure = sum of
    |distance - desired distance| over that structure's constrained pairs."""
    from .optimization_methods import constraint_scores
    score, _ = constraint_scores(structures, np.asarray(constrained_indices), np.asarray(constrained_distances, dtype=np.float64))
    return score
