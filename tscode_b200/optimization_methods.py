"""B200-native drop-in for the one hot-path function of tscode/optimization_methods.py:

    prune_by_moment_of_inertia(structures, atomnos, max_deviation=1e-2) -> (structures[mask], mask)   # :327-358

It runs right before the RMSD prune in Embedder.similarity_refining (embedder.py:1349).  Principal moments
of the heavy atoms and the first-match pair scan run on the GPU (tfd_moi.cu); the connected-component
survivor choice is the reference's networkx code, verbatim.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from ._lib import check, lib, ptr, require_cuda, stream_ptr

# standard atomic masses (u) for Z = 1..54, as the `periodictable` package the reference reads them from
_MASS = [0.0, 1.00794, 4.002602, 6.941, 9.012182, 10.811, 12.0107, 14.0067, 15.9994, 18.9984032, 20.1797, 22.98977,
         24.305, 26.981538, 28.0855, 30.973761, 32.065, 35.453, 39.948, 39.0983, 40.078, 44.95591, 47.867, 50.9415,
         51.9961, 54.938049, 55.845, 58.9332, 58.6934, 63.546, 65.409, 69.723, 72.64, 74.9216, 78.96, 79.904, 83.798,
         85.4678, 87.62, 88.90585, 91.224, 92.90638, 95.94, 98.0, 101.07, 102.9055, 106.42, 107.8682, 112.411, 114.818,
         118.71, 121.76, 127.6, 126.90447, 131.293]


def _masses_for(atomnos):
    try:                                   # the drop-in scenario: the reference's own table
        from tscode.pt import pt
        return np.array([pt[int(a)].mass for a in atomnos], dtype=np.float64)
    except Exception:
        return np.array([_MASS[int(a)] for a in atomnos], dtype=np.float64)


def moments_of_inertia(structures, atomnos, masses=None):
    """get_inertia_moments (algebra.py:166-187) of every structure's heavy atoms: (N, 3) numpy, ascending."""
    torch = require_cuda()
    structures = np.ascontiguousarray(structures, dtype=np.float64)
    atomnos = np.asarray(atomnos)
    heavy = np.flatnonzero(atomnos != 1).astype(np.int32)
    m = np.ascontiguousarray(_masses_for(atomnos[heavy]) if masses is None else np.asarray(masses, dtype=np.float64)[heavy])
    N = structures.shape[0]
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    out = torch.zeros((max(N, 1), 3), dtype=torch.float64, device=dev)
    d_S, d_heavy, d_m = torch.from_numpy(structures).to(dev), torch.from_numpy(heavy).to(dev), torch.from_numpy(m).to(dev)
    if N and heavy.size:
        check(lib().tsc_moi_moments(ptr(d_S), N, structures.shape[1], ptr(d_heavy), int(heavy.size), ptr(d_m), ptr(out),
                                    stream_ptr()), "tsc_moi_moments")
    return out[:N]


def prune_by_moment_of_inertia(structures, atomnos, max_deviation=1e-2, *, masses=None):
    """Drop-in for tscode.optimization_methods.prune_by_moment_of_inertia (:327-358).  `masses` (per atom,
    optional) overrides the atomic-mass table."""
    import networkx as nx
    torch = require_cuda()
    structures = np.asarray(structures)
    N = structures.shape[0]
    mask = np.ones(N, dtype=bool)
    if N == 0:
        return structures, mask
    mom = moments_of_inertia(structures, atomnos, masses)
    first = torch.full((N,), N, dtype=torch.int32, device=mom.device)
    near = torch.zeros(1, dtype=torch.int64, device=mom.device)
    check(lib().tsc_moi_scan(ptr(mom), N, float(max_deviation), ptr(first), ptr(near), stream_ptr()), "tsc_moi_scan")
    fh = first.cpu().numpy()
    rows = np.flatnonzero(fh < N)
    matches = [(int(i), int(fh[i])) for i in rows]                  # np.where order: ascending rows (algebra.py:200-201)
    G = nx.Graph(matches)                                           # optimization_methods.py:341-355, verbatim
    subgraphs = [G.subgraph(c) for c in nx.connected_components(G)]
    groups = [tuple(graph.nodes) for graph in subgraphs]
    best_of_cluster = [group[0] for group in groups]
    rejects_sets = [set(a) - {b} for a, b in zip(groups, best_of_cluster)]
    for _s in rejects_sets:
        for i in _s:
            mask[i] = False
    prune_by_moment_of_inertia.last_near_threshold = int(near.item())
    return structures[mask], mask


def constraint_scores(structures, constraints, targets):
    """Batched distance-constraint scores: (score_abs float32 (P,), error_signed float64 (P,)) numpy.
    constraints: (K, 2) shared by all structures or (P, K, 2) per structure; targets: (K,) or (P, K), None / NaN for a
    constraint without target."""
    torch = require_cuda()
    S = np.ascontiguousarray(structures, dtype=np.float64)
    P, A = S.shape[0], S.shape[1]
    cons = np.ascontiguousarray(np.asarray(constraints, dtype=np.int32))
    tg = np.array([[np.nan if v is None else v for v in row] for row in np.atleast_2d(np.asarray(targets, dtype=object))],
                  dtype=np.float64)
    per_pose = cons.ndim == 3
    K = cons.shape[-2] if cons.size else 0
    tg = np.ascontiguousarray(tg.reshape(P, K) if per_pose else tg.reshape(K))
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    d_S, d_c, d_t = torch.from_numpy(S).to(dev), torch.from_numpy(cons.reshape(-1)).to(dev), torch.from_numpy(tg.reshape(-1)).to(dev)
    score = torch.zeros(max(P, 1), dtype=torch.float32, device=dev)
    err = torch.zeros(max(P, 1), dtype=torch.float64, device=dev)
    if P:
        check(lib().tsc_constraint_scores(ptr(d_S), P, A, ptr(d_c), ptr(d_t), K, 1 if per_pose else 0, ptr(score), ptr(err),
                                          stream_ptr()), "tsc_constraint_scores")
    return score[:P].cpu().numpy(), err[:P].cpu().numpy()


def fitness_check(coords, constraints, targets, threshold) -> bool:
    """Drop-in for tscode.optimization_methods.fitness_check (:544-557): True if the signed sum of
    (distance - target) over the constraints with a target is below `threshold`.  (One launch per call: use
    constraint_scores for a whole ensemble, as Embedder.fitness_refining's loop should, embedder.py:1283-1290.)"""
    if len(constraints) == 0:
        return bool(0 < threshold)
    _, err = constraint_scores(np.asarray(coords, dtype=np.float64)[None], np.asarray(constraints), list(targets))
    return bool(err[0] < threshold)
