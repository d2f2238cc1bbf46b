"""numpy restatement of the reference hot path.  TEST INFRASTRUCTURE — see oracle/__init__.py.

Uses LAPACK's SVD through numpy exactly where the reference does (rmsd_pruning.py:19), so it is
the closest thing to the reference that can travel to the GPU box; pure-Python loops, so only
for small cases.  Cross-checks oracle.c (different SVD) in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

LADDER = (5e5, 2e5, 1e5, 5e4, 2e4, 1e4, 5000, 2000, 1000, 500, 200, 100, 50, 20, 10, 5, 2, 1)


def rmsd_and_max(p, q):
    """rmsd_pruning.py:6-41 — rotation-only Kabsch about the origin, RMSD and max deviation."""
    p = np.asarray(p, float); q = np.asarray(q, float)
    cov = p.T @ q                                        # :15
    v, _, w = np.linalg.svd(cov)                         # :19
    if np.linalg.det(v) * np.linalg.det(w) < 0.0:        # :20-23
        v[:, -1] = -v[:, -1]
    rot = v @ w                                          # :26
    diff = p @ rot - q                                   # :29-32
    rmsd = np.sqrt((diff * diff).sum() / len(diff))      # :35
    maxdev = np.sqrt((diff * diff).sum(axis=1)).max()    # :39
    return float(rmsd), float(maxdev)


def sim_matrix(H, thr):
    """sim[i, j] (i < j) = rmsd < thr and maxdev < 2*thr   (rmsd_pruning.py:75, :95)."""
    N = len(H)
    sim = np.zeros((N, N), bool)
    for i in range(N):
        for j in range(i + 1, N):
            r, d = rmsd_and_max(H[i], H[j])
            sim[i, j] = (r < thr) and (d < 2 * thr)
    return sim


def ladder_model(sim, N, gate=20):
    """SURVEY Appendix A.2 — the elimination of rmsd_pruning.py:81-206 on a precomputed sim
    matrix (callable or array).  Returns (mask, rounds_run)."""
    get = sim if callable(sim) else (lambda i, j: sim[i, j])
    mask = np.ones(N, bool)
    cache = set()
    rounds = []
    for k in LADDER:
        if k == 1 or gate * k < np.count_nonzero(mask):                     # :192
            rounds.append(k)
            cs = int(N // k)                                               # :136
            new = np.zeros(N, bool)
            add = []
            for c in range(int(k)):
                first = c * cs
                last = N if c == k - 1 else cs * (c + 1)                   # :141-144
                for i in range(first, last):
                    if not mask[i]:
                        continue
                    keep = True
                    for j in range(i + 1, last):
                        if mask[j]:
                            key = (first, first + j - i)                   # :65
                            if key in cache:                               # :66-67
                                break
                            if get(i, j):                                  # :75-77
                                add.append(key); keep = False
                                break
                    new[i] = keep
            cache.update(add)                                              # :204
            mask = new
    return mask, rounds


def prune_conformers_rmsd(structures, atomnos, rmsd_thr=0.5):
    """rmsd_pruning.py:164-206."""
    structures = np.asarray(structures)
    H = structures[:, np.asarray(atomnos) != 1]
    memo = {}

    def get(i, j):
        if (i, j) not in memo:
            r, d = rmsd_and_max(H[i], H[j])
            memo[(i, j)] = (r < rmsd_thr) and (d < 2 * rmsd_thr)
        return memo[(i, j)]
    mask, _ = ladder_model(get, len(H))
    return structures[mask], mask


def all_dists(A, B):
    """algebra.py:98-157."""
    d = A[:, None, :] - B[None, :, :]
    return np.sqrt((d * d).sum(-1))


def compenetration_check(coords, ids=None, thresh=1.5, max_clashes=0) -> int:
    """numba_functions.py:59-105."""
    coords = np.asarray(coords, float)
    if ids is None:
        D = all_dists(coords, coords)
        return 0 if np.count_nonzero((D < 0.5) & (D > 0)) > max_clashes else 1       # :49-56, :71-72
    if len(ids) == 2:
        m1, m2 = coords[:ids[0]], coords[ids[0]:]
        return 0 if np.count_nonzero(all_dists(m2, m1) < thresh) > max_clashes else 1  # :74-81
    m1 = coords[:ids[0]]; m2 = coords[ids[0]:ids[0] + ids[1]]; m3 = coords[ids[0] + ids[1]:]
    clashes = 0
    for a, b in ((m2, m1), (m3, m2), (m1, m3)):                                        # :92-103
        clashes += np.count_nonzero(all_dists(a, b) < thresh)
        if clashes > max_clashes:
            return 0
    return 1


def get_embed(frags, conf_ids, R, t):
    """embeds.py:961-969."""
    return np.concatenate([(R[k] @ frags[k][c].T).T + t[k] for k, c in enumerate(conf_ids)])


# --- pose (R, t) builders -------------------------------------------------------------------
def quaternion_to_rotation_matrix(Q):
    """algebra.py:284-323 (scalar-last input)."""
    q0, q1, q2, q3 = Q[3], Q[0], Q[1], Q[2]
    return np.array([[2 * (q0 * q0 + q1 * q1) - 1, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
                     [2 * (q1 * q2 + q0 * q3), 2 * (q0 * q0 + q2 * q2) - 1, 2 * (q2 * q3 - q0 * q1)],
                     [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 2 * (q0 * q0 + q3 * q3) - 1]])


def rot_mat_from_pointer(pointer, angle):
    """algebra.py:325-344 (angle in degrees)."""
    pointer = np.asarray(pointer, float)
    pointer = pointer / np.sqrt((pointer * pointer).sum())
    a = angle * np.pi / 180
    return quaternion_to_rotation_matrix(np.array([np.sin(a / 2) * pointer[0], np.sin(a / 2) * pointer[1],
                                                   np.sin(a / 2) * pointer[2], np.cos(a / 2)]))


def rotation_matrix_from_vectors(vec1, vec2):
    """utils.py:183-208."""
    a = vec1 / np.sqrt((vec1 * vec1).sum()); b = vec2 / np.sqrt((vec2 * vec2).sum())
    v = np.cross(a, b)
    s = np.sqrt((v * v).sum())
    if s != 0:
        c = a @ b
        K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
        return np.eye(3) + K + K @ K * ((1 - c) / s ** 2)
    if np.sqrt(((a + b) ** 2).sum()) == 0:
        return rot_mat_from_pointer(np.array([0., 0., 1.]), 180)
    return np.eye(3)


def string_embed_params(centers, vecs, angles):
    """Pose parameters of the string embed, embeds.py:91-114: loop order conformer pairs (cartesian_product:
    first index major), centre pairs, angles.  Returns (conf (P, 2), R (P, 2, 3, 3), t (P, 2, 3))."""
    (c1, c2), (v1, v2) = centers, vecs
    conf, R, t = [], [], []
    for a in range(c1.shape[0]):
        for b in range(c2.shape[0]):
            for ai1 in range(c1.shape[1]):
                for ai2 in range(c2.shape[1]):
                    for angle in angles:
                        p1, p2, ref_vec, mol_vec = c1[a, ai1], c2[b, ai2], v1[a, ai1], v2[b, ai2]
                        rot = rotation_matrix_from_vectors(mol_vec, -ref_vec)          # :103
                        if angle != 0:                                                # :105-107
                            rot = rot_mat_from_pointer(ref_vec, angle) @ rot
                        conf.append((a, b)); R.append((np.eye(3), rot)); t.append((np.zeros(3), p1 - rot @ p2))   # :109
    return np.array(conf), np.array(R), np.array(t)


def cyclical_embed_params(ref2, tgt2, axis_src, apm, vec_mean, pivot_mean, systematic_angles):
    """Pose parameters of the cyclical embeds, embeds.py:657-709, per group g and angle combination c:
    (R (G*C, F, 3, 3), t (G*C, F, 3)), pose index g*C + c."""
    ref2, tgt2 = np.asarray(ref2, float), np.asarray(tgt2, float)
    G, F = ref2.shape[:2]
    ang = np.asarray(systematic_angles, float).reshape(-1, F)
    R, t = [], []
    for g in range(G):
        for angles in ang:
            Rg, tg = [], []
            for i in range(F):
                A = align_vec_pair(ref2[g, i], tgt2[g, i])                       # :682
                step = rot_mat_from_pointer(A @ axis_src[g][i], angles[i])       # :688-696
                cor = A @ apm[g][i]                                              # :698
                pos = vec_mean[g][i] - A @ pivot_mean[g][i]                      # :706
                Rg.append(step @ A); tg.append(cor - step @ cor + pos)           # :703, :707
            R.append(Rg); t.append(tg)
    return np.array(R), np.array(t)


def align_vec_pair(ref, tgt):
    """algebra.py:258-282."""
    B = np.einsum("ji,jk->ik", np.asarray(ref, float), np.asarray(tgt, float))
    u, s, vh = np.linalg.svd(B)
    if np.linalg.det(u @ vh) < 0:
        u[:, -1] = -u[:, -1]
    return u @ vh


# --- rotor-corrected RMSD (torsion_module.py:953-1161) ------------------------------------------
def rotate_dihedral(coords, torsion, angle, mask):
    """utils.py:389-414 — rotate coords[mask] by `angle` degrees about the i2->i3 bond axis
    (axis = coords[i2] - coords[i3], centre coords[i3]).  Returns a new array."""
    _, i2, i3, _ = torsion
    out = np.array(coords, dtype=float, copy=True)
    mat = rot_mat_from_pointer(out[i2] - out[i3], angle)
    c = out[i3].copy()
    out[mask] = (mat @ (out[mask] - c).T).T + c
    return out


def kabsch_rmsd(P, Q):
    """rmsd==1.4 kabsch_rmsd(P, Q, translate=False): rotation-only Kabsch of P onto Q, then RMSD.
    (third-party, not vendored: restated from the published algorithm)."""
    return rmsd_and_max(P, Q)[0]


def rotationally_corrected_rmsd(ref, coord, heavy_mask, torsions, angles, rot_masks, node_lists):
    """torsion_module.py:953-1011, stateless (works on a copy of `coord`).
    torsions: list of 4-tuples oriented so that the dummy side is last (:1049);
    angles[t]: tuple of degrees (:112-118); rot_masks[t]: bool (A,) = _get_rotation_mask(graph, t)
    (:301-325); node_lists[t]: heavy atoms of the component holding t[1] once every OTHER torsion's
    central bond is cut (:964-977).  Returns (rmsd, corrections, corrected copy of coord)."""
    coord = np.array(coord, dtype=float, copy=True)
    corrections = [0] * len(torsions)
    for t, torsion in enumerate(torsions):
        best = 1e10
        nodes = node_lists[t]
        for angle in angles[t]:                                               # :982-999
            trial = rotate_dihedral(coord, torsion, angle, rot_masks[t])
            local = kabsch_rmsd(ref[nodes], trial[nodes])
            if local < best:
                best = local
                corrections[t] = angle
    coord = apply_rotor_state(coord, torsions, corrections, rot_masks)       # :1004-1008
    return kabsch_rmsd(ref[heavy_mask], coord[heavy_mask]), corrections, coord  # :1011


def apply_rotor_state(coord, torsions, angles_deg, rot_masks):
    """Rotate every rotor by its angle, in torsion order, each about the CURRENT bond axis."""
    for torsion, ang, m in zip(torsions, angles_deg, rot_masks):
        coord = rotate_dihedral(coord, torsion, ang, m)
    return coord


def rotcorr_ladder_model(similar, N, best_angles=None):
    """SURVEY Appendix A.4b — literal restatement of the grouping loop of
    prune_conformers_rmsd_rot_corr (torsion_module.py:1076-1152) on a precomputed boolean matrix
    similar[i, j] = (rot-corrected rmsd(i, j) < max_rmsd), i < j.

    The reference mutates the second structure of every comparison in place (utils.py:412 via
    torsion_module.py:1004-1008): its rotors are left at the angles that best matched the first
    structure *in that structure's current, possibly already mutated, state*.  Because every
    angle set is a full n-fold orbit, the state of rotor t of structure j after comparing (i, j) is
        state[j][t] = (stateless_best_angle(i, j)[t] + state[i][t]) mod 360,
    which this model tracks when best_angles (N, N, T degrees, from the ORIGINAL coordinates) is
    given.  Returns (final_mask, state) — state is (N, T) degrees, all zero without best_angles."""
    import networkx as nx
    final_mask = np.ones(N, dtype=bool)
    cache_set = set()
    T = 0 if best_angles is None else best_angles.shape[2]
    state = np.zeros((N, T))
    for k in LADDER:
        num_active_str = np.count_nonzero(final_mask)
        if k == 1 or 5 * k < num_active_str:                                  # :1083
            d = int(N // k)
            for step in range(int(k)):
                if step == k - 1:
                    _l = len(range(d * step, num_active_str))                  # :1093-1094 (quirk)
                else:
                    _l = len(range(d * step, int(d * (step + 1))))
                matches = set()
                for i_rel in range(_l):
                    for j_rel in range(i_rel + 1, _l):
                        i_abs = i_rel + d * step
                        j_abs = j_rel + d * step
                        if (i_abs, j_abs) not in cache_set:                   # :1107
                            if T:
                                state[j_abs] = (best_angles[i_abs, j_abs] + state[i_abs]) % 360.0
                            if similar[i_abs, j_abs]:                         # :1118
                                matches.add((i_rel, j_rel))
                                break
                            cache_set.add((i_abs, j_abs))
                g = nx.Graph(matches)                                         # node order = set order
                groups = [tuple(g.subgraph(c).nodes) for c in nx.connected_components(g)]
                for group in groups:                                          # :1141-1152
                    for i in set(group) - {group[0]}:
                        final_mask[i + d * step] = 0
    return final_mask, state


# --- TFD and MOI pruning (numba_functions.py:142-268, optimization_methods.py:327-358, algebra.py:24-57, 166-203)
def dihedral(p):
    """algebra.py:24-57 (Praxeolitic formula), degrees."""
    p0, p1, p2, p3 = (np.asarray(x, float) for x in p)
    b0 = -1.0 * (p1 - p0); b1 = p2 - p1; b2 = p3 - p2
    b1 = b1 / np.sqrt((b1 * b1).sum())
    v = b0 - np.dot(b0, b1) * b1
    w = b2 - np.dot(b2, b1) * b1
    return np.degrees(np.arctan2(np.dot(np.cross(b1, v), w), np.dot(v, w)))


def torsion_fingerprints(structures, quadruplets):
    """_get_tf_mat / get_torsion_fingerprint (numba_functions.py:233-239, 258-268): (N, Q) float32."""
    out = np.zeros((len(structures), len(quadruplets)), dtype=np.float32)
    for i, s in enumerate(structures):
        for k, q in enumerate(quadruplets):
            out[i, k] = dihedral([s[q[0]], s[q[1]], s[q[2]], s[q[3]]])
    return out


def tfd_similarity(tfp1, tfp2, thresh=10):
    """numba_functions.py:241-256: float32 difference, wrapped and summed in float64."""
    deltas = np.abs(tfp1 - tfp2)
    deltas = np.abs(deltas.astype(np.float64) - (deltas > 180) * 360)
    return bool(np.sum(deltas) < thresh)


def prune_conformers_tfd(structures, quadruplets, thresh=10):
    """numba_functions.py:142-231: the grouping loop is the one of rot_corr (rotcorr_ladder_model) on the
    TFD similarity matrix."""
    structures = np.asarray(structures)
    tf = torsion_fingerprints(structures, quadruplets)
    N = len(structures)
    d = np.abs(tf[:, None, :] - tf[None, :, :])
    d = np.abs(d.astype(np.float64) - (d > 180) * 360).sum(axis=2)
    mask, _ = rotcorr_ladder_model(np.triu(d < thresh, 1), N)
    return structures[mask], mask


def get_inertia_moments(coords, masses):
    """algebra.py:166-187 (symmetric eigenvalues instead of eig + B^-1 A B: same values to rounding)."""
    coords = np.asarray(coords, float)
    c = coords - (coords * masses[:, None]).sum(axis=0) / masses.sum()
    r2 = (c * c).sum(axis=1)
    I = (masses * r2).sum() * np.eye(3) - np.einsum("n,ni,nj->ij", masses, c, c)
    ev = np.linalg.eigvalsh(I)
    return ev[np.argsort(np.abs(ev))]


def prune_by_moment_of_inertia(structures, atomnos, masses, max_deviation=1e-2):
    """optimization_methods.py:327-358 with get_moi_similarity_matches (algebra.py:189-203); `masses` per atom."""
    import networkx as nx
    structures = np.asarray(structures)
    atomnos = np.asarray(atomnos)
    heavy = atomnos != 1
    mom = np.array([get_inertia_moments(s[heavy], np.asarray(masses, float)[heavy]) for s in structures])
    matches = []
    for i in range(len(structures)):
        for j in range(i + 1, len(structures)):
            if np.all(np.abs(mom[i] - mom[j]) / mom[i] < max_deviation):
                matches.append((i, j))
                break
    G = nx.Graph(matches)
    groups = [tuple(G.subgraph(c).nodes) for c in nx.connected_components(G)]
    mask = np.ones(len(structures), dtype=bool)
    for g in groups:
        for i in set(g) - {g[0]}:
            mask[i] = False
    return structures[mask], mask


def score_embed_poses(structures, constrained_indices, constrained_distances):
    """numba_functions.py:273-288: float32 accumulation of float64 terms."""
    scores = np.zeros(len(structures), dtype=np.float32)
    for j in range(len(structures)):
        for i, (i1, i2) in enumerate(constrained_indices[j]):
            v = structures[j][i1] - structures[j][i2]
            dist = np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
            scores[j] += np.abs(dist - constrained_distances[j][i])
    return scores


def fitness_check(coords, constraints, targets, threshold):
    """optimization_methods.py:544-557."""
    error = 0
    for (a, b), target in zip(constraints, targets):
        if target is not None:
            v = coords[a] - coords[b]
            error += (np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) - target)
    return error < threshold


def xyz_reader_model(text):
    """The XYZ reader behind the reference's read_xyz (utils.py:128-135 = cclib's ccread; cclib==1.7 is pinned in the
    reference's setup.py:50 and is NOT in the reference tree or this image), restated from its published source
    (cclib/io/xyzreader.py, XYZ.generate_repr): lines from str.splitlines(); per frame one optional blank line, the
    atom count = int(first token), the comment line, `count` lines of >= 4 whitespace-separated tokens (symbol, x, y,
    z; further columns ignored); StopIteration anywhere ends the parse, dropping an incomplete frame (its comment is
    kept); symbols are those of the last complete frame; ccData turns the coordinate strings into float64.
    Returns (atomcoords (n, A, 3), symbols, comments).  PARITY UNPINNED against cclib itself (absent); pinned against
    Python's float() and the reference's own write_xyz format (tests/test_capi_and_host.py)."""
    it = iter(text.splitlines())
    all_atomcoords, comments, atomsyms = [], [], []
    while True:
        try:
            line = next(it)
            if line.strip() == '':
                line = next(it)
            tokens = line.split()
            assert len(tokens) >= 1
            natom = int(tokens[0])
            comments.append(next(it))
            lines = []
            for _ in range(natom):
                line = next(it)
                tokens = line.split()
                assert len(tokens) >= 4
                lines.append(tokens)
            assert len(lines) == natom
            atomsyms = [ln[0] for ln in lines]
            all_atomcoords.append([[float(x) for x in ln[1:4]] for ln in lines])
        except StopIteration:
            break
    return np.array(all_atomcoords, dtype=np.float64), atomsyms, comments
