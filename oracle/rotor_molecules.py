"""Idealised rotor-bearing test molecules for rot_corr (SURVEY Appendix A.4 / A.4b builders).
TEST INFRASTRUCTURE: inputs for oracle/gen_golden.py and the tests (pure geometry, no reference
code involved)."""
import numpy as np


def rotmat(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    a = np.deg2rad(ang)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def tetra(center, parent, bl, n, phase):
    """n substituents on `center`, tetrahedral w.r.t. the bond to `parent`, azimuth phase+120k."""
    ax = center - parent
    ax = ax / np.linalg.norm(ax)
    t = np.cross(ax, [1, 0, 0])
    if np.linalg.norm(t) < 1e-3:
        t = np.cross(ax, [0, 1, 0])
    t /= np.linalg.norm(t)
    th = np.deg2rad(180 - 109.47)
    return [center + bl * (np.cos(th) * ax + np.sin(th) * (rotmat(ax, phase + 120 * k) @ t)) for k in range(n)]


def neopentyl_chloride(phi_tbu, a_cl):
    """17 atoms: [Cq, C1(H2Cl), Me x3, Cl, H, H, 9 x H]."""
    cq = np.zeros(3)
    c1 = np.array([0, 0, 1.54])
    me = tetra(cq, c1, 1.54, 3, phi_tbu)
    sub = tetra(c1, cq, 1.0, 3, a_cl)
    cl = c1 + (sub[0] - c1) * 1.79
    h1 = c1 + (sub[1] - c1) * 1.09
    h2 = c1 + (sub[2] - c1) * 1.09
    hs = []
    for m in me:
        hs += tetra(m, cq, 1.09, 3, 60.0)
    coords = np.array([cq, c1] + me + [cl, h1, h2] + hs)
    atomnos = np.array([6, 6, 6, 6, 6, 17, 1, 1] + [1] * 9)
    return coords, atomnos


def di_tbu_benzene(phi1, phi2):
    """36 atoms: 6 ring C, 4 ring H, then per tBu [Cq, 3 Me C, 9 H] (para positions 0 and 3)."""
    ring = np.array([[1.39 * np.cos(np.deg2rad(60 * k)), 1.39 * np.sin(np.deg2rad(60 * k)), 0.0] for k in range(6)])
    hs = [ring[k] * (1.39 + 1.08) / 1.39 for k in (1, 2, 4, 5)]
    atoms = [r for r in ring] + hs
    atomnos = [6] * 6 + [1] * 4
    for pos, phi in ((0, phi1), (3, phi2)):
        cq = ring[pos] * (1.39 + 1.53) / 1.39
        me = tetra(cq, ring[pos], 1.54, 3, phi)
        mh = []
        for m in me:
            mh += tetra(m, cq, 1.09, 3, 60.0)
        atoms += [cq] + me + mh
        atomnos += [6] * 4 + [1] * 9
    return np.array(atoms), np.array(atomnos)


def ensemble_neopentyl(seed, N):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(N):
        c, atomnos = neopentyl_chloride(rng.choice([0, 120, 240]) + rng.normal(0, 3.0), rng.choice([0, 120, 240]))
        out.append(c + rng.normal(0, 0.02, size=c.shape))
    return np.array(out), atomnos


def ensemble_ditbu(seed, N):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(N):
        c, atomnos = di_tbu_benzene(rng.choice([0, 120, 240]) + rng.normal(0, 2.0),
                                    rng.choice([0, 40, 80, 120]) + rng.normal(0, 2.0))
        out.append(c + rng.normal(0, 0.02, size=c.shape))
    return np.array(out), atomnos


def tri_tbu_tripropynylbenzene(phis):
    """63 atoms: 1,3,5-tri-tert-butyl-2,4,6-tri(prop-1-ynyl)benzene.  Order: 6 ring C, then for ring
    positions 0,2,4 a tBu [Cq, 3 Me C, 9 H], then for positions 1,3,5 a propynyl [C, C, C, 3 H]
    pointing radially (uncrowded, so the bond perception sees the intended graph)."""
    ring = np.array([[1.39 * np.cos(np.deg2rad(60 * k)), 1.39 * np.sin(np.deg2rad(60 * k)), 0.0] for k in range(6)])
    atoms = [r for r in ring]
    atomnos = [6] * 6
    for pos, phi in zip((0, 2, 4), phis):
        cq = ring[pos] * (1.39 + 1.53) / 1.39
        me = tetra(cq, ring[pos], 1.54, 3, phi)
        mh = []
        for m in me:
            mh += tetra(m, cq, 1.09, 3, 60.0)
        atoms += [cq] + me + mh
        atomnos += [6] * 4 + [1] * 9
    for pos in (1, 3, 5):
        u = ring[pos] / 1.39
        c1 = u * (1.39 + 1.43); c2 = u * (1.39 + 1.43 + 1.20); c3 = u * (1.39 + 1.43 + 1.20 + 1.46)
        atoms += [c1, c2, c3] + tetra(c3, c2, 1.09, 3, 15.0)
        atomnos += [6, 6, 6, 1, 1, 1]
    return np.array(atoms), np.array(atomnos)


def ensemble_tritbu63(seed, N):
    """BASELINE configs[3] shape (~60 atoms, three symmetric 3-fold rotors)."""
    rng = np.random.default_rng(seed)
    base = np.array([0.0, 40.0, 80.0])
    out = []
    atomnos = None
    for _ in range(N):
        phis = rng.choice(base, size=3) + rng.normal(0, 2.0, size=3) + rng.choice([0, 120, 240], size=3)
        c, atomnos = tri_tbu_tripropynylbenzene(phis)
        out.append(c + rng.normal(0, 0.02, size=c.shape))
    return np.array(out), atomnos
