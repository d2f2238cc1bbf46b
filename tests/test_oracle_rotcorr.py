"""Pin the numpy restatement of rot_corr (oracle_np) against the live reference's fixtures
(stub harness for rmsd==1.4 — see tests/golden/rotcorr.json 'note').  CPU only."""
import json
import os
import sys

import numpy as np
import pytest

from oracle import oracle_np
from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rotor_molecules as rm  # noqa: E402
from tscode_b200.synth import mask_digest  # noqa: E402

FX = json.load(open(os.path.join(GOLDEN, "rotcorr.json")))["fixtures"]
BUILD = {"neopentyl": rm.ensemble_neopentyl, "ditbu": rm.ensemble_ditbu, "tritbu63": rm.ensemble_tritbu63}


def load(name):
    f = FX[name]
    g = np.load(os.path.join(GOLDEN, f"rotcorr_{name}.npz"))
    S, atomnos = BUILD[name.split("_")[0]](f["seed"], f["N"])
    assert np.array_equal(atomnos, g["atomnos"])
    Sc = np.array([s - s.mean(axis=0) for s in S])
    info = dict(torsions=[tuple(t) for t in f["torsions"]], angles=[tuple(a) for a in f["angles"]],
                rot_masks=g["rot_masks"].astype(bool), node_lists=[np.flatnonzero(n) for n in g["node_lists"]],
                heavy=atomnos != 1)
    return f, g, S, Sc, atomnos, info


def rc(info, ref, coord):
    return oracle_np.rotationally_corrected_rmsd(ref, coord, info["heavy"], info["torsions"], info["angles"],
                                                 info["rot_masks"], info["node_lists"])


@pytest.mark.parametrize("name", list(FX))
def test_pair_values_match_reference(name):
    f, g, S, Sc, atomnos, info = load(name)
    for a, b, v, mut in zip(g["pair_i"], g["pair_j"], g["pair_rmsd"], g["pair_mutated"]):
        r, corr, out = rc(info, Sc[a], Sc[b])
        assert abs(r - v) < 1e-9, (name, a, b, r, v)
        assert np.abs(out - mut).max() < 1e-9          # the in-place mutation the reference leaves behind


@pytest.mark.parametrize("name", ["neopentyl_s1", "neopentyl_s2", "ditbu_s0"])
def test_stateless_matrix_plus_ladder_reproduces_reference_mask_and_structures(name):
    f, g, S, Sc, atomnos, info = load(name)
    N = f["N"]
    R = np.full((N, N), np.inf)
    best = np.zeros((N, N, len(info["torsions"])))
    for i in range(N):
        for j in range(i + 1, N):
            R[i, j], best[i, j], _ = rc(info, Sc[i], Sc[j])
    near = int((np.abs(R[np.isfinite(R)] - f["thr"]) < 1e-6).sum())
    mask, state = oracle_np.rotcorr_ladder_model(R < f["thr"], N, best)
    assert near == 0
    assert int(mask.sum()) == f["survivors"] and mask_digest(mask) == f["digest"]
    assert np.array_equal(mask, g["mask"])
    # returned structures: centred, rotors left as found against the last reference compared with
    out = []
    for j in np.flatnonzero(mask):
        out.append(oracle_np.apply_rotor_state(Sc[j], info["torsions"], state[j], info["rot_masks"]))
    out = np.array(out)
    dev = np.abs(out - g["out"]).max()
    print(name, "max |returned - reference| =", dev)
    assert dev < 1e-6


@pytest.mark.parametrize("name", ["neopentyl_s2", "ditbu_s0"])
def test_product_ladder_replay_equals_literal_model(name):
    """tscode_b200.torsion_module.ladder_replay (vectorised host code of the product) against the
    literal model, on the oracle's stateless matrix."""
    from tscode_b200.torsion_module import ladder_replay
    f, g, S, Sc, atomnos, info = load(name)
    N = f["N"]
    R = np.full((N, N), np.inf)
    best = np.zeros((N, N, len(info["torsions"])))
    for i in range(N):
        for j in range(i + 1, N):
            R[i, j], best[i, j], _ = rc(info, Sc[i], Sc[j])
    m0, s0 = oracle_np.rotcorr_ladder_model(R < f["thr"], N, best)
    m1, s1 = ladder_replay(R < f["thr"], N, best)
    assert np.array_equal(m0, m1) and np.array_equal(s0, s1)
    assert mask_digest(m1) == f["digest"]
    # random matrices stress the quirks (last-chunk length from num_active, set/nx survivor choice)
    rng = np.random.default_rng(3)
    for n, dens in ((97, 0.05), (300, 0.01), (750, 0.004)):
        sim = np.triu(rng.random((n, n)) < dens, 1)
        ba = rng.choice([0.0, 120.0, 240.0], size=(n, n, 2))
        a, sa = oracle_np.rotcorr_ladder_model(sim, n, ba)
        b, sb = ladder_replay(sim, n, ba)
        assert np.array_equal(a, b) and np.array_equal(sa, sb)
