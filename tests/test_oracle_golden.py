"""Pin the oracle (oracle.c and oracle_np.py) against outputs of the LIVE reference frozen in
tests/golden/ by oracle/gen_golden.py.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle_c, oracle_np
from tscode_b200.synth import gen_ensemble, gen_poses, materialise_poses, mask_digest

from conftest import GOLDEN


def _atomnos(r):
    a = np.full(r["M"], 6)
    if r.get("mixed_h"):
        a[np.random.default_rng(r["seed"]).random(r["M"]) < 0.3] = 1
    return a


def test_rmsd_and_max_matches_reference_pairs():
    g = np.load(os.path.join(GOLDEN, "rmsd_pairs.npz"))
    out = g["out"]
    for i in range(len(out)):
        p, q = g[f"p{i}"], g[f"q{i}"]
        rc, dc = oracle_c.rmsd_and_max(p, q)
        rn, dn = oracle_np.rmsd_and_max(p, q)
        M = len(p)
        if M == 3 and np.linalg.svd(p.T @ q)[1][2] < 1e-9:
            continue
        # RMSD within 1e-9 A of the reference (north star); in practice ~1e-14
        assert abs(rc - out[i, 0]) < 1e-9 and abs(dc - out[i, 1]) < 1e-9, (i, M, rc, out[i])
        assert abs(rn - out[i, 0]) < 1e-12 and abs(dn - out[i, 1]) < 1e-12


_rows = json.load(open(os.path.join(GOLDEN, "prune_masks.json")))["rows"]


@pytest.mark.parametrize("r", _rows, ids=[f"s{r['seed']}_N{r['N']}_M{r['M']}" for r in _rows])
def test_prune_mask_matches_reference(r):
    S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"])
    out, mask = oracle_c.prune_conformers_rmsd(S, _atomnos(r), r["thr"])
    ref = np.unpackbits(np.frombuffer(bytes.fromhex(r["mask_hex"]), np.uint8))[:r["N"]].astype(bool)
    assert mask.sum() == r["survivors"]
    assert np.array_equal(mask, ref)
    assert mask_digest(mask) == r["digest"]
    assert np.array_equal(out, S[mask])


_aniso = json.load(open(os.path.join(GOLDEN, "prune_masks_aniso.json")))["rows"]


@pytest.mark.parametrize("r", _aniso, ids=[f"s{r['seed']}_N{r['N']}_M{r['M']}" for r in _aniso])
def test_prune_mask_matches_reference_anisotropic_molecules(r):
    """Elongated / planar / rod-like molecules (fixtures from the live reference, oracle/gen_golden.py --only
    prune_aniso): the shapes on which the pre-screen's Samuelson bound excludes nothing."""
    S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"], scale=np.array(r["scale"]))
    out, mask = oracle_c.prune_conformers_rmsd(S, _atomnos(r), r["thr"])
    ref = np.unpackbits(np.frombuffer(bytes.fromhex(r["mask_hex"]), np.uint8))[:r["N"]].astype(bool)
    assert np.array_equal(mask, ref) and mask_digest(mask) == r["digest"] and np.array_equal(out, S[mask])


def test_prune_big_anisotropic_digest_matches_reference():
    """20 000 x 80 with a planar base molecule: C oracle against the live reference's digest (the 50 000 x 80
    elongated row of the same fixture file was checked the same way in the build container: 11 s)."""
    r = json.load(open(os.path.join(GOLDEN, "prune_masks_aniso_big.json")))["rows"][0]
    S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"], scale=np.array(r["scale"]))
    mask, _, _ = oracle_c.prune_heavy(S, r["thr"])
    assert int(mask.sum()) == r["survivors"] and mask_digest(mask) == r["digest"]


def test_prune_numpy_oracle_small():
    for r in _rows:
        if r["N"] > 400:
            continue
        S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"])
        _, mask = oracle_np.prune_conformers_rmsd(S, _atomnos(r), r["thr"])
        assert mask_digest(mask) == r["digest"], r["seed"]


def test_ladder_replay_on_precomputed_sim():
    """orc_prune_rmsd with a precomputed sim matrix == lazy evaluation == numpy ladder model."""
    r = _rows[5]
    S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"])
    H = oracle_c.heavy(S, _atomnos(r))
    sim = oracle_c.sim_rows(H, r["thr"], 0, r["N"])
    m1, ne, rounds = oracle_c.prune_heavy(H, r["thr"], sim_bytes=sim)
    assert ne == 0 and mask_digest(m1) == r["digest"]
    m2, rounds2 = oracle_np.ladder_model(sim.astype(bool), r["N"])
    assert np.array_equal(m1, m2) and rounds == list(rounds2)


def test_rmsd_similarity_matches_reference():
    g = json.load(open(os.path.join(GOLDEN, "rmsd_similarity.json")))
    S = gen_ensemble(g["seed"], g["N"], g["M"], g["n_clusters"], sigma_noise=g["sigma_noise"])
    out = [oracle_c.rmsd_similarity(S[i], S[i + 1:i + 1 + g["window"]], g["rmsd_thr"]) for i in range(len(g["out"]))]
    assert out == g["out"]


_clash = json.load(open(os.path.join(GOLDEN, "clash_verdicts.json")))


@pytest.mark.parametrize("r", _clash["rows"], ids=[f"s{r['seed']}_P{r['P']}" for r in _clash["rows"]])
def test_clash_verdicts_match_reference(r):
    frags, conf, R, t = gen_poses(r["seed"], r["P"], tuple(r["n_atoms"]))
    v = oracle_c.embed_clash_batch(frags, conf, R, t, r["thresh"], r["max_clashes"])
    ref = np.unpackbits(np.frombuffer(bytes.fromhex(r["verdict_hex"]), np.uint8))[:r["P"]]
    assert int(v.sum()) == r["passes"]
    assert np.array_equal(v, ref)
    assert mask_digest(v) == r["digest"]
    # materialised route + numpy oracle on a slice
    sel = np.arange(0, min(r["P"], 300))
    S = materialise_poses(frags, conf, R, t, sel)
    v2 = oracle_c.clash_structs(S, np.array(r["n_atoms"]), r["thresh"], r["max_clashes"])
    assert np.array_equal(v2, ref[sel])
    v3 = [oracle_np.compenetration_check(S[i], np.array(r["n_atoms"]), r["thresh"], r["max_clashes"]) for i in range(100)]
    assert np.array_equal(np.array(v3), ref[:100])


def test_clash_ids_none_matches_reference():
    g = _clash["ids_none"]
    frags, conf, R, t = gen_poses(g["seed"], g["P"], tuple(g["n_atoms"]), blob=g["blob"], dmin=g["dmin"], dmax=g["dmax"])
    S = materialise_poses(frags, conf, R, t)
    for row in g["rows"]:
        ref = np.unpackbits(np.frombuffer(bytes.fromhex(row["verdict_hex"]), np.uint8))[:g["P"]]
        v = oracle_c.clash_structs(S, None, 1.5, row["max_clashes"])
        assert np.array_equal(v, ref)
        assert [oracle_np.compenetration_check(S[i], None, 1.5, row["max_clashes"]) for i in range(50)] == list(ref[:50])


def test_get_embed_and_rotation_builders():
    frags, conf, R, t = gen_poses(0, 16, (5, 9, 4))
    for p in range(16):
        a = oracle_c.get_embed(frags, conf[p], R[p], t[p])
        b = oracle_np.get_embed(frags, conf[p], R[p], t[p])
        assert np.abs(a - b).max() < 1e-13
    g = np.load(os.path.join(GOLDEN, "rotation_builders.npz"))
    for i in range(len(g["v1"])):
        assert np.abs(oracle_np.rotation_matrix_from_vectors(g["v1"][i], g["v2"][i]) - g["rmv"][i]).max() < 1e-12
        assert np.abs(oracle_np.rot_mat_from_pointer(g["v1"][i], g["ang"][i]) - g["rmp"][i]).max() < 1e-12
        assert np.abs(oracle_np.align_vec_pair(g["ref"][i], g["tgt"][i]) - g["avp"][i]).max() < 1e-10


def test_oracle_dedup_groups_vs_live_reference():
    """(f)-1: clash test + group-local `_rmsd_similarity` de-duplication of the cyclical embeds
    (embeds.py:713-718), oracle loop against the masks the live reference produced."""
    from tscode_b200.synth import gen_pose_groups, materialise_poses
    rows = json.load(open(os.path.join(GOLDEN, "dedup_groups.json")))["rows"]
    for r in rows:
        frags, conf, R, t, gid = gen_pose_groups(r["seed"], r["n_groups"], r["steps"], tuple(r["n_atoms"]))
        S = materialise_poses(frags, conf, R, t)
        ids = np.array(r["n_atoms"])
        want_pass = np.unpackbits(np.frombuffer(bytes.fromhex(r["passed_hex"]), np.uint8))[:len(S)]
        want_keep = np.unpackbits(np.frombuffer(bytes.fromhex(r["keep_hex"]), np.uint8))[:len(S)]
        keep, passed = np.zeros(len(S), np.uint8), np.zeros(len(S), np.uint8)
        for g in range(r["n_groups"]):
            kept = []
            for p in np.flatnonzero(gid == g):
                if oracle_c.compenetration_check(S[p], ids, r["thresh"], 0):
                    passed[p] = 1
                    if not oracle_c.rmsd_similarity(S[p], kept, r["rmsd_thr"]):
                        kept.append(S[p]); keep[p] = 1
        assert np.array_equal(passed, want_pass) and np.array_equal(keep, want_keep)


def test_oracle_string_embed_params_vs_live_reference():
    """(f)-2: pose parameters of the string embed (embeds.py:91-114), numpy oracle vs the live reference's builders,
    including the parallel (identity) and antiparallel (180 degree flip) branches of rotation_matrix_from_vectors."""
    g = np.load(os.path.join(GOLDEN, "string_embed_params.npz"))
    conf, R, t = oracle_np.string_embed_params((g["c1"], g["c2"]), (g["v1"], g["v2"]), list(g["angles"]))
    assert R.shape[0] == g["R"].shape[0] == 3 * 2 * 2 * 3 * 6
    assert np.abs(R[:, 1] - g["R"]).max() < 1e-14 and np.abs(t[:, 1] - g["t"]).max() < 1e-13
    assert np.array_equal(R[:, 0], np.broadcast_to(np.eye(3), R[:, 0].shape)) and not t[:, 0].any()


_tm = json.load(open(os.path.join(GOLDEN, "tfd_moi.json")))["rows"]


@pytest.mark.parametrize("r", _tm, ids=[f"{r['kind']}{r['seed']}" for r in _tm])
def test_oracle_tfd_moi_pruning_vs_live_reference(r):
    """(f)-3: prune_conformers_tfd / prune_by_moment_of_inertia, numpy oracle vs the live reference's masks."""
    S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"])
    want = np.unpackbits(np.frombuffer(bytes.fromhex(r["mask_hex"]), np.uint8))[:r["N"]].astype(bool)
    if r["kind"] == "tfd":
        tf = oracle_np.torsion_fingerprints(S[:1], r["quads"])
        assert np.array_equal(tf[0], np.array(r["tf_row0"], dtype=np.float32))
        _, mask = oracle_np.prune_conformers_tfd(S, r["quads"], r["thresh"])
    else:
        atomnos, masses = np.array(r["atomnos"]), np.array(r["masses"])
        mom = oracle_np.get_inertia_moments(S[0][atomnos != 1], masses[atomnos != 1])
        assert np.allclose(mom, r["moments_row0"], rtol=1e-12)
        _, mask = oracle_np.prune_by_moment_of_inertia(S, atomnos, masses, r["max_deviation"])
    assert mask_digest(mask) == r["digest"] and np.array_equal(mask, want)


def test_oracle_cyclical_embed_params_vs_live_reference():
    """(f)-2: pose parameters of the cyclical embeds (embeds.py:657-709), numpy oracle vs the live reference's
    align_vec_pair / rot_mat_from_pointer."""
    g = np.load(os.path.join(GOLDEN, "cyclical_embed_params.npz"))
    R, t = oracle_np.cyclical_embed_params(g["ref2"], g["tgt2"], g["axis_src"], g["apm"], g["vmean"], g["pmean"],
                                           g["sys_angles"])
    assert R.shape == g["R"].shape == (7 * 27, 3, 3, 3)
    assert np.abs(R - g["R"]).max() < 1e-13 and np.abs(t - g["t"]).max() < 1e-12


def test_oracle_constraint_scores_vs_live_reference():
    g = json.load(open(os.path.join(GOLDEN, "constraint_scores.json")))
    S = gen_ensemble(g["seed"], g["N"], g["M"], g["n_clusters"], sigma_noise=g["sigma_noise"])
    cons, dists = np.array(g["cons"]), np.array(g["dists"])
    sc = oracle_np.score_embed_poses(S, cons, dists)
    assert np.array_equal(sc, np.array(g["scores"], dtype=np.float32))
    for p in range(g["N"]):
        tg = [None if (p + k) % 5 == 0 else float(dists[p, k]) for k in range(3)]
        assert bool(oracle_np.fitness_check(S[p], [tuple(c) for c in cons[p]], tg, g["fitness_threshold"])) == g["fitness"][p]


def test_oracle_embed_pipeline_small_vs_live_reference():
    """BASELINE configs[4] in small: the oracles' restatement of the generator loop (pose parameters, get_embed,
    compenetration_check, group-local _rmsd_similarity, prune_conformers_rmsd) reproduces every bit the live
    reference produced (oracle/gen_golden_c5.py, tests/golden/embed_pipeline.json)."""
    import json
    import os
    from oracle import oracle_np
    from tscode_b200.synth import gen_cyclical_groups, materialise_poses
    r = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "embed_pipeline.json")))["rows"]["small"]
    d = gen_cyclical_groups(r["seed"], r["n_groups"])
    R, t = oracle_np.cyclical_embed_params(d["ref2"], d["tgt2"], d["axis_src"], d["atomic_pivot_mean"], d["vec_mean"],
                                           d["pivot_mean"], d["systematic_angles"])
    G, C = r["n_groups"], d["systematic_angles"].shape[0]
    conf = np.repeat(d["group_conf"], C, axis=0)
    v = oracle_c.embed_clash_batch(d["frags"], conf, R, t, 1.5, 0)
    unpack = lambda h, n: np.unpackbits(np.frombuffer(bytes.fromhex(h), np.uint8))[:n].astype(bool)
    assert np.array_equal(v.astype(bool), unpack(r["verdict_hex"], r["poses"]))
    idx = np.flatnonzero(v)
    P = materialise_poses(d["frags"], conf, R, t, idx)
    gid = np.repeat(np.arange(G), C)[idx]
    kept = np.zeros(r["poses"], bool)
    for g in np.unique(gid):
        members, keep = np.flatnonzero(gid == g), []
        for i in members:
            if not keep or not oracle_c.rmsd_similarity(P[i], P[keep], 1.0):
                keep.append(i)
                kept[idx[i]] = True
    assert np.array_equal(kept, unpack(r["kept_hex"], r["poses"]))
    A = P.shape[1]
    _, mask = oracle_c.prune_conformers_rmsd(P[kept[idx]], np.full(A, 6), 0.5)
    assert np.array_equal(mask, unpack(r["mask_hex"], r["kept"]))
