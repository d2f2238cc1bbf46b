"""install_into(tscode) rebinding logic — needs the reference tree, which exists only in the build
container (skipped on the GPU box).  No compute is executed."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")
def test_install_rebinds_every_importer():
    ref_harness.install(full=True)
    import tscode.rmsd_pruning as rp
    import tscode.numba_functions as nf
    import tscode.embeds as emb
    import tscode.torsion_module as tm
    from tscode_b200 import install, numba_functions, rmsd_pruning, torsion_module
    orig = rp.prune_conformers_rmsd
    # default: whole-ensemble functions and the orchestrator's loops only — the per-call scalars stay numba's
    patched = install.install_into()
    try:
        assert rp.prune_conformers_rmsd is rmsd_pruning.prune_conformers_rmsd
        assert nf.compenetration_check is not numba_functions.compenetration_check
        assert emb.get_embed is not numba_functions.get_embed
        assert rp.rmsd_and_max_numba is not rmsd_pruning.rmsd_and_max_numba
        from tscode.embedder import RunEmbedding
        assert RunEmbedding.compenetration_refining is install.compenetration_refining
        assert RunEmbedding.fitness_refining is install.fitness_refining
        assert ("tscode.embedder.RunEmbedding", "compenetration_refining") in patched
    finally:
        install.uninstall()
    from tscode.embedder import RunEmbedding
    assert RunEmbedding.compenetration_refining is not install.compenetration_refining
    patched = install.install_into(scalars=True)
    try:
        assert rp.prune_conformers_rmsd is rmsd_pruning.prune_conformers_rmsd
        assert rp._rmsd_similarity is rmsd_pruning._rmsd_similarity
        assert nf.compenetration_check is numba_functions.compenetration_check
        assert emb.get_embed is numba_functions.get_embed
        assert emb.compenetration_check is numba_functions.compenetration_check       # `from ... import` binding
        assert emb._rmsd_similarity is rmsd_pruning._rmsd_similarity
        assert tm.prune_conformers_rmsd_rot_corr is install.prune_conformers_rmsd_rot_corr   # GPU path + own fallback
        assert ("tscode.embeds", "compenetration_check") in patched
        import tscode.optimization_methods as om
        import tscode.operators as ops
        from tscode_b200 import optimization_methods
        assert nf.prune_conformers_tfd is numba_functions.prune_conformers_tfd
        assert tm.prune_conformers_tfd is numba_functions.prune_conformers_tfd          # torsion_module.py:35
        assert ops.prune_conformers_tfd is numba_functions.prune_conformers_tfd         # operators.py:38
        assert om.prune_by_moment_of_inertia is optimization_methods.prune_by_moment_of_inertia
    finally:
        install.uninstall()
    assert rp.prune_conformers_rmsd is orig
    assert emb.compenetration_check is not numba_functions.compenetration_check


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")
def test_install_io_rebinds_xyz_helpers_with_reference_fallback(tmp_path):
    """install_into(io=True): utils.write_xyz / read_xyz and their `from ... import` bindings go to the
    native formatter / reader; a `.xyz` file is read natively, anything else (and a file the native reader rejects) is
    handed to the reference's own read_xyz; off by default; uninstall restores."""
    ref_harness.install(full=True)
    import tscode.utils as ut
    import tscode.hypermolecule_class as hm
    import tscode.optimization_methods as om
    from tscode_b200 import install, utils
    orig_r, orig_w = ut.read_xyz, ut.write_xyz
    install.install_into()
    try:
        assert ut.read_xyz is orig_r and ut.write_xyz is orig_w
    finally:
        install.uninstall()
    patched = install.install_into(io=True)
    try:
        assert ut.read_xyz is install.read_xyz and ut.write_xyz is utils.write_xyz
        assert hm.read_xyz is install.read_xyz and om.write_xyz is utils.write_xyz
        assert ("tscode.hypermolecule_class", "read_xyz") in patched
        S = np.random.default_rng(0).normal(size=(3, 5, 3))
        at = np.array([6, 1, 1, 8, 7])
        fn = tmp_path / "ens.xyz"
        with open(fn, "w") as f:
            for k, c in enumerate(S):
                ut.write_xyz(c, at, f, title=f"frame {k}")
        mol = hm.read_xyz(str(fn))
        assert isinstance(mol, utils.XyzEnsemble) and np.array_equal(mol.atomnos, at)
        assert np.abs(mol.atomcoords - S).max() <= 0.50001e-6 and mol.metadata["comments"] == ["frame 0", "frame 1", "frame 2"]
        log = tmp_path / "calc.out"
        log.write_text("not an xyz file")
        assert not isinstance(ut.read_xyz(str(log)), utils.XyzEnsemble)            # the reference's ccread wrapper
        bad = tmp_path / "bad.xyz"
        bad.write_text("2\nc\nH 0 0 0\nH 0 0\n")
        assert not isinstance(ut.read_xyz(str(bad)), utils.XyzEnsemble)            # rejected natively -> reference
    finally:
        install.uninstall()
    assert ut.read_xyz is orig_r and ut.write_xyz is orig_w and hm.read_xyz is orig_r


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")
def test_install_io_write_structures_writes_the_reference_file(tmp_path, monkeypatch):
    """Embedder.write_structures (embedder.py:996-1043) as patched by install_into(io=True) — one native call for all
    frames — writes byte for byte the file the reference's per-structure loop writes, for both alignments, with and
    without energies, and truncates / logs the same way."""
    import types
    ref_harness.install(full=True)
    from tscode.embedder import Embedder
    from tscode_b200 import install
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(3)
    atomnos = np.array([6, 6, 8, 7, 1, 1, 1, 1, 17, 16, 9, 1])        # elements of the harness periodic table
    base = rng.normal(size=(12, 3)) * 2.0

    def stub(n, let=False):
        st = types.SimpleNamespace()
        st.structures = base[None] + rng.normal(size=(n, 12, 3)) * 0.3
        st.energies = rng.normal(size=n) * 5.0
        st.atomnos = atomnos
        st.options = types.SimpleNamespace(let=let)
        st.stamp = "stamp"
        st.lines = []
        st.log = lambda string='', p=True: st.lines.append(string)
        return st

    ref_method = Embedder.write_structures
    cases = [dict(tag="a", n=7), dict(tag="b", n=7, energies=False, extra=" | extra text"), dict(tag="c", n=5, align="moi"),
             dict(tag="d", n=6, relative=False, p=False), dict(tag="e", n=4, indices=np.array([0, 2, 3])),
             dict(tag="f", n=10003)]                                  # truncated to 10 000 frames, with the log line
    for case in cases:
        n = case.pop("n")
        tag = case.pop("tag")
        state = rng.bit_generator.state
        a = stub(n)
        rng.bit_generator.state = state
        b = stub(n)
        ref_method(a, tag, **case)
        want = open(a.outname, "rb").read()
        os.remove(a.outname)
        install.write_structures(b, tag, **case)
        got = open(b.outname, "rb").read()
        assert got == want and len(want) > 100, tag
        assert a.lines == b.lines and a.outname == b.outname and np.array_equal(a.energies, b.energies)
    patched = install.install_into(io=True)
    try:
        assert Embedder.write_structures is install.write_structures
        assert ("tscode.embedder.Embedder", "write_structures") in patched
    finally:
        install.uninstall()
    assert Embedder.write_structures is ref_method


class _StubRun:
    """The attributes RunEmbedding.compenetration_refining / fitness_refining touch (embedder.py:1119-1134, :973-984,
    :1230-1313), around real Embedder methods."""

    def __init__(self, embed, structures, ids, constrained_indices, targets_by_pair):
        import types
        from tscode.embedder import RunEmbedding
        self.embed, self.structures, self.ids = embed, structures, ids
        self.constrained_indices = constrained_indices
        self.options = types.SimpleNamespace(max_clashes=0, clash_thresh=1.5)
        self.energies = np.zeros(len(structures))
        self.exit_status = np.ones(len(structures), dtype=bool)
        self.lines = []
        self._targets = targets_by_pair
        self.apply_mask = types.MethodType(RunEmbedding.apply_mask, self)

    def log(self, string='', p=True):
        self.lines.append(string)

    def zero_candidates_check(self):
        assert len(self.structures) > 0

    def get_pairing_dists_from_constrained_indices(self, pair):
        return self._targets.get((int(pair[0]), int(pair[1])))


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")
def test_batched_loop_patches_equal_the_reference_methods(monkeypatch):
    """install.compenetration_refining / fitness_refining against the reference's own methods run on the same stub
    `self` (build container: no GPU, so the two batched device calls are stood in for by the oracles — the device calls
    themselves are parity-tested on the GPU): masks, surviving arrays and log lines must be identical."""
    import re
    ref_harness.install(full=True)
    from tscode.embedder import RunEmbedding
    from oracle import oracle_c, oracle_np
    from tscode_b200 import install, numba_functions, optimization_methods
    from tscode_b200.synth import gen_poses, materialise_poses

    monkeypatch.setattr(numba_functions, "compenetration_check_batch",
                        lambda S, ids=None, thresh=1.5, max_clashes=0, **kw: oracle_c.clash_structs(np.asarray(S), ids, thresh, max_clashes))

    def scores(S, cons, targets):
        S, cons = np.asarray(S), np.asarray(cons)
        err = np.zeros(len(S))
        for p in range(len(S)):
            e = 0
            for (a, b), t in zip(cons[p], targets[p]):
                if t is not None:
                    e += np.linalg.norm(S[p][a] - S[p][b]) - t
            err[p] = e
        return np.abs(err).astype(np.float32), err
    monkeypatch.setattr(optimization_methods, "constraint_scores", scores)

    frags, conf, R, t = gen_poses(3, 400, (20, 25), dmin=2.0, dmax=7.0)
    S = materialise_poses(frags, conf, R, t)
    ids = np.array([20, 25])
    rng = np.random.default_rng(0)
    cons = np.stack([np.array([[rng.integers(0, 20), 20 + rng.integers(0, 25)], [rng.integers(0, 20), 20 + rng.integers(0, 25)]])
                     for _ in range(400)])
    targets = {(int(a), int(b)): float(rng.uniform(2.0, 6.0)) for a, b in cons[::2, 0]}       # half of the first pairs have a target
    strip = lambda lines: [re.sub(r"\(\d+ left, .*\)", "(left)", l) for l in lines]
    for embed in ("multiembed", "cyclical"):
        a, b = _StubRun(embed, S.copy(), ids, cons.copy(), targets), _StubRun(embed, S.copy(), ids, cons.copy(), targets)
        RunEmbedding.compenetration_refining(a)
        install.compenetration_refining(b)
        assert np.array_equal(a.structures, b.structures) and np.array_equal(a.constrained_indices, b.constrained_indices)
        assert 0 < len(a.structures) and (embed == "cyclical" or len(a.structures) < 400)
        assert strip(a.lines) == strip(b.lines)
        assert np.array_equal(a.energies, b.energies) and np.array_equal(a.exit_status, b.exit_status)
        for thr in (5, 0.5):
            a2, b2 = (_StubRun(embed, x.structures.copy(), ids, x.constrained_indices.copy(), targets) for x in (a, b))
            RunEmbedding.fitness_refining(a2, threshold=thr, verbose=True)
            install.fitness_refining(b2, threshold=thr, verbose=True)
            assert np.array_equal(a2.structures, b2.structures) and a2.lines == b2.lines
            assert np.array_equal(a2.energies, b2.energies) and len(a2.structures) > 0



@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")
def test_rot_corr_with_too_many_rotors_goes_to_the_reference(monkeypatch):
    """a patched entry point must not raise where the reference worked.  More symmetric rotors than the
    kernels' 64-bit code holds (torsion_module.MAX_T) -> UnsupportedRotors -> the saved original is called with the
    reference's own signature."""
    ref_harness.install(full=True)
    import tscode.torsion_module as tm
    from tscode_b200 import install, torsion_module
    calls = []

    def fake_reference(structures, atomnos, graph, max_rmsd=0.25, verbose=False, logfunction=None):
        calls.append((len(structures), max_rmsd, verbose))
        return structures, np.ones(len(structures), dtype=bool)

    def gpu_path(*a, **k):
        raise torsion_module.UnsupportedRotors("21 rotors")

    monkeypatch.setattr(tm, "prune_conformers_rmsd_rot_corr", fake_reference)
    install.install_into()
    try:
        monkeypatch.setattr(torsion_module, "prune_conformers_rmsd_rot_corr", gpu_path)
        S = np.zeros((3, 4, 3))
        out, mask = tm.prune_conformers_rmsd_rot_corr(S, np.full(4, 6), None, max_rmsd=0.3, verbose=True)
        assert calls == [(3, 0.3, True)] and mask.all() and out is S
    finally:
        install.uninstall()
    assert issubclass(torsion_module.UnsupportedRotors, ValueError)
