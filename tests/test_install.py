"""install_into(tscode) rebinding logic — needs the reference tree, which exists only in the build
container (skipped on the GPU box).  No compute is executed."""
import os
import sys

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
    patched = install.install_into()
    try:
        assert rp.prune_conformers_rmsd is rmsd_pruning.prune_conformers_rmsd
        assert rp._rmsd_similarity is rmsd_pruning._rmsd_similarity
        assert nf.compenetration_check is numba_functions.compenetration_check
        assert emb.get_embed is numba_functions.get_embed
        assert emb.compenetration_check is numba_functions.compenetration_check       # `from ... import` binding
        assert emb._rmsd_similarity is rmsd_pruning._rmsd_similarity
        assert tm.prune_conformers_rmsd_rot_corr is torsion_module.prune_conformers_rmsd_rot_corr
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
