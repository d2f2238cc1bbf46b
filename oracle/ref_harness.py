"""Import harness for the LIVE reference (/root/reference) — golden-vector generation only.

TEST INFRASTRUCTURE.  Never imported by the product (tscode_b200/), by bench.py or by the
`-m gpu` tests: /root/reference does not exist on the GPU box.  oracle/gen_golden.py uses it
in the build container to run the unmodified reference functions and freeze their outputs
under tests/golden/.

Stubs follow SURVEY.md Appendix A.6: MagicMock modules for the third-party packages the
hot path never calls (cclib, ase, sella, matplotlib, periodictable, openbabel, _tkinter), a
7-element periodic table, and a functional stand-in for the un-vendored `rmsd==1.4`
package (rotation-only Kabsch, translate=False default) which rot_corr calls
(torsion_module.py:24,989,1011).
"""
import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("TSCODE_REFERENCE", "/root/reference")

_RMSD_STUB = '''
import numpy as np
def kabsch(P, Q):
    C = np.dot(np.transpose(P), Q); V, S, W = np.linalg.svd(C)
    if (np.linalg.det(V) * np.linalg.det(W)) < 0.0: S[-1] = -S[-1]; V[:, -1] = -V[:, -1]
    return np.dot(V, W)
def kabsch_rotate(P, Q): return np.dot(P, kabsch(P, Q))
def rmsd(V, W): d = np.array(V) - np.array(W); return np.sqrt((d * d).sum() / len(V))
def kabsch_rmsd(P, Q, W=None, translate=False):
    if translate: Q = Q - Q.mean(axis=0); P = P - P.mean(axis=0)
    return rmsd(kabsch_rotate(P, Q), Q)
'''


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tscode"))


def install(full: bool = False):
    """Put the reference on sys.path.  full=True also installs the stub modules needed to
    import tscode.numba_functions / torsion_module / embeds / utils."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not full:
        return
    try:
        import rmsd  # noqa: F401
    except ImportError:
        d = tempfile.mkdtemp(prefix="rmsd_stub_")
        os.makedirs(os.path.join(d, "rmsd"))
        with open(os.path.join(d, "rmsd", "__init__.py"), "w") as f:
            f.write(_RMSD_STUB)
        sys.path.insert(0, d)
    for name in ['_tkinter', 'cclib', 'cclib.io', 'ase', 'ase.calculators', 'ase.calculators.calculator',
                 'ase.calculators.gaussian', 'ase.calculators.mopac', 'ase.calculators.orca', 'ase.constraints',
                 'ase.dyneb', 'ase.optimize', 'ase.vibrations', 'ase.atoms', 'ase.gui', 'ase.gui.gui',
                 'ase.gui.images', 'sella', 'matplotlib', 'matplotlib.pyplot', 'periodictable',
                 'periodictable.core', 'periodictable.covalent_radius', 'periodictable.mass', 'openbabel',
                 'prettytable']:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()

    class _El:
        def __init__(s, sym, r, m):
            s.symbol, s.covalent_radius, s.mass = sym, r, m

    ptmod = types.ModuleType('tscode.pt')
    ptmod.pt = {1: _El('H', 0.31, 1.008), 6: _El('C', 0.76, 12.011), 7: _El('N', 0.71, 14.007),
                8: _El('O', 0.66, 15.999), 9: _El('F', 0.57, 18.998), 16: _El('S', 1.05, 32.06),
                17: _El('Cl', 1.02, 35.45)}
    import tscode  # noqa: F401
    sys.modules['tscode.pt'] = ptmod
    import networkx as nx
    if not hasattr(nx, 'from_numpy_matrix'):
        nx.from_numpy_matrix = nx.from_numpy_array
