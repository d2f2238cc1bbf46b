"""Drop the B200 path into a live TSCoDe: `install_into(tscode)`.

Every caller in the reference binds the hot-path functions with `from module import name`
(embedder.py:48-59, embeds.py:28-33, operators.py:38-43, optimization_methods.py:32,
automep.py:12), so replacing them in the defining module is not enough: the name must be
rebound in each importing module's namespace too.  `uninstall()` restores the originals.
"""
from __future__ import annotations

import importlib
import sys

from . import numba_functions as _nf
from . import optimization_methods as _om
from . import rmsd_pruning as _rp
from . import torsion_module as _tm

# defining module -> {name: replacement}
_PATCHES = {
    "tscode.rmsd_pruning": {
        "prune_conformers_rmsd": _rp.prune_conformers_rmsd,
        "rmsd_and_max_numba": _rp.rmsd_and_max_numba,
        "_rmsd_similarity": _rp._rmsd_similarity,
    },
    "tscode.numba_functions": {
        "compenetration_check": _nf.compenetration_check,
        "prune_conformers_tfd": _nf.prune_conformers_tfd,
        "_score_embed_poses": _nf._score_embed_poses,
    },
    "tscode.optimization_methods": {
        "prune_by_moment_of_inertia": _om.prune_by_moment_of_inertia,
        "fitness_check": _om.fitness_check,
    },
    "tscode.embeds": {
        "get_embed": _nf.get_embed,
    },
    "tscode.torsion_module": {
        "prune_conformers_rmsd_rot_corr": _tm.prune_conformers_rmsd_rot_corr,
    },
}
# modules that import those names with `from ... import`
_IMPORTERS = ("tscode.embedder", "tscode.embeds", "tscode.operators", "tscode.optimization_methods",
              "tscode.automep", "tscode.atropisomer_module", "tscode.multiembed", "tscode.torsion_module")

_saved = []


def install_into(tscode_pkg=None, strict: bool = False):
    """Rebind the reference's hot-path names to the CUDA implementations.  Returns the list of
    (module, name) pairs that were patched.  Modules that are not importable in this
    environment are skipped unless strict=True."""
    patched = []
    originals = {}
    for modname, names in _PATCHES.items():
        try:
            mod = importlib.import_module(modname)
        except Exception:
            if strict:
                raise
            continue
        for name, repl in names.items():
            if hasattr(mod, name):
                originals[name] = getattr(mod, name)
                _saved.append((mod, name, originals[name]))
                setattr(mod, name, repl)
                patched.append((modname, name))
    for modname in _IMPORTERS:
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                if strict:
                    raise
                continue
        for names in _PATCHES.values():
            for name, repl in names.items():
                cur = getattr(mod, name, None)
                if cur is not None and cur is originals.get(name):
                    _saved.append((mod, name, cur))
                    setattr(mod, name, repl)
                    patched.append((modname, name))
    return patched


def uninstall():
    while _saved:
        mod, name, orig = _saved.pop()
        setattr(mod, name, orig)
