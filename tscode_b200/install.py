"""Drop the B200 path into a live TSCoDe: `install_into(tscode)`.

Every caller in the reference binds the hot-path functions with `from module import name`
(embedder.py:48-59, embeds.py:28-33, operators.py:38-43, optimization_methods.py:32,
automep.py:12), so replacing them in the defining module is not enough: the name must be
rebound in each importing module's namespace too.  `uninstall()` restores the originals.

What is patched by default is what the GPU is for — whole-ensemble operations:

  * the pruners  prune_conformers_rmsd, prune_conformers_rmsd_rot_corr, prune_conformers_tfd,
    prune_by_moment_of_inertia, _score_embed_poses;
  * the two per-structure LOOPS of the orchestrator, RunEmbedding.compenetration_refining
    (embedder.py:1230-1268: `for structure in self.structures: compenetration_check(...)`) and
    RunEmbedding.fitness_refining (:1270-1313: `fitness_check` per structure), re-stated with ONE batched call each
    (compenetration_check_batch / constraint_scores); everything around the loop — logging, apply_mask,
    zero_candidates_check — is the reference's own, line by line.

The per-call scalars (compenetration_check, get_embed, fitness_check, rmsd_and_max_numba, _rmsd_similarity) are NOT
patched unless `scalars=True`: one structure per call means one upload, one launch and one read-back per call
(~60-100 us measured, profiles/r02_scalar_latency.json) against ~13 us for the numba originals, which the
generators call once per pose inside Python loops (embeds.py:116-118, 713-715).  Those loops are replaced as a whole
by tscode_b200.embeds.cyclical_embed_pipeline / string_embed_poses, not call by call.
"""
from __future__ import annotations

import importlib
import sys
import time

import numpy as np

from . import numba_functions as _nf
from . import optimization_methods as _om
from . import rmsd_pruning as _rp
from . import torsion_module as _tm
from . import utils as _ut

# defining module -> {name: replacement}
_PATCHES = {
    "tscode.rmsd_pruning": {
        "prune_conformers_rmsd": _rp.prune_conformers_rmsd,
    },
    "tscode.numba_functions": {
        "prune_conformers_tfd": _nf.prune_conformers_tfd,
        "_score_embed_poses": _nf._score_embed_poses,
    },
    "tscode.optimization_methods": {
        "prune_by_moment_of_inertia": _om.prune_by_moment_of_inertia,
    },
    "tscode.torsion_module": {
        "prune_conformers_rmsd_rot_corr": None,          # filled in below: GPU path with the reference as its own fallback
    },
}
_reference_rot_corr = []                                  # the original, saved by install_into


def prune_conformers_rmsd_rot_corr(structures, atomnos, graph, max_rmsd=0.25, verbose=False, logfunction=None, **kw):
    """tscode_b200.torsion_module.prune_conformers_rmsd_rot_corr; a molecule with more symmetric rotors (or rotor
    angles) than the kernels' 64-bit code holds (torsion_module.MAX_T / MAX_ANG) is handed to the reference's own
    function instead of raising from a patched entry point."""
    try:
        return _tm.prune_conformers_rmsd_rot_corr(structures, atomnos, graph, max_rmsd, verbose, logfunction, **kw)
    except _tm.UnsupportedRotors:
        if not _reference_rot_corr:
            raise
        return _reference_rot_corr[0](structures, atomnos, graph, max_rmsd=max_rmsd, verbose=verbose, logfunction=logfunction)


_PATCHES["tscode.torsion_module"]["prune_conformers_rmsd_rot_corr"] = prune_conformers_rmsd_rot_corr
_SCALAR_PATCHES = {
    "tscode.rmsd_pruning": {
        "rmsd_and_max_numba": _rp.rmsd_and_max_numba,
        "_rmsd_similarity": _rp._rmsd_similarity,
    },
    "tscode.numba_functions": {
        "compenetration_check": _nf.compenetration_check,
    },
    "tscode.optimization_methods": {
        "fitness_check": _om.fitness_check,
    },
    "tscode.embeds": {
        "get_embed": _nf.get_embed,
    },
}
# modules that import those names with `from ... import`
_IMPORTERS = ("tscode.embedder", "tscode.embeds", "tscode.operators", "tscode.optimization_methods",
              "tscode.automep", "tscode.atropisomer_module", "tscode.multiembed", "tscode.torsion_module")
# `from tscode.utils import read_xyz / write_xyz` (and pka.py:23, which takes write_xyz from optimization_methods)
_IO_IMPORTERS = ("tscode.hypermolecule_class", "tscode.embedder_options", "tscode.pka", "tscode.concurrent_test")
_reference_read_xyz = []                                  # the original, saved by install_into(io=True)


def read_xyz(filename):
    """Installed in place of tscode.utils.read_xyz by install_into(io=True): `.xyz` files go through the native reader
    (utils.read_xyz: .atomcoords / .atomnos / .metadata['comments'], what the reference's callers use); any other
    format ccread understands, and any file the native reader rejects, is handed to the reference's own function."""
    if str(filename).lower().endswith(".xyz"):
        try:
            return _ut.read_xyz(filename)
        except (ValueError, KeyError, AssertionError):
            if not _reference_read_xyz:
                raise
    if not _reference_read_xyz:
        raise ValueError(f"{filename}: only .xyz files can be read without the reference's read_xyz")
    return _reference_read_xyz[0](filename)


_IO_PATCHES = {
    "tscode.utils": {
        "write_xyz": _ut.write_xyz,
        "read_xyz": read_xyz,
    },
}

_saved = []


def compenetration_refining(self):
    """RunEmbedding.compenetration_refining (embedder.py:1230-1268) with the per-structure loop (:1243-1248) replaced
    by one batched clash screen; log lines, masking and the final initialisations as in the reference."""
    if self.embed not in ('string', 'cyclical', 'monomolecular'):
        from tscode.utils import time_to_string
        self.log('--> Checking structures for compenetrations')
        t_start = time.perf_counter()
        structures = np.asarray(self.structures)
        if len(structures):
            mask = _nf.compenetration_check_batch(structures, self.ids, thresh=self.options.clash_thresh,
                                                  max_clashes=self.options.max_clashes).astype(bool)
        else:
            mask = np.zeros(0, dtype=bool)
        self.apply_mask(('structures', 'constrained_indices'), mask)
        t_end = time.perf_counter()
        if False in mask:
            self.log(f'Discarded {len([b for b in mask if not b])} candidates for compenetration '
                     f'({len([b for b in mask if b])} left, {time_to_string(t_end-t_start)})')
        else:
            self.log(f'All {len(mask)} structures passed the compenetration check')
        self.log()
        self.zero_candidates_check()
    self.energies = np.full(len(self.structures), 1E10)
    self.exit_status = np.zeros(len(self.structures), dtype=bool)


def fitness_refining(self, threshold=5, verbose=False):
    """RunEmbedding.fitness_refining (embedder.py:1270-1313) with the per-structure fitness_check loop (:1283-1290)
    replaced by one batched constraint-score call (optimization_methods.py:544-557: signed sum of distance - target over
    the constraints that have a target, rejected when not below `threshold`)."""
    if verbose:
        self.log(' \n--> Fitness pruning - removing inaccurate structures')
    n = len(self.structures)
    mask = np.ones(n, dtype=bool)
    if n:
        cons = np.asarray(self.constrained_indices)
        targets = [[self.get_pairing_dists_from_constrained_indices(_c) for _c in constraints]
                   for constraints in self.constrained_indices]
        if cons.ndim == 3 and cons.shape[1] > 0:
            _, err = _om.constraint_scores(np.asarray(self.structures), cons, targets)
            mask = err < threshold
        else:                                                   # no constraints: the sum is 0 (:549-557)
            mask[:] = 0 < threshold
    attr = ('structures', 'energies', 'constrained_indices', 'exit_status')
    self.apply_mask(attr, mask)
    if False in mask:
        self.log(f'Discarded {len([b for b in mask if not b])} candidates for unfitness ({len([b for b in mask if b])} left)')
    else:
        if verbose:
            self.log('All candidates meet the imposed criteria.')
    self.log()
    self.zero_candidates_check()


def write_structures(self, tag, indices=None, energies=True, relative=True, extra='', align='indices', p=True):
    """Embedder.write_structures (embedder.py:996-1043) with the per-structure write_xyz loop (:1034-1041) replaced by
    one call of the native formatter on all frames (utils.xyz_text); alignment, titles, truncation and log lines are
    the reference's own, line by line (including `rel_e -= min`, which edits self.energies in place)."""
    from tscode.hypermolecule_class import align_by_moi, align_structures
    align_functions = {
        'indices': align_structures,
        'moi': align_by_moi,
    }
    if energies:
        rel_e = self.energies
        if relative:
            rel_e -= np.min(self.energies)
    # truncate if there are too many (embed debug first dump)
    if len(self.structures) > 10000 and not self.options.let:
        self.log(f'Truncated {tag} output structures to 10000 (from {len(self.structures)} - keyword LET to override).')
        output_structures = self.structures[0:10000]
    else:
        output_structures = self.structures
    self.outname = f'tscode_{tag}_{self.stamp}.xyz'
    with open(self.outname, 'w') as f:
        aligned = align_functions[align](output_structures, atomnos=self.atomnos, indices=indices)
        titles = []
        for i in range(len(aligned)):
            title = f'Strucure {i+1} - {tag}'
            if energies:
                title += f' - Rel. E. = {round(rel_e[i], 3)} kcal/mol '
            title += extra
            titles.append(title)
        if len(aligned):
            f.write(_ut.xyz_text(np.asarray(aligned), self.atomnos, titles).decode())
    if p:
        self.log(f'Wrote {len(output_structures)} {tag} structures to {self.outname} file.\n')


_IO_METHOD_PATCHES = {
    ("tscode.embedder", "Embedder"): {
        "write_structures": write_structures,
    },
}

_METHOD_PATCHES = {
    ("tscode.embedder", "RunEmbedding"): {
        "compenetration_refining": compenetration_refining,
        "fitness_refining": fitness_refining,
    },
}


def install_into(tscode_pkg=None, strict: bool = False, scalars: bool = False, loops: bool = True, io: bool = False):
    """Rebind the reference's hot-path names to the CUDA implementations.  Returns the list of
    (module, name) pairs that were patched.  Modules that are not importable in this
    environment are skipped unless strict=True.  scalars / loops: see the module docstring.  io=True also rebinds
    utils.write_xyz / read_xyz to the native formatter and reader (read_xyz above: `.xyz` only, the
    reference's function for everything else)."""
    patched = []
    originals = {}
    tables = [_PATCHES] + ([_SCALAR_PATCHES] if scalars else []) + ([_IO_PATCHES] if io else [])
    for table in tables:
        for modname, names in table.items():
            try:
                mod = importlib.import_module(modname)
            except Exception:
                if strict:
                    raise
                continue
            for name, repl in names.items():
                if hasattr(mod, name):
                    originals[name] = getattr(mod, name)
                    if name == "prune_conformers_rmsd_rot_corr" and not _reference_rot_corr:
                        _reference_rot_corr.append(originals[name])
                    if name == "read_xyz" and not _reference_read_xyz:
                        _reference_read_xyz.append(originals[name])
                    _saved.append((mod, name, originals[name]))
                    setattr(mod, name, repl)
                    patched.append((modname, name))
    for modname in _IMPORTERS + (_IO_IMPORTERS if io else ()):
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                if strict:
                    raise
                continue
        for table in tables:
            for names in table.values():
                for name, repl in names.items():
                    cur = getattr(mod, name, None)
                    if cur is not None and cur is originals.get(name):
                        _saved.append((mod, name, cur))
                        setattr(mod, name, repl)
                        patched.append((modname, name))
    method_tables = ([_METHOD_PATCHES] if loops else []) + ([_IO_METHOD_PATCHES] if io else [])
    for method_table in method_tables:
        for (modname, clsname), methods in method_table.items():
            try:
                cls = getattr(importlib.import_module(modname), clsname)
            except Exception:
                if strict:
                    raise
                continue
            for name, repl in methods.items():
                if hasattr(cls, name):
                    _saved.append((cls, name, getattr(cls, name)))
                    setattr(cls, name, repl)
                    patched.append((f"{modname}.{clsname}", name))
    return patched


def uninstall():
    del _reference_rot_corr[:]
    del _reference_read_xyz[:]
    while _saved:
        mod, name, orig = _saved.pop()
        setattr(mod, name, orig)
