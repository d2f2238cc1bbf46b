"""Drop-in for the ensemble-output helper of tscode/utils.py:

    write_xyz(coords, atomnos, output, title='temp')                                   # utils.py:114-126

plus the batched form `xyz_text(structures, atomnos, titles)` that Embedder.write_structures-style loops
(embedder.py:996-1043: one write_xyz call per structure) should use: the text of all frames is produced by the
native formatter (tsc_host_write_xyz, a few host threads) byte-identically to the reference's
'%s     % .6f % .6f % .6f\\n' per atom.  Host-only: no GPU involved.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from ._lib import lib

_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
            "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu "
            "Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()


def _symbols_for(atomnos):
    try:                                   # the drop-in scenario: the reference's own table
        from tscode.pt import pt
        return [pt[int(a)].symbol for a in atomnos]
    except Exception:
        return [_SYMBOLS[int(a)] for a in atomnos]


def xyz_text(structures, atomnos, titles=None, n_threads=None) -> bytes:
    """Multi-frame XYZ text of `structures` (n_frames, A, 3): what calling write_xyz once per structure into the
    same file produces.  titles: one string per frame (default 'temp')."""
    S = np.ascontiguousarray(structures, dtype=np.float64)
    if S.ndim == 2:
        S = S[None]
    n, A = S.shape[0], S.shape[1]
    atomnos = np.asarray(atomnos)
    assert atomnos.shape[0] == A and S.shape[2] == 3
    sym = b"".join(s.encode().ljust(4, b"\0") for s in _symbols_for(atomnos))
    tt = None
    if titles is not None:
        titles = [titles] * n if isinstance(titles, str) else list(titles)
        assert len(titles) == n
        tt = b"".join(str(t).encode() + b"\0" for t in titles)
    nt = int(n_threads) if n_threads else min(os.cpu_count() or 1, 16)
    L = lib()
    cap = n * (A * 48 + 64) + (len(tt) if tt else 8 * n) + 64        # typical size; the call reports -needed if short
    while True:
        buf = np.empty(max(cap, 1), dtype=np.uint8)
        got = int(L.tsc_host_write_xyz(S.ctypes.data, n, A, sym, tt, buf.ctypes.data, cap, nt))
        if got >= 0:
            return buf[:got].tobytes()
        cap = -got


def write_xyz(coords, atomnos, output, title='temp'):
    """Drop-in for tscode.utils.write_xyz (utils.py:114-126): `output` is a text file object."""
    coords = np.asarray(coords)
    assert np.asarray(atomnos).shape[0] == coords.shape[0]
    assert coords.shape[1] == 3
    output.write(xyz_text(coords[None], atomnos, [title], n_threads=1).decode())
