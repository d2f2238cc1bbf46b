"""Drop-in for the ensemble-output helper of tscode/utils.py:

    write_xyz(coords, atomnos, output, title='temp')                                   # utils.py:114-126
    read_xyz(filename) -> object with .atomcoords (n_frames, A, 3), .atomnos (A,)         # utils.py:128-135

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


class XyzEnsemble:
    """What the reference's callers use of the ccData object read_xyz returns: `.atomcoords` (n_frames, A, 3) float64,
    `.atomnos` (A,) ints (hypermolecule_class.py:163-168, operators.py:109, 169, 285), and cclib's
    `.metadata['comments']` (the frames' title lines)."""

    def __init__(self, atomcoords, atomnos, comments):
        self.atomcoords, self.atomnos, self.metadata = atomcoords, atomnos, {"comments": comments}
        self.natom = int(atomnos.shape[0])


_XYZ_ERRORS = {-1: "bad arguments", -2: "malformed frame (atom count line, or an atom line with fewer than 4 columns)",
               -3: "frames with different numbers of atoms", -4: "a coordinate is not a number", -5: "buffer too small"}


def parse_xyz(text, n_threads=None) -> XyzEnsemble:
    """Multi-frame XYZ text (bytes or str) -> XyzEnsemble, by the library's host parser (tsc_host_read_xyz: the
    algorithm of cclib's XYZ reader — optional blank line, atom count, comment line, `count` lines "symbol x y z
    [ignored]", an incomplete last frame dropped, symbols of the last frame — with numbers converted like float())."""
    data = text.encode() if isinstance(text, str) else bytes(text)
    L = lib()
    nt = int(n_threads) if n_threads else min(os.cpu_count() or 1, 16)
    n_at = ctypes.c_int32(0)
    n = int(L.tsc_host_read_xyz(data, len(data), ctypes.byref(n_at), None, 0, None, None, 1))
    if n < 0:
        raise ValueError("XYZ text: " + _XYZ_ERRORS.get(n, str(n)))
    A = int(n_at.value)
    coords = np.empty((n, A, 3), dtype=np.float64)
    sym = np.zeros((max(A, 1), 4), dtype=np.uint8)
    spans = np.zeros((max(n, 1), 2), dtype=np.int64)
    if n:
        rc = int(L.tsc_host_read_xyz(data, len(data), ctypes.byref(n_at), coords.ctypes.data, n, sym.ctypes.data,
                                     spans.ctypes.data, nt))
        if rc != n:
            raise ValueError("XYZ text: " + _XYZ_ERRORS.get(rc, str(rc)))
    number = {s: z for z, s in enumerate(_SYMBOLS)}
    try:                                   # the drop-in scenario: the reference's own table
        from tscode.pt import pt
        number.update({pt[z].symbol: z for z in range(1, 119)})
    except Exception:
        pass
    symbols = [bytes(r).rstrip(b"\0").decode() for r in sym[:A]] if n else []
    try:
        atomnos = np.array([number[s] for s in symbols], dtype=int)
    except KeyError as exc:
        raise KeyError(f"XYZ text: unknown element symbol {exc.args[0]!r}") from None
    comments = [data[o:o + ln].decode(errors="replace") for o, ln in spans[:n].tolist()]
    return XyzEnsemble(coords, atomnos, comments)


def read_xyz(filename) -> XyzEnsemble:
    """Drop-in for tscode.utils.read_xyz (utils.py:128-135) for .xyz files: the fields the reference's callers use."""
    with open(filename, "rb") as f:
        mol = parse_xyz(f.read())
    assert mol.atomcoords.shape[0] > 0, f'Reading molecule {filename} failed - check its integrity.'
    return mol
