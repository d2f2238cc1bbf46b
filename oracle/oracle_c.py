"""ctypes front-end of liboracle.so (oracle.c).  TEST INFRASTRUCTURE — see oracle/__init__.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_rmsd_and_max.argtypes = [_dp, _dp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_kabsch_rotation.argtypes = [_dp, _dp, C.c_int, _dp]
        L.orc_sim_rows.argtypes = [_dp, C.c_long, C.c_int, C.c_double, C.c_long, C.c_long, _bp, C.c_void_p, C.c_void_p]
        L.orc_eval_pairs.argtypes = [_dp, C.c_int, C.c_double, _lp, _lp, C.c_long]
        L.orc_eval_pairs.restype = C.c_long
        L.orc_prune_rmsd.argtypes = [_dp, C.c_long, C.c_int, C.c_double, C.c_void_p, _bp,
                                     C.POINTER(C.c_longlong), C.c_void_p]
        L.orc_prune_rmsd.restype = C.c_long
        L.orc_rmsd_similarity.argtypes = [_dp, _dp, C.c_long, C.c_int, C.c_double]
        L.orc_rmsd_similarity.restype = C.c_int
        L.orc_compenetration_check.argtypes = [_dp, C.c_long, C.c_void_p, C.c_int, C.c_double, C.c_long]
        L.orc_compenetration_check.restype = C.c_int
        L.orc_get_embed.argtypes = [_dp, _lp, _ip, C.c_int, _lp, _dp, _dp, _dp]
        L.orc_embed_clash_batch.argtypes = [_dp, _lp, _ip, C.c_int, _lp, _dp, _dp, C.c_long, C.c_double,
                                            C.c_long, _bp]
        L.orc_clash_structs.argtypes = [_dp, C.c_long, C.c_long, C.c_void_p, C.c_int, C.c_double, C.c_long, _bp]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_threads(n: int):
    lib().orc_set_threads(int(n))


def rmsd_and_max(p, q):
    p, q = _c(p), _c(q)
    r, d = C.c_double(), C.c_double()
    lib().orc_rmsd_and_max(p, q, p.shape[0], C.byref(r), C.byref(d))
    return r.value, d.value


def kabsch_rotation(p, q):
    p, q = _c(p), _c(q)
    R = np.empty((3, 3))
    lib().orc_kabsch_rotation(p, q, p.shape[0], R)
    return R


def heavy(structures, atomnos):
    """rmsd_pruning.py:178-179."""
    structures = np.asarray(structures, dtype=np.float64)
    return np.ascontiguousarray(structures[:, np.asarray(atomnos) != 1])


def sim_rows(H, thr, row_begin, row_end, want_values=False):
    H = _c(H)
    N, M = H.shape[0], H.shape[1]
    n = row_end - row_begin
    sim = np.zeros((n, N), np.uint8)
    if want_values:
        r = np.zeros((n, N)); d = np.zeros((n, N))
        lib().orc_sim_rows(H, N, M, thr, row_begin, row_end, sim, r.ctypes.data, d.ctypes.data)
        return sim, r, d
    lib().orc_sim_rows(H, N, M, thr, row_begin, row_end, sim, None, None)
    return sim


def eval_pairs(H, thr, ii, jj):
    H = _c(H)
    ii, jj = _c(ii, np.int64), _c(jj, np.int64)
    return lib().orc_eval_pairs(H, H.shape[1], thr, ii, jj, ii.shape[0])


def prune_heavy(H, thr, sim_bytes=None):
    """Ladder prune on the heavy-atom array.  Returns (mask, n_eval, rounds)."""
    H = _c(H)
    N, M = H.shape[0], H.shape[1]
    mask = np.zeros(max(N, 1), np.uint8)
    ne = C.c_longlong(0)
    rounds = np.zeros(19)
    sb = None
    if sim_bytes is not None:
        sim_bytes = _c(sim_bytes, np.uint8)
        assert sim_bytes.shape == (N, N)
        sb = sim_bytes.ctypes.data
    lib().orc_prune_rmsd(H, N, M, thr, sb, mask, C.byref(ne), rounds.ctypes.data)
    return mask[:N].astype(bool), ne.value, [k for k in rounds if k > 0]


def prune_conformers_rmsd(structures, atomnos, rmsd_thr=0.5):
    """rmsd_pruning.py:164-206 — same signature and return."""
    structures = np.asarray(structures)
    mask, _, _ = prune_heavy(heavy(structures, atomnos), rmsd_thr)
    return structures[mask], mask


def rmsd_similarity(ref, structures, rmsd_thr=0.5):
    S = _c(np.asarray(structures).reshape(-1, np.asarray(ref).shape[0], 3))
    if S.shape[0] == 0:
        return False
    return bool(lib().orc_rmsd_similarity(_c(ref), S, S.shape[0], S.shape[1], rmsd_thr))


def compenetration_check(coords, ids=None, thresh=1.5, max_clashes=0) -> int:
    coords = _c(coords)
    if ids is None:
        return int(lib().orc_compenetration_check(coords, coords.shape[0], None, 0, thresh, max_clashes))
    ids = _c(ids, np.int64)
    return int(lib().orc_compenetration_check(coords, coords.shape[0], ids.ctypes.data, len(ids), thresh, max_clashes))


def pack_frags(frags):
    """Flatten a fragment library [(n_conf, n_k, 3), ...] -> (lib, off, n_atoms)."""
    off, n_atoms, parts, o = [], [], [], 0
    for f in frags:
        f = _c(f)
        off.append(o); n_atoms.append(f.shape[1]); parts.append(f.ravel()); o += f.size
    return np.concatenate(parts), np.array(off, np.int64), np.array(n_atoms, np.int32)


def get_embed(frags, conf_ids, R, t):
    flib, off, na = pack_frags(frags)
    out = np.empty((int(na.sum()), 3))
    lib().orc_get_embed(flib, off, na, len(na), _c(conf_ids, np.int64), _c(R), _c(t), out)
    return out


def embed_clash_batch(frags, conf, R, t, thresh=1.5, max_clashes=0):
    flib, off, na = pack_frags(frags)
    conf = _c(conf, np.int64)
    P = conf.shape[0]
    v = np.zeros(max(P, 1), np.uint8)
    lib().orc_embed_clash_batch(flib, off, na, len(na), conf, _c(R), _c(t), P, thresh, max_clashes, v)
    return v[:P]


def clash_structs(S, ids=None, thresh=1.5, max_clashes=0):
    S = _c(S)
    P, A = S.shape[0], S.shape[1]
    v = np.zeros(max(P, 1), np.uint8)
    if ids is None:
        lib().orc_clash_structs(S, P, A, None, 0, thresh, max_clashes, v)
    else:
        ids = _c(ids, np.int64)
        lib().orc_clash_structs(S, P, A, ids.ctypes.data, len(ids), thresh, max_clashes, v)
    return v[:P]
