"""CPU checks of the per-pair device math (tsc_math.cuh compiled for the host) vs the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_c
from tscode_b200.synth import gen_ensemble

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(HERE, "hostmath", "hostmath.cpp")
    so = os.path.join(HERE, "hostmath", "libhostmath.so")
    hdr = os.path.join(HERE, "..", "tscode_b200", "csrc", "tsc_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    L.hm_screen.argtypes = [_dp, _dp, C.c_int, C.c_double]
    L.hm_screen_nomargin.argtypes = [_dp, _dp, C.c_int, C.c_double]
    L.hm_rmsd_and_max.argtypes = [_dp, _dp, C.c_int] + [C.POINTER(C.c_double)] * 4
    return L


def _rm(L, p, q):
    r, d, lam, gap = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    L.hm_rmsd_and_max(np.ascontiguousarray(p), np.ascontiguousarray(q), len(p), C.byref(r), C.byref(d), C.byref(lam), C.byref(gap))
    return r.value, d.value, lam.value, gap.value


def test_verify_math_matches_oracle(hm):
    worst = 0.0
    for M, noise in ((4, 0.3), (17, 0.05), (40, 0.2), (80, 1.0), (80, 1e-4)):
        S = gen_ensemble(7 + M, 60, M, 5, sigma_noise=noise)
        for i in range(0, 60, 2):
            r0, d0 = oracle_c.rmsd_and_max(S[i], S[i + 1])
            r1, d1, lam, gap = _rm(hm, S[i], S[i + 1])
            worst = max(worst, abs(r0 - r1), abs(d0 - d1))
            G = (S[i] ** 2).sum() + (S[i + 1] ** 2).sum()
            # closed form E = G - 2 lambda agrees with explicit differences
            assert abs(np.sqrt(max(G - 2 * lam, 0) / M) - r0) < 1e-6
    assert worst < 1e-11, worst


def test_verify_math_reflection_case(hm):
    """det(cov) < 0: the reference flips the last singular vector (rmsd_pruning.py:20-23)."""
    rng = np.random.default_rng(5)
    n = 0
    for _ in range(200):
        p = rng.normal(size=(6, 3)); q = rng.normal(size=(6, 3))
        if np.linalg.det(p.T @ q) < 0:
            n += 1
            r0, d0 = oracle_c.rmsd_and_max(p, q)
            r1, d1, _, gap = _rm(hm, p, q)
            if gap > 1e-6:
                assert abs(r0 - r1) < 1e-11 and abs(d0 - d1) < 1e-9
    assert n > 50


def test_screen_never_loses_a_similar_pair(hm):
    """Screen (Budan-Fourier sign test at lam_t) must keep every pair with rmsd < thr, and should
    reject almost everything else."""
    kept = rej = 0
    for seed, M, noise, thr in ((1, 40, 0.05, 0.5), (2, 80, 0.3, 0.5), (3, 12, 0.2, 0.25), (4, 33, 0.29, 0.5)):
        S = gen_ensemble(seed, 120, M, 6, sigma_noise=noise)
        for i in range(0, 120):
            for j in range(i + 1, min(i + 20, 120)):
                r, _ = oracle_c.rmsd_and_max(S[i], S[j])
                c = hm.hm_screen(S[i], S[j], M, thr)
                c0 = hm.hm_screen_nomargin(S[i], S[j], M, thr)
                if r < thr:
                    assert c == 1, (seed, i, j, r)
                if abs(r - thr) > 1e-7:
                    assert c0 == int(r < thr), (seed, i, j, r, c0)   # exact equivalence away from thr
                    assert c == int(r < thr) or abs(r - thr) < 1e-5
                kept += c; rej += 1 - c
    assert kept > 100 and rej > 100


def test_screen_degenerate_inputs(hm):
    z = np.zeros((5, 3))
    assert hm.hm_screen(z, z, 5, 0.5) == 1                      # lam_t <= 0 -> candidate
    p = np.zeros((5, 3)); p[:, 0] = np.arange(5) * 3.0          # collinear
    q = p.copy(); q[:, 0] += 0.01
    assert hm.hm_screen(p, q, 5, 0.5) == 1
    q2 = p.copy(); q2[:, 0] *= 3
    assert hm.hm_screen(p, q2, 5, 0.5) == 0


def test_screen_far_from_origin_never_loses_a_pair(hm):
    """Ensembles far from the origin (rotation-only Kabsch does not centre, rmsd_pruning.py:6-41) make two roots of
    the key-matrix quartic nearly coincide at ~1e11: without the rounding guard of quartic_excluded the sign test
    dropped true pairs."""
    S = gen_ensemble(11, 120, 24, 8, sigma_noise=0.05)
    lost = sim = 0
    for shift in (7e4, 3e3, 1e6):
        T = S.copy(); T[:, :, 0] += shift
        for i in range(0, 120):
            for j in range(i + 1, min(i + 25, 120)):
                r, d = oracle_c.rmsd_and_max(T[i], T[j])
                if r < 0.5:
                    sim += 1
                    lost += 1 - hm.hm_screen(T[i], T[j], 24, 0.5)
    assert sim > 100 and lost == 0, (sim, lost)
