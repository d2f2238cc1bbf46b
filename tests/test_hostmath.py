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
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(HERE, "hostmath", "hostmath.cpp")
    so = os.path.join(HERE, "hostmath", "libhostmath.so")
    hdr = os.path.join(HERE, "..", "tscode_b200", "csrc", "tsc_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    L.hm_screen.argtypes = [_dp, _dp, C.c_int, C.c_double]
    L.hm_screen_nomargin.argtypes = [_dp, _dp, C.c_int, C.c_double]
    L.hm_rmsd_and_max.argtypes = [_dp, _dp, C.c_int] + [C.POINTER(C.c_double)] * 4
    L.hm_q32_errors.argtypes = [_fp, _fp, C.c_int, _dp]
    L.hm_q32_excluded.argtypes = [_fp, C.c_float, C.POINTER(C.c_double)]
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


def _q32_covariances(rng, kind, n):
    if kind == 0:        # anything, 8 decades of magnitude
        S = rng.normal(size=(n, 9)) * 10 ** rng.uniform(-3, 5, size=(n, 1))
    elif kind == 1:      # elongated molecule, similar orientation: near-PSD anisotropic covariance
        P = rng.normal(size=(n, 20, 3)) * np.array([6, 2, 1.0]); Q = P + rng.normal(size=(n, 20, 3)) * 0.3
        S = np.einsum("nma,nmb->nab", P, Q).reshape(n, 9)
    elif kind == 2:      # rank one
        a = rng.normal(size=(n, 3)); b = rng.normal(size=(n, 3))
        S = (a[:, :, None] * b[:, None, :]).reshape(n, 9) * 1e3
    elif kind == 3:      # ensemble far from the origin: two roots of the quartic nearly coincide
        P = rng.normal(size=(n, 10, 3)) + np.array([7e3, 0, 0]); Q = P + rng.normal(size=(n, 10, 3)) * 0.05
        S = np.einsum("nma,nmb->nab", P, Q).reshape(n, 9)
    else:                # many exact zeros
        S = rng.normal(size=(n, 9)) * (rng.random((n, 9)) < 0.4) * 1e2
    return np.ascontiguousarray(S, dtype=np.float32)


def test_fp32_quartic_forward_errors_within_tolerances(hm):
    """quartic32_values (FP32 second stage of the tcgen05 pre-screens, tsc_math.cuh): measured forward errors of
    P, P', P'' against long-double evaluation stay well inside the tolerances quartic32_margins applies
    (64 u R^2, 72 u R^1.5, 68 u R with R = lam^2 + 4 f >= rho^2)."""
    rng = np.random.default_rng(0)
    u = 2.0 ** -24
    worst = np.zeros(3)
    for trial in range(10):
        n = 20000
        S = _q32_covariances(rng, trial % 5, n)
        s = np.sqrt((S.astype(np.float64) ** 2).sum(1))
        lam = (s * 10 ** rng.uniform(-2, 1, size=n)).astype(np.float32)
        out = np.zeros(3 * n)
        hm.hm_q32_errors(S, lam, n, out)
        o = out.reshape(n, 3)
        o = o[np.isfinite(o).all(1)]
        assert len(o) > 0.9 * n
        worst = np.maximum(worst, o.max(axis=0))
    assert worst[0] < 64 * u / 8 and worst[1] < 72 * u / 3 and worst[2] < 68 * u / 2, worst / u


def test_fp32_quartic_never_excludes_a_pair_above_the_test_point(hm):
    """Whenever the FP32 stage excludes, lambda_max (FP64 eigen-solve of the same covariance) is below the test
    point; and it does exclude the anisotropic pairs Samuelson's bound cannot (tools/aniso_probe.py)."""
    rng = np.random.default_rng(1)
    n = 3000
    base = rng.normal(size=(80, 3)) * np.array([6, 2, 1.0])
    P = base + rng.normal(size=(n, 80, 3)); Q = base + rng.normal(size=(n, 80, 3))
    S = np.ascontiguousarray(np.einsum("nma,nmb->nab", P, Q).reshape(n, 9), dtype=np.float32)
    Gp, Gq = (P ** 2).sum((1, 2)), (Q ** 2).sum((1, 2))
    lam_t = (0.5 * (Gp + Gq - 80 * 0.25) - 1.05e-3 * np.sqrt(3.0) * np.sqrt(Gp * Gq)).astype(np.float32)
    lm = C.c_double()
    excluded = samuelson = 0
    for k in range(n):
        e = hm.hm_q32_excluded(S[k], float(lam_t[k]), C.byref(lm))
        assert not (e and lm.value > lam_t[k])
        excluded += e
        samuelson += bool(3.00004 * float((S[k].astype(np.float64) ** 2).sum()) <= float(lam_t[k]) ** 2)
    assert excluded == n and samuelson == 0
    # test points swept through lambda_max: the decision flips from "not excluded" to "excluded" within a relative
    # band of 1e-3 above lambda_max, never below it
    for k in range(0, n, 10):
        hm.hm_q32_excluded(S[k], 1.0, C.byref(lm))
        for rel in (-1e-2, -1e-4, -1e-6, 0.0, 1e-6):
            assert hm.hm_q32_excluded(S[k], float(np.float32(lm.value * (1 + rel))) , C.byref(lm)) == 0 or rel > 0
        assert hm.hm_q32_excluded(S[k], float(np.float32(lm.value * (1 + 1e-3))), C.byref(lm)) == 1
    # NaN / inf / zero inputs are never excluded
    bad = np.zeros(9, np.float32)
    assert hm.hm_q32_excluded(bad, 0.0, C.byref(lm)) == 0
    bad[0] = np.inf
    assert hm.hm_q32_excluded(bad, 10.0, C.byref(lm)) == 0
    bad[0] = np.nan
    assert hm.hm_q32_excluded(bad, 10.0, C.byref(lm)) == 0


def _rd32(x):
    y = np.float32(x)
    return np.nextafter(y, np.float32(-np.inf)) if float(y) > x else y


def _ru32(x):
    y = np.float32(x)
    return np.nextafter(y, np.float32(np.inf)) if float(y) < x else y


@pytest.mark.parametrize("case", [
    dict(seed=0, N=260, M=40, ncl=26, noise=0.05, thr=0.5, scale=3.0),
    dict(seed=13, N=240, M=80, ncl=12, noise=0.2, thr=0.5, scale=3.0),                     # many pairs near the threshold
    dict(seed=31, N=260, M=40, ncl=26, noise=0.05, thr=0.5, scale=[6.0, 2.0, 1.0]),        # elongated
    dict(seed=32, N=220, M=80, ncl=16, noise=0.08, thr=0.5, scale=[4.0, 4.0, 0.5]),        # planar
    dict(seed=33, N=260, M=17, ncl=20, noise=0.05, thr=0.3, scale=[8.0, 1.0, 1.0]),        # rod
], ids=["iso", "near_thr", "elongated", "planar", "rod"])
def test_f16_prescreen_model_never_loses_a_similar_pair(hm, case):
    """Host model of the default pre-screen: coordinates rounded to FP16 as tsc_pack_f16 does,
    covariance accumulated in FP32, threshold eigenvalue lowered by the operand error bound with the directed
    roundings of screen_row_consts / the column terms, stage 1 (Samuelson) and stage 2 (FP32 quartic through the same
    tsc_math.cuh code the device runs).  Every pair the oracle calls similar must survive both stages; almost
    everything else must be excluded."""
    S = gen_ensemble(case["seed"], case["N"], case["M"], case["ncl"], sigma_noise=case["noise"],
                     scale=np.array(case["scale"]) if isinstance(case["scale"], list) else case["scale"])
    N, M, thr = case["N"], case["M"], case["thr"]
    sim = oracle_c.sim_rows(S, thr, 0, N).astype(bool)
    X = S.astype(np.float16).astype(np.float32)
    G = (S ** 2).sum((1, 2))
    sG = np.sqrt(G)
    hs, cc, e_thr = 0.5 * (1.0 - 1e-10), np.sqrt(3.0) * 1.05e-3, M * thr * thr * (1.0 + 1e-6)
    Af = [_rd32(hs * G[i] - 0.5 * e_thr) for i in range(N)]
    Cf = [_ru32(cc * sG[i]) for i in range(N)]
    Bf = [_rd32(hs * G[j]) for j in range(N)]
    Df = [_ru32(sG[j]) for j in range(N)]
    lm = C.c_double()
    lost = excluded = dissimilar = stage2 = 0
    for i in range(N):
        cov = np.einsum("ma,jmb->jab", X[i], X[i + 1:]).reshape(-1, 9).astype(np.float32)
        for k, j in enumerate(range(i + 1, N)):
            ab = np.float32(Af[i] + Bf[j])
            lf = np.float32(np.float64(ab) - np.float64(Cf[i]) * np.float64(Df[j]))        # one rounding, as the FMA
            f = np.float32(0)
            for q in range(9):
                f = np.float32(np.float64(cov[k, q]) * np.float64(cov[k, q]) + np.float64(f))
            t = np.float32(3.00004) * f - lf * lf
            out = bool(lf > 0 and t < 0)
            if not out:
                stage2 += 1
                lam = np.float32(np.float64(lf) - 2e-7 * (abs(float(ab)) + abs(float(lf))))
                out = bool(hm.hm_q32_excluded(np.ascontiguousarray(cov[k]), float(lam), C.byref(lm)))
            if sim[i, j]:
                lost += out
            else:
                dissimilar += 1
                excluded += out
    assert lost == 0
    assert excluded > 0.9 * dissimilar or case["noise"] >= 0.2, (excluded, dissimilar)
    print(case, "stage 2 ran for", stage2, "of", N * (N - 1) // 2, "pairs; excluded", excluded, "of", dissimilar)


# ---- T = S^T S form of the FP32 stage (component-sequential screen, rmsd_screen.cu) ------------------------------------
@pytest.fixture(scope="module")
def hmt(hm):
    hm.hm_q32t_excluded.argtypes = [_fp, C.c_float, C.POINTER(C.c_double)]
    hm.hm_q32t_det.argtypes = [_fp, C.c_int, _dp]
    hm.hm_q32t_excluded_scaled.argtypes = [_fp, _dp, C.c_float, C.POINTER(C.c_double)]
    return hm


def test_fp32_T_form_det_bound_is_an_upper_bound(hmt):
    """d = sqrt(max(det T~, 0) + 8 u f^3) must never be below |det S| (long double), for covariances of every kind —
    it replaces the signed determinant in the quartic's linear coefficient (tsc_math.cuh)."""
    rng = np.random.default_rng(3)
    u = 2.0 ** -24
    for trial in range(10):
        n = 20000
        S = _q32_covariances(rng, trial % 5, n)
        out = np.zeros(3 * n)
        hmt.hm_q32t_det(S, n, out)
        o = out.reshape(n, 3)
        ok = np.isfinite(o).all(1)
        d, true, f = o[ok, 0], o[ok, 1], o[ok, 2]
        # (kind 3, an ensemble 7e3 A from the origin, overflows f^3 in FP32: the bound is +inf, nothing gets excluded)
        assert ok.sum() > 0.9 * n or trial % 5 == 3
        assert (d * (1 + 4 * u) >= true).all(), float((true - d).max())
        # and it is not wildly loose: within the margin's own root of |det S|
        assert (d <= true + 1.01 * np.sqrt(8 * u * f ** 3) + 1e-30).all()


def test_fp32_T_form_never_excludes_a_pair_above_the_test_point(hmt):
    """As test_fp32_quartic_never_excludes_a_pair_above_the_test_point for the T form: sound on every kind of
    covariance (incl. det S < 0, where it is weaker by design), and still excluding the anisotropic pairs."""
    rng = np.random.default_rng(4)
    lm = C.c_double()
    # soundness on generic covariances with test points around lambda_max
    for kind in range(5):
        S = _q32_covariances(rng, kind, 400)
        for k in range(400):
            hmt.hm_q32t_excluded(S[k], 1.0, C.byref(lm))
            top = lm.value
            if not np.isfinite(top) or top <= 0:
                continue
            for rel in (-0.5, -1e-2, -1e-4, -1e-6, 0.0):
                assert hmt.hm_q32t_excluded(S[k], float(np.float32(top * (1 + rel))), C.byref(lm)) == 0
    # tiny covariances: the device's flush-to-zero root returns d = 0 once det T + 8 u f^3 is subnormal; the test point
    # then has to exceed Q32_LAM_MIN = 1e-4 (tsc_math.cuh, quartic32_decide) — still sound, and nothing below the cut
    # is ever excluded
    for scale in (1e-2, 1e-4, 1e-5, 1e-6, 3e-7):
        for kind in (0, 1, 2, 4):
            S = np.ascontiguousarray(_q32_covariances(rng, kind, 200) * np.float32(scale))
            for k in range(200):
                hmt.hm_q32t_excluded(S[k], 1.0, C.byref(lm))
                top = lm.value
                if not np.isfinite(top) or top <= 0:
                    continue
                for rel in (-0.5, -1e-3, -1e-6, 0.0):
                    assert hmt.hm_q32t_excluded(S[k], float(np.float32(top * (1 + rel))), C.byref(lm)) == 0
                for lam in (1e-5, 9.9e-5, 1e-4):
                    assert hmt.hm_q32t_excluded(S[k], float(np.float32(lam)), C.byref(lm)) == 0
    # screening power on the elongated sample of the signed form's test
    n = 3000
    base = rng.normal(size=(80, 3)) * np.array([6, 2, 1.0])
    P = base + rng.normal(size=(n, 80, 3)); Q = base + rng.normal(size=(n, 80, 3))
    S = np.ascontiguousarray(np.einsum("nma,nmb->nab", P, Q).reshape(n, 9), dtype=np.float32)
    Gp, Gq = (P ** 2).sum((1, 2)), (Q ** 2).sum((1, 2))
    lam_t = (0.5 * (Gp + Gq - 80 * 0.25) - 1.05e-3 * np.sqrt(3.0) * np.sqrt(Gp * Gq)).astype(np.float32)
    excluded = 0
    for k in range(n):
        e = hmt.hm_q32t_excluded(S[k], float(lam_t[k]), C.byref(lm))
        assert not (e and lm.value > lam_t[k])
        excluded += e
    assert excluded == n
    for k in range(0, n, 10):     # det S > 0 here: the decision flips within 2e-3 (relative) above lambda_max
        hmt.hm_q32t_excluded(S[k], 1.0, C.byref(lm))
        assert hmt.hm_q32t_excluded(S[k], float(np.float32(lm.value * (1 + 2e-3))), C.byref(lm)) == 1
    bad = np.zeros(9, np.float32)
    assert hmt.hm_q32t_excluded(bad, 0.0, C.byref(lm)) == 0
    bad[0] = np.inf
    assert hmt.hm_q32t_excluded(bad, 10.0, C.byref(lm)) == 0
    bad[0] = np.nan
    assert hmt.hm_q32t_excluded(bad, 10.0, C.byref(lm)) == 0


@pytest.mark.parametrize("case", [
    dict(seed=0, N=200, M=40, ncl=20, noise=0.05, thr=0.5, scale=3.0),
    dict(seed=13, N=200, M=80, ncl=10, noise=0.2, thr=0.5, scale=3.0),
    dict(seed=31, N=200, M=40, ncl=20, noise=0.05, thr=0.5, scale=[6.0, 2.0, 1.0]),
    dict(seed=32, N=200, M=80, ncl=16, noise=0.08, thr=0.5, scale=[4.0, 4.0, 0.5]),
    dict(seed=33, N=200, M=17, ncl=20, noise=0.05, thr=0.3, scale=[8.0, 1.0, 1.0]),
    dict(seed=34, N=200, M=30, ncl=20, noise=0.05, thr=0.5, scale=[5.0, 5.0, 0.02]),       # flat (aromatic-like)
], ids=["iso", "near_thr", "elongated", "planar", "rod", "flat"])
def test_screen_model_T_form_never_loses_a_similar_pair(hmt, case):
    """Host model of the component-sequential screen: FP16 operands, FP32 covariance rows, T accumulated row by
    row, f = tr T, Samuelson then the T-form quartic.  No similar pair may be lost; most others are excluded."""
    S = gen_ensemble(case["seed"], case["N"], case["M"], case["ncl"], sigma_noise=case["noise"],
                     scale=np.array(case["scale"]) if isinstance(case["scale"], list) else case["scale"])
    N, M, thr = case["N"], case["M"], case["thr"]
    sim = oracle_c.sim_rows(S, thr, 0, N).astype(bool)
    X = S.astype(np.float16).astype(np.float32)
    G = (S ** 2).sum((1, 2))
    sG = np.sqrt(G)
    hs, cc, e_thr = 0.5 * (1.0 - 1e-10), np.sqrt(3.0) * 1.05e-3, M * thr * thr * (1.0 + 1e-6)
    Af = [_rd32(hs * G[i] - 0.5 * e_thr) for i in range(N)]
    Cf = [_ru32(cc * sG[i]) for i in range(N)]
    Bf = [_rd32(hs * G[j]) for j in range(N)]
    Df = [_ru32(sG[j]) for j in range(N)]
    lm = C.c_double()
    lost = excluded = dissimilar = stage2 = 0
    for i in range(N):
        cov = np.einsum("ma,jmb->jab", X[i], X[i + 1:]).reshape(-1, 9).astype(np.float32)
        for k, j in enumerate(range(i + 1, N)):
            ab = np.float32(Af[i] + Bf[j])
            lf = np.float32(np.float64(ab) - np.float64(Cf[i]) * np.float64(Df[j]))
            c = cov[k].astype(np.float64)
            f = np.float32(np.float32(np.float32((c[0::3] ** 2).sum()) + np.float32((c[1::3] ** 2).sum())) + np.float32((c[2::3] ** 2).sum()))
            t = np.float32(3.00004) * f - lf * lf
            out = bool(lf > 0 and t < 0)
            if not out:
                stage2 += 1
                lam = np.float32(np.float64(lf) - 2e-7 * (abs(float(ab)) + abs(float(lf))))
                out = bool(hmt.hm_q32t_excluded(np.ascontiguousarray(cov[k]), float(lam), C.byref(lm)))
            if sim[i, j]:
                lost += out
            else:
                dissimilar += 1
                excluded += out
    assert lost == 0
    assert excluded > 0.9 * dissimilar or case["noise"] >= 0.2, (excluded, dissimilar)
    print(case, "stage 2 ran for", stage2, "of", N * (N - 1) // 2, "pairs; excluded", excluded, "of", dissimilar)


def test_fp32_T_form_with_scaled_columns_stays_sound(hmt):
    """ScFrame: the column-side operand is scaled per axis and the quartic stage un-scales T^ = diag(t) T diag(t) with
    float constants before it runs (one more rounding per entry: 5 u instead of 3 u, inside the doubled tolerances,
    tsc_math.cuh).  For weights as _host.screen_frame produces them (elongated, planar, rod, extreme) the stage must
    never exclude a pair whose lambda_max reaches the test point, and must still exclude clearly separated ones."""
    rng = np.random.default_rng(11)
    lm = C.c_double()
    shapes = ([6.0, 2.0, 1.0], [4.0, 4.0, 0.5], [8.0, 1.0, 1.0], [1.0, 1.0, 1.0], [30.0, 1.0, 0.3])
    excluded = total = 0
    for sc in shapes:
        lam3 = np.maximum(np.array(sc) ** 2, 1e-4 * float(np.sum(np.array(sc) ** 2)))
        w = lam3.sum() / lam3 * (1 + 1e-9)
        t3 = np.ascontiguousarray(np.sqrt(w / 3.0))
        assert (1.0 / w).sum() <= 1.0
        base = rng.normal(size=(60, 3)) * np.array(sc)
        for k in range(400):
            P = base + rng.normal(size=base.shape) * rng.choice([0.05, 0.5, 2.0])
            Q = base + rng.normal(size=base.shape) * rng.choice([0.05, 0.5, 2.0])
            S = np.ascontiguousarray((P.T @ Q).reshape(9), dtype=np.float32)
            hmt.hm_q32t_excluded_scaled(S, t3, 1.0, C.byref(lm))
            top = lm.value
            if not np.isfinite(top) or top <= 0:
                continue
            for rel in (-0.3, -1e-2, -1e-4, -1e-6, 0.0):
                assert hmt.hm_q32t_excluded_scaled(S, t3, float(np.float32(top * (1 + rel))), C.byref(lm)) == 0
            total += 1
            excluded += hmt.hm_q32t_excluded_scaled(S, t3, float(np.float32(top * 1.01)), C.byref(lm))
    # (screening power: not sharp for the heavily perturbed pairs, whose det S can be negative — weaker by design)
    assert excluded > 0.5 * total, (excluded, total)


def _f32_dir(x, up):
    """float32 rounding of a float64 towards +inf (up) or -inf."""
    y = np.float32(x)
    if up and float(y) < x:
        y = np.nextafter(y, np.float32(np.inf))
    if (not up) and float(y) > x:
        y = np.nextafter(y, np.float32(-np.inf))
    return y


def test_weighted_samuelson_stage_with_fp16_operands_never_excludes_a_similar_pair():
    """numpy model of the whole first stage as rmsd_screen.cu runs it with a frame (ScFrame): rotate, scale the column
    side, round both sides to FP16 (flush below 2^-14), covariance from the rounded operands, f^ = ||S^~||_F^2 in float32,
    directed-rounded row / column constants, lf = A_i + B_j - C_i D_j, excluded iff lf > 0 and 3.00004 f^ < lf^2.  Pairs
    are generated AROUND the similarity threshold (where a wrong exclusion would change the mask) for several shapes;
    an excluded pair must have lambda_max below the true threshold eigenvalue (float64 SVD of the unrounded data)."""
    from tscode_b200 import _host
    rng = np.random.default_rng(5)
    eps, thr = 1.05e-3, 0.5
    n_excl = n_tot = 0
    for sc in ([3.0, 3.0, 3.0], [6.0, 2.0, 1.0], [4.0, 4.0, 0.5], [8.0, 1.0, 1.0], [20.0, 0.5, 0.2]):
        for M in (12, 40, 80):
            base = rng.normal(size=(M, 3)) * np.array(sc)
            R0, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            base = base @ R0.T                                        # arbitrary orientation: the frame has to find it
            fr, _ = _host.screen_frame(base)
            Q, t = fr[:9].reshape(3, 3), fr[9:]
            e_thr = M * thr * thr * (1 + 1e-6)
            for k in range(300):
                s = rng.choice([0.1, 0.3, 0.42, 0.5, 0.6, 1.0])      # per-atom noise: rmsd from well below to above thr
                P = base + rng.normal(size=base.shape) * 0.05
                Y = P + rng.normal(size=base.shape) * s / np.sqrt(3.0)
                # ---- pack ----
                Pa, Yb = P @ Q.T, (Y @ Q.T) * t
                ha = Pa.astype(np.float16).astype(np.float64); ha[np.abs(Pa) < 2.0 ** -14] = 0.0
                hb = Yb.astype(np.float16).astype(np.float64); hb[np.abs(Yb) < 2.0 ** -14] = 0.0
                tiny_a = int(((Pa != 0) & (np.abs(Pa) < 2.0 ** -14)).sum())
                tiny_b = int(((Yb != 0) & (np.abs(Yb) < 2.0 ** -14)).sum())
                Gi, Gj = float((P ** 2).sum()), float((Y ** 2).sum())
                alpha = 2.0 ** -14 * (1 + 2.0 ** -10) / eps
                sg_i = np.sqrt(Gi) * (1 + 1e-12) + alpha * np.sqrt(tiny_a)
                sgb = np.sqrt(float((Yb ** 2).sum())) * (1 + 1e-12) + alpha * np.sqrt(tiny_b)
                sgu = np.sqrt(Gj) * (1 + 1e-12) + alpha / t.min() * np.sqrt(tiny_b)
                sg_j_row = np.sqrt(Gj) * (1 + 1e-12)
                hs = 0.5 * (1 - 1e-10)
                A = _f32_dir(hs * Gi - 0.5 * e_thr, up=False)
                Cc = _f32_dir(np.sqrt(3.0) * eps * sg_i, up=True)
                B = _f32_dir(hs * Gj, up=False)
                D = _f32_dir(max(sg_j_row, sgb, sgu), up=True)
                # ---- kernel ----
                S32 = (ha.T @ hb).astype(np.float32)                  # FP32 accumulators (products of FP16 values)
                f = np.float32(0)
                for v in S32.ravel():
                    f = np.float32(f + np.float32(v * v))
                ab = np.float32(A + B)
                lf = np.float32(np.float64(ab) - np.float64(Cc) * np.float64(D))        # one rounding: fma
                tv = np.float32(np.float64(np.float32(3.00004)) * np.float64(f) - np.float64(np.float32(lf * lf)))
                excluded = bool(lf > 0 and tv < 0)
                # ---- truth ----
                sv = np.linalg.svd(P.T @ Y, compute_uv=False)
                lam_max = sv[0] + sv[1] + (sv[2] if np.linalg.det(P.T @ Y) >= 0 else -sv[2])
                lam_true = 0.5 * (Gi + Gj - M * thr * thr)
                assert not (excluded and lam_max >= lam_true), (sc, M, s, lam_max, lam_true)
                n_excl += excluded; n_tot += 1
    assert 0.05 * n_tot < n_excl < n_tot                             # the stage does exclude, and not everything
