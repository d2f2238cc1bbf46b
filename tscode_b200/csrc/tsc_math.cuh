// tsc_math.cuh — per-pair closed-form math shared by the RMSD kernels.
//
// Everything here is __host__ __device__ so that tests/hostmath (a test-only harness compiled
// with nvcc for the CPU) can check the numerics in the GPU-less build container.  The product
// library only ever calls these from device code.
//
// Reference semantics being reproduced: rmsd_and_max_numba, tscode/rmsd_pruning.py:6-41
// (rotation-only Kabsch about the origin, improper-rotation fix, RMSD and max deviation).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TSC_HD __host__ __device__ __forceinline__
#else
#define TSC_HD inline
#endif

namespace tsc {

// ------------------------------------------------------------------------------------------
// Horn's 4x4 key matrix of the cross-covariance S (S[a][b] = sum_m p_m[a] * q_m[b]).
// Its largest eigenvalue is max over PROPER rotations R of sum_m (R p_m) . q_m, i.e.
// sigma1 + sigma2 + sign(det S) * sigma3 — exactly what Kabsch with the reflection fix
// (rmsd_pruning.py:20-23) attains.  Stored as the 10 upper-triangular entries.
// ------------------------------------------------------------------------------------------
struct Key4 {
    double k00, k01, k02, k03, k11, k12, k13, k22, k23, k33;
};

TSC_HD Key4 key_matrix(const double S[9]) {
    Key4 K;
    K.k00 = S[0] + S[4] + S[8];
    K.k01 = S[5] - S[7];
    K.k02 = S[6] - S[2];
    K.k03 = S[1] - S[3];
    K.k11 = S[0] - S[4] - S[8];
    K.k12 = S[1] + S[3];
    K.k13 = S[6] + S[2];
    K.k22 = -S[0] + S[4] - S[8];
    K.k23 = S[5] + S[7];
    K.k33 = -S[0] - S[4] + S[8];
    return K;
}

// Characteristic polynomial of the (traceless) key matrix: P(x) = x^4 + c2 x^2 + c1 x + c0,
//   c2 = -2 ||S||_F^2,  c1 = -8 det S,  c0 = det K.
TSC_HD void key_charpoly(const double S[9], double& c2, double& c1, double& c0) {
    double f = S[0] * S[0];
    f = fma(S[1], S[1], f); f = fma(S[2], S[2], f);
    f = fma(S[3], S[3], f); f = fma(S[4], S[4], f); f = fma(S[5], S[5], f);
    f = fma(S[6], S[6], f); f = fma(S[7], S[7], f); f = fma(S[8], S[8], f);
    c2 = -2.0 * f;
    double d = S[0] * fma(S[4], S[8], -S[5] * S[7]);
    d = fma(-S[1], fma(S[3], S[8], -S[5] * S[6]), d);
    d = fma(S[2], fma(S[3], S[7], -S[4] * S[6]), d);
    c1 = -8.0 * d;
    Key4 K = key_matrix(S);
    // det of the symmetric 4x4 through 2x2 minors of rows (0,1) and rows (2,3)
    double a01 = fma(K.k00, K.k11, -K.k01 * K.k01);
    double a02 = fma(K.k00, K.k12, -K.k02 * K.k01);
    double a03 = fma(K.k00, K.k13, -K.k03 * K.k01);
    double a12 = fma(K.k01, K.k12, -K.k02 * K.k11);
    double a13 = fma(K.k01, K.k13, -K.k03 * K.k11);
    double a23 = fma(K.k02, K.k13, -K.k03 * K.k12);
    double b01 = fma(K.k02, K.k13, -K.k12 * K.k03);
    double b02 = fma(K.k02, K.k23, -K.k22 * K.k03);
    double b03 = fma(K.k02, K.k33, -K.k23 * K.k03);
    double b12 = fma(K.k12, K.k23, -K.k22 * K.k13);
    double b13 = fma(K.k12, K.k33, -K.k23 * K.k13);
    double b23 = fma(K.k22, K.k33, -K.k23 * K.k23);
    double det = a01 * b23;
    det = fma(-a02, b13, det);
    det = fma(a03, b12, det);
    det = fma(a12, b03, det);
    det = fma(-a13, b02, det);
    det = fma(a23, b01, det);
    c0 = det;
}

// ------------------------------------------------------------------------------------------
// Screen: can this pair possibly have rmsd < thr ?
//   E_opt = G - 2*lambda_max  (G = |p|^2 + |q|^2),   rmsd^2 = E_opt / M.
//   rmsd < thr  <=>  lambda_max > lam_t := (G - M*thr^2) / 2.
// All roots of P are real (K symmetric), so by Budan-Fourier the number of roots above x
// equals the number of sign changes in (P, P', P'', P''', P'''')(x).  P'''' = 24 > 0, hence
//   "no root above lam_t"  <=>  P, P', P'', P''' all > 0 at lam_t.
// No eigen-solve, no iteration, no divergence.  `e_thr_pad` is M*thr^2 plus a safety margin
// (callers use M*thr^2*(1+1e-6) + 1e-10*G): every pair the screen passes is re-evaluated
// exactly (rotation applied, explicit differences) before its bit is trusted, so the margin
// only costs a few extra verifications and can never lose a true pair.
// Returns true when the pair is a CANDIDATE (cannot be excluded).
// ------------------------------------------------------------------------------------------
// Rounding guard.  The sign test is only as good as the evaluated polynomial: for an ensemble far from the
// origin (|centroid| ~ 7e4 A) two roots of P sit within ~1e3 of each other at ~1e11, P(lam_t) ~ 1e26 drowns in the
// ~1e28 rounding noise of the coefficients, and the unguarded test LOST true pairs (caught by
// test_f16_screen_extreme_coordinates).  Every root and lam_t are below r = lam_t + 2 ||S||_F, the evaluation
// errors of P, P', P'' (coefficients through 2x2 minors + Horner) are below ~30 eps (r/2)^4-ish; a pair is
// excluded only if the three values clear 4e-13 r^4, 16e-13 r^3, 48e-13 r^2 (with r^2 over-estimated by
// 2 (lam_t^2 + 4 ||S||_F^2), which needs no square root) — for ordinary, centred ensembles
// (r ~ 1e3..1e4) that is ~1e2 against values of 1e10 and more, so no screening power is lost.
TSC_HD bool quartic_excluded(double lam, double c2, double c1, double c0) {
    if (!(lam > 0.0)) return false;
    const double l2 = lam * lam;
    const double R2 = 2.0 * fma(-2.0, c2, l2);                 // 2 (lam^2 + 4 ||S||_F^2) >= r^2  (no square root)
    const double tol = 4e-13;
    const double p2 = fma(12.0, l2, 2.0 * c2);                 // P''
    const double p1 = fma(fma(4.0, l2, 2.0 * c2), lam, c1);    // P'
    const double p0 = fma(fma(l2 + c2, lam, c1), lam, c0);     // P
    // P > tol r^4,  P' > 4 tol r^3 (compared through squares),  P'' > 12 tol r^2;  NaN -> not excluded
    return (p0 > tol * R2 * R2) & (p1 > 0.0) & (p1 * p1 > 16.0 * tol * tol * R2 * R2 * R2) & (p2 > 12.0 * tol * R2);
}

TSC_HD bool screen_candidate(const double S[9], double G, double e_thr_pad) {
    double lam = 0.5 * (G - e_thr_pad);
    if (!(lam > 0.0)) return true;
    double c2, c1, c0;
    key_charpoly(S, c2, c1, c0);
    return !quartic_excluded(lam, c2, c1, c0);
}

// ------------------------------------------------------------------------------------------
// FP32 form of the sign test: second stage of the tcgen05 pre-screen (rmsd_screen.cu).
//
// Samuelson's bound sqrt(3) ||S||_F >= lambda_max is only sharp for near-isotropic covariances; for an
// elongated or planar molecule it exceeds the threshold eigenvalue for EVERY pair (tools/aniso_probe.py:
// the FP64 test then ran for all pairs and the screen was 3x slower).  The quartic's values P, P', P'' at a
// test point lam <= lam_t are therefore also evaluated in FP32, straight from the FP32 accumulators, and
// a pair is excluded only if they clear forward error bounds; whatever this stage cannot exclude becomes a
// candidate and is decided exactly by the verify kernel, so the bounds only have to be safe, not sharp.
//
// Coefficients: c2 = -2 f (f = ||S||_F^2), c1 = -8 det S, and — the roots being the signed sums
// +-s1 +-s2 +-s3 of the singular values — c0 = det K = sum s^4 - 2 sum s_i^2 s_j^2 = 2 ||S^T S||_F^2 - f^2,
// which costs 28 operations through the six entries of S^T S instead of 44 through the 2x2 minors of K and
// has a much smaller error bound.
// Error bounds (standard model, u = 2^-24, s = ||S||_F, rho = max(lam, 2 s) >= every root and lam):
//   T = S^T S: |T~_ab - T_ab| <= 3 u |col_a| |col_b|;  p4 = ||T||_F^2 <= f^2, error <= 13 u f^2;  f~ = f (1 + 9 u)
//   c0: error <= 46 u s^4 <= 3 u rho^4;   c1 = -8 det S: error <= 72 u s^3 <= 9 u rho^3;   c2: error <= 4.5 u rho^2
//   Horner: |P~ - P| <= 26 u rho^4,  |P'~ - P'| <= 36 u rho^3,  |P''~ - P''| <= 34 u rho^2
// The tolerances are 64 u R^2, 72 u R^(3/2), 68 u R with R = lam^2 + 4 f in [rho^2, 2 rho^2] (no max, no square
// root: P' is compared through squares): at least twice the bounds.  tests/test_hostmath.py measures the actual
// errors against long-double evaluation on 2e5 covariances of five kinds (isotropic, elongated, rank one, far from
// the origin, sparse): a few u.
// The operations are written through a policy class so that the host (float) and the device (two pairs per
// instruction, fma / mul / add / sub .f32x2) run the same sequence.
// ------------------------------------------------------------------------------------------
struct OpsF32 {
    typedef float T;
#ifdef __CUDA_ARCH__
    static TSC_HD T fma(T a, T b, T c) { return __fmaf_rn(a, b, c); }
    static TSC_HD T mul(T a, T b) { return __fmul_rn(a, b); }
    static TSC_HD T add(T a, T b) { return __fadd_rn(a, b); }
    static TSC_HD T sub(T a, T b) { return __fsub_rn(a, b); }
#else
    static TSC_HD T fma(T a, T b, T c) { return fmaf(a, b, c); }
    static TSC_HD T mul(T a, T b) { volatile float r = a * b; return r; }      // (no contraction on the host)
    static TSC_HD T add(T a, T b) { volatile float r = a + b; return r; }
    static TSC_HD T sub(T a, T b) { volatile float r = a - b; return r; }
#endif
    static TSC_HD T bc(float x) { return x; }
};

constexpr float Q32_T0 = 3.82e-6f;      // 64 u
constexpr float Q32_T1SQ = 1.85e-11f;   // (72 u)^2 = 1.842e-11, rounded up
constexpr float Q32_T2 = 4.06e-6f;      // 68 u

// P, P', P'' of the key-matrix quartic of S at lam;  f = ||S||_F^2 (already computed by the caller)
template <class O>
TSC_HD void quartic32_values(const typename O::T* S, typename O::T f, typename O::T lam, typename O::T& p0,
                             typename O::T& p1, typename O::T& p2) {
    typedef typename O::T T;
    // T = S^T S (S row-major: S[3 k + a])
    const T t00 = O::fma(S[6], S[6], O::fma(S[3], S[3], O::mul(S[0], S[0])));
    const T t11 = O::fma(S[7], S[7], O::fma(S[4], S[4], O::mul(S[1], S[1])));
    const T t22 = O::fma(S[8], S[8], O::fma(S[5], S[5], O::mul(S[2], S[2])));
    const T t01 = O::fma(S[6], S[7], O::fma(S[3], S[4], O::mul(S[0], S[1])));
    const T t02 = O::fma(S[6], S[8], O::fma(S[3], S[5], O::mul(S[0], S[2])));
    const T t12 = O::fma(S[7], S[8], O::fma(S[4], S[5], O::mul(S[1], S[2])));
    const T dg = O::fma(t22, t22, O::fma(t11, t11, O::mul(t00, t00)));
    const T og = O::fma(t12, t12, O::fma(t02, t02, O::mul(t01, t01)));
    const T p4 = O::fma(og, O::bc(2.0f), dg);
    const T c0 = O::sub(O::mul(p4, O::bc(2.0f)), O::mul(f, f));
    // c1 = -8 det S
    const T m0 = O::sub(O::mul(S[4], S[8]), O::mul(S[5], S[7]));
    const T m1 = O::sub(O::mul(S[5], S[6]), O::mul(S[3], S[8]));
    const T m2 = O::sub(O::mul(S[3], S[7]), O::mul(S[4], S[6]));
    T d = O::mul(S[0], m0);
    d = O::fma(S[1], m1, d);
    d = O::fma(S[2], m2, d);
    const T c1 = O::mul(d, O::bc(-8.0f));
    const T l2 = O::mul(lam, lam);
    const T c2 = O::mul(f, O::bc(-2.0f)), c2x2 = O::mul(f, O::bc(-4.0f));
    p2 = O::fma(O::bc(12.0f), l2, c2x2);
    p1 = O::fma(O::fma(O::bc(4.0f), l2, c2x2), lam, c1);
    p0 = O::fma(O::fma(O::add(l2, c2), lam, c1), lam, c0);
}

// ------------------------------------------------------------------------------------------
// The same test from T = S^T S alone (component-sequential screen, rmsd_screen.cu).
//
// That kernel receives the covariance one ROW at a time (row a of S = the three sums over atoms of
// x_a(i) * x_b(j), b = 0..2: one accumulator buffer per (tile, a)) and can afford to keep six numbers per
// pair between rows, not nine.  T = S^T S is the sum of the outer products of the rows, so it accumulates
// row by row (mul for the first row, fma for the others: the very operation order of quartic32_values);
// f = tr T and c0 = 2 ||T||_F^2 - f^2 follow as before.  Only det S couples the rows.  It is replaced by
//     d >= |det S|,    d = sqrt(max(det T~, 0) + 8 u f^3) (1 + 4 u)        (det T = (det S)^2 exactly)
// and c1 by -8 d.  For x >= 0 the true quartic P(x) = x^4 - 2 f x^2 - 8 det S x + c0 is then >= the evaluated
// one, Pd(x); if Pd and its first three derivatives (the third is 24 lam) are positive at lam, Taylor's formula
// gives Pd > 0 on [lam, inf), hence P > 0 there and lambda_max < lam: the pair is excluded.  When det S >= 0
// (every pair that could be similar: the optimal superposition is then a proper rotation already) nothing is
// lost against the signed form beyond the margin under the root.
// Margin: every Leibniz term of det T is at most T00 T11 T22 <= (f/3)^3; the entries of T~ carry a relative
// error <= 3 u of sqrt(T_aa T_bb), each term thus <= 9 u f^3 / 27, six terms 2 u f^3; the evaluation (nine
// operations on quantities <= 2 f^3 / 27) adds less than u f^3.  8 u f^3 is more than twice the sum.
// sqrt is the approximate flush-to-zero instruction on the device (error <= 2 ulp; the factor 1 + 4 u covers it and
// the addition under the root; arguments below 2^-126 give 0, see quartic32_decide).  P and P' take -8 d through the same Horner steps as before, so the forward-error
// tolerances of quartic32_margins (which even include an allowance for c1's own rounding) stay valid.
// ------------------------------------------------------------------------------------------
// With a scaled column side (rmsd_screen.cu, ScFrame) the kernel accumulates T^ = diag(t) T diag(t) and multiplies every
// entry by the float constant 1 / (t_b t_c) before this stage: one more rounding of the constant and one of the product,
// i.e. entries with 5 u instead of 3 u of relative error.  The budgets above scale accordingly — p4: 17 u f^2, c0: 57 u
// s^4 <= 3.6 u rho^4 (total for P still < 30 u rho^4 against the 64 u R^2 tolerance), the Leibniz terms of det T:
// 3.3 u f^3 + u f^3 of evaluation = 4.3 u f^3 against the 8 u f^3 margin — and stay inside the tolerances
// (tests/test_hostmath.py: test_fp32_T_form_with_scaled_columns_stays_sound).
constexpr float Q32_DET_MARGIN = 4.77e-7f;    // 8 u
constexpr float Q32_C1_SCALE = -8.000004f;    // -8 (1 + 4 u), rounded away from zero

// t[6] = (T00, T11, T22, T01, T02, T12) -> f = tr T, c0, and det_arg = det T~ + 8 u f^3 (the caller clamps it at
// zero and takes the root: packed types have no square-root instruction)
template <class O>
TSC_HD void quartic32_T_coeffs(const typename O::T* t, typename O::T& f, typename O::T& c0, typename O::T& det_arg) {
    typedef typename O::T T;
    f = O::add(O::add(t[0], t[1]), t[2]);
    const T dg = O::fma(t[2], t[2], O::fma(t[1], t[1], O::mul(t[0], t[0])));
    const T og = O::fma(t[5], t[5], O::fma(t[4], t[4], O::mul(t[3], t[3])));
    const T p4 = O::fma(og, O::bc(2.0f), dg);
    const T ff = O::mul(f, f);
    c0 = O::sub(O::mul(p4, O::bc(2.0f)), ff);
    // det T expanded along the first row
    const T m0 = O::sub(O::mul(t[1], t[2]), O::mul(t[5], t[5]));
    const T m1 = O::sub(O::mul(t[5], t[4]), O::mul(t[3], t[2]));
    const T m2 = O::sub(O::mul(t[3], t[5]), O::mul(t[1], t[4]));
    T d = O::mul(t[0], m0);
    d = O::fma(t[3], m1, d);
    d = O::fma(t[4], m2, d);
    det_arg = O::fma(O::mul(ff, f), O::bc(Q32_DET_MARGIN), d);
}

// P, P', P'' at lam with c1 = Q32_C1_SCALE * d  (d >= 0: the caller's root of max(det_arg, 0))
template <class O>
TSC_HD void quartic32_T_values(typename O::T f, typename O::T c0, typename O::T d, typename O::T lam, typename O::T& p0,
                               typename O::T& p1, typename O::T& p2) {
    typedef typename O::T T;
    const T c1 = O::mul(d, O::bc(Q32_C1_SCALE));
    const T l2 = O::mul(lam, lam);
    const T c2 = O::mul(f, O::bc(-2.0f)), c2x2 = O::mul(f, O::bc(-4.0f));
    p2 = O::fma(O::bc(12.0f), l2, c2x2);
    p1 = O::fma(O::fma(O::bc(4.0f), l2, c2x2), lam, c1);
    p0 = O::fma(O::fma(O::add(l2, c2), lam, c1), lam, c0);
}

// margins of the three values over their tolerances: the pair is excluded iff lam, p1, m0, m1, m2 are all > 0
template <class O>
TSC_HD void quartic32_margins(typename O::T p0, typename O::T p1, typename O::T p2, typename O::T f, typename O::T lam,
                              typename O::T& m0, typename O::T& m1, typename O::T& m2) {
    typedef typename O::T T;
    const T R = O::fma(f, O::bc(4.0f), O::mul(lam, lam));
    const T R2 = O::mul(R, R);
    m0 = O::fma(R2, O::bc(-Q32_T0), p0);
    m2 = O::fma(R, O::bc(-Q32_T2), p2);
    m1 = O::fma(O::mul(R2, R), O::bc(-Q32_T1SQ), O::mul(p1, p1));
}

// true = all roots of the quartic are provably below lam (NaN / inf anywhere -> false).
// lam has to exceed Q32_LAM_MIN rather than 0: the device takes the root of det_arg with the flush-to-zero
// approximate instruction (one MUFU instead of a five-instruction sequence), i.e. d = 0 for det_arg < 2^-126, short of
// the true bound by at most 1.1e-19; then f <= 2.9e-11 (det_arg >= 8 u f^3), the terms -8 d lam of P and -8 d of P' are
// off by <= 8.7e-19 lam and 8.7e-19, and the unused halves of the tolerances, 32 u R^2 >= 1.9e-6 lam^4 and
// 36 u R^(3/2) >= 2.1e-6 lam^3, cover that as soon as lam >= 7.7e-5.  (A threshold eigenvalue that small means
// G_i + G_j ~ M thr^2, a molecule the size of the threshold; the pair simply stays a candidate.)
// On the device: two three-input NaN-propagating minima and one comparison instead of five comparisons and four selects.
constexpr float Q32_LAM_MIN = 1e-4f;
TSC_HD bool quartic32_decide(float lam, float p1, float m0, float m1, float m2) {
#ifdef __CUDA_ARCH__
    float t, lg = __fsub_rn(lam, Q32_LAM_MIN);              // > 0 iff lam > Q32_LAM_MIN (exact near the cut: Sterbenz)
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(m0), "f"(m1), "f"(m2));
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(t), "f"(p1), "f"(lg));
    return t > 0.0f;
#else
    return (lam > Q32_LAM_MIN) & (p1 > 0.0f) & (m0 > 0.0f) & (m1 > 0.0f) & (m2 > 0.0f);
#endif
}

TSC_HD bool quartic32_excluded(const float S[9], float f, float lam) {
    float p0, p1, p2, m0, m1, m2;
    quartic32_values<OpsF32>(S, f, lam, p0, p1, p2);
    quartic32_margins<OpsF32>(p0, p1, p2, f, lam, m0, m1, m2);
    return quartic32_decide(lam, p1, m0, m1, m2);
}

// scalar form of the T-based test (host model / tests; the device runs the packed form of the same sequence)
TSC_HD bool quartic32_T_excluded(const float t[6], float lam) {
    float f, c0, da, p0, p1, p2, m0, m1, m2;
    quartic32_T_coeffs<OpsF32>(t, f, c0, da);
    // a NaN argument gives d = 0, but then f is NaN too: not excluded; below the smallest normal the device's root is 0
    const float d = da >= 1.17549435e-38f ? sqrtf(da) : 0.0f;
    quartic32_T_values<OpsF32>(f, c0, d, lam, p0, p1, p2);
    quartic32_margins<OpsF32>(p0, p1, p2, f, lam, m0, m1, m2);
    return quartic32_decide(lam, p1, m0, m1, m2);
}

// T = S^T S accumulated row by row as the component-sequential screen does (S row-major)
TSC_HD void quartic32_T_from_rows(const float S[9], float t[6]) {
    t[0] = OpsF32::mul(S[0], S[0]); t[1] = OpsF32::mul(S[1], S[1]); t[2] = OpsF32::mul(S[2], S[2]);
    t[3] = OpsF32::mul(S[0], S[1]); t[4] = OpsF32::mul(S[0], S[2]); t[5] = OpsF32::mul(S[1], S[2]);
    for (int a = 1; a < 3; a++) {
        const float x = S[3 * a], y = S[3 * a + 1], z = S[3 * a + 2];
        t[0] = OpsF32::fma(x, x, t[0]); t[1] = OpsF32::fma(y, y, t[1]); t[2] = OpsF32::fma(z, z, t[2]);
        t[3] = OpsF32::fma(x, y, t[3]); t[4] = OpsF32::fma(x, z, t[4]); t[5] = OpsF32::fma(y, z, t[5]);
    }
}

// ------------------------------------------------------------------------------------------
// Largest eigenpair of the key matrix by cyclic Jacobi (always converges, any eigenvector of a
// degenerate top eigenspace is an optimal rotation).  q = (q0, q1, q2, q3) unit quaternion,
// scalar first.  Returns lambda_max; *gap receives lambda_1 - lambda_2.
// ------------------------------------------------------------------------------------------
TSC_HD double key_top_eigen(const Key4& K, double q[4], double* gap) {
    double A[4][4] = {{K.k00, K.k01, K.k02, K.k03},
                      {K.k01, K.k11, K.k12, K.k13},
                      {K.k02, K.k12, K.k22, K.k23},
                      {K.k03, K.k13, K.k23, K.k33}};
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    double scale = 0.0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) scale += A[i][j] * A[i][j];
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0;
        for (int i = 0; i < 3; i++)
            for (int j = i + 1; j < 4; j++) off += A[i][j] * A[i][j];
        if (off <= 1e-32 * scale) break;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int r = p + 1; r < 4; r++) {
                double apr = A[p][r];
                if (apr == 0.0) continue;
                double theta = (A[r][r] - A[p][p]) / (2.0 * apr);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                double c = 1.0 / sqrt(fma(t, t, 1.0)), s = t * c;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    double akp = A[k][p], akr = A[k][r];
                    A[k][p] = c * akp - s * akr;
                    A[k][r] = s * akp + c * akr;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    double apk = A[p][k], ark = A[r][k];
                    A[p][k] = c * apk - s * ark;
                    A[r][k] = s * apk + c * ark;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    double vkp = V[k][p], vkr = V[k][r];
                    V[k][p] = c * vkp - s * vkr;
                    V[k][r] = s * vkp + c * vkr;
                }
            }
    }
    double best = A[0][0], second = -1e300;
    int bi = 0;
#pragma unroll
    for (int i = 1; i < 4; i++) {
        if (A[i][i] > best) { second = best; best = A[i][i]; bi = i; }
        else if (A[i][i] > second) second = A[i][i];
    }
    // select column bi without dynamic indexing (keeps V in registers on the device)
#pragma unroll
    for (int k = 0; k < 4; k++)
        q[k] = (bi == 0) ? V[k][0] : (bi == 1) ? V[k][1] : (bi == 2) ? V[k][2] : V[k][3];
    double n = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
    if (gap) *gap = best - second;
    return best;
}

// Rotation matrix (column-vector convention, x' = R x) of a unit quaternion, scalar first.
TSC_HD void quat_to_rot(const double q[4], double R[9]) {
    double q00 = q[0] * q[0], q11 = q[1] * q[1], q22 = q[2] * q[2], q33 = q[3] * q[3];
    double q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
    double q12 = q[1] * q[2], q13 = q[1] * q[3], q23 = q[2] * q[3];
    R[0] = q00 + q11 - q22 - q33; R[1] = 2.0 * (q12 - q03);       R[2] = 2.0 * (q13 + q02);
    R[3] = 2.0 * (q12 + q03);       R[4] = q00 - q11 + q22 - q33; R[5] = 2.0 * (q23 - q01);
    R[6] = 2.0 * (q13 - q02);       R[7] = 2.0 * (q23 + q01);       R[8] = q00 - q11 - q22 + q33;
}

// Optimal proper rotation taking p onto q (x' = R x) from their cross-covariance.
// The reference's `rot_mat` (rmsd_pruning.py:26, row-vector convention p @ rot) is R^T.
TSC_HD void kabsch_rot_from_cov(const double S[9], double R[9], double* lambda_max, double* gap) {
    Key4 K = key_matrix(S);
    double q[4];
    double lam = key_top_eigen(K, q, gap);
    if (lambda_max) *lambda_max = lam;
    quat_to_rot(q, R);
}

}  // namespace tsc
