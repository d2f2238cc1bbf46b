// Test-only CPU harness around tscode_b200/csrc/tsc_math.cuh (the __host__ __device__ per-pair
// math of the RMSD kernels).  Lets the GPU-less container check the screen / eigen-solve
// numerics against the oracle.  Not part of the product library.
#include "../../tscode_b200/csrc/tsc_math.cuh"

extern "C" {

static void cov_and_g(const double* p, const double* q, int M, double S[9], double* G) {
    for (int i = 0; i < 9; i++) S[i] = 0;
    double g = 0;
    for (int m = 0; m < M; m++) {
        for (int a = 0; a < 3; a++) {
            g += p[3 * m + a] * p[3 * m + a] + q[3 * m + a] * q[3 * m + a];
            for (int b = 0; b < 3; b++) S[3 * a + b] += p[3 * m + a] * q[3 * m + b];
        }
    }
    *G = g;
}

// returns 1 if the screen keeps the pair as a candidate for rmsd < thr
int hm_screen(const double* p, const double* q, int M, double thr) {
    double S[9], G;
    cov_and_g(p, q, M, S, &G);
    double e = (double)M * thr * thr;
    return tsc::screen_candidate(S, G, e * (1.0 + 1e-6) + 1e-10 * G) ? 1 : 0;
}

// exact screen (no margin) — for checking the Budan-Fourier equivalence itself
int hm_screen_nomargin(const double* p, const double* q, int M, double thr) {
    double S[9], G;
    cov_and_g(p, q, M, S, &G);
    return tsc::screen_candidate(S, G, (double)M * thr * thr) ? 1 : 0;
}

// the verify path: explicit rotation, explicit differences
void hm_rmsd_and_max(const double* p, const double* q, int M, double* rmsd, double* maxdev, double* lam, double* gap) {
    double S[9], G, R[9];
    cov_and_g(p, q, M, S, &G);
    tsc::kabsch_rot_from_cov(S, R, lam, gap);
    double ss = 0, mx = 0;
    for (int m = 0; m < M; m++) {
        double x = p[3 * m], y = p[3 * m + 1], z = p[3 * m + 2];
        double dx = fma(R[0], x, fma(R[1], y, R[2] * z)) - q[3 * m];
        double dy = fma(R[3], x, fma(R[4], y, R[5] * z)) - q[3 * m + 1];
        double dz = fma(R[6], x, fma(R[7], y, R[8] * z)) - q[3 * m + 2];
        double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
        ss += d2;
        if (d2 > mx) mx = d2;
    }
    *rmsd = sqrt(ss / M);
    *maxdev = sqrt(mx);
}

// ---- FP32 second stage of the tcgen05 pre-screens (quartic32_*) ----
static float frob32(const float* S) {                 // the device's FFMA chain
    float f = 0.f;
    for (int q = 0; q < 9; q++) f = fmaf(S[q], S[q], f);
    return f;
}

// For n samples (9 floats each) and test points lam: forward errors of the FP32 values of P, P', P'' against
// long-double evaluation of the same polynomial, scaled by rho^4, rho^3, rho^2 (rho = max(lam, 2 ||S||_F)).
// out[3*k + 0..2]
void hm_q32_errors(const float* S, const float* lam, int n, double* out) {
    for (int k = 0; k < n; k++) {
        const float* s = S + 9 * k;
        float p0, p1, p2;
        tsc::quartic32_values<tsc::OpsF32>(s, frob32(s), lam[k], p0, p1, p2);
        long double Sd[9], f = 0;
        for (int q = 0; q < 9; q++) { Sd[q] = s[q]; f += Sd[q] * Sd[q]; }
        long double K[4][4];
        K[0][0] = Sd[0] + Sd[4] + Sd[8]; K[0][1] = Sd[5] - Sd[7]; K[0][2] = Sd[6] - Sd[2]; K[0][3] = Sd[1] - Sd[3];
        K[1][1] = Sd[0] - Sd[4] - Sd[8]; K[1][2] = Sd[1] + Sd[3]; K[1][3] = Sd[6] + Sd[2];
        K[2][2] = -Sd[0] + Sd[4] - Sd[8]; K[2][3] = Sd[5] + Sd[7]; K[3][3] = -Sd[0] - Sd[4] + Sd[8];
        for (int a = 0; a < 4; a++) for (int b = 0; b < a; b++) K[a][b] = K[b][a];
        // determinant by cofactor expansion (exact enough in long double)
        auto det3 = [&](int r0, int r1, int r2, int c0, int c1, int c2) {
            return K[r0][c0] * (K[r1][c1] * K[r2][c2] - K[r1][c2] * K[r2][c1])
                 - K[r0][c1] * (K[r1][c0] * K[r2][c2] - K[r1][c2] * K[r2][c0])
                 + K[r0][c2] * (K[r1][c0] * K[r2][c1] - K[r1][c1] * K[r2][c0]);
        };
        long double c0 = K[0][0] * det3(1, 2, 3, 1, 2, 3) - K[0][1] * det3(1, 2, 3, 0, 2, 3)
                       + K[0][2] * det3(1, 2, 3, 0, 1, 3) - K[0][3] * det3(1, 2, 3, 0, 1, 2);
        long double dS = Sd[0] * (Sd[4] * Sd[8] - Sd[5] * Sd[7]) - Sd[1] * (Sd[3] * Sd[8] - Sd[5] * Sd[6])
                       + Sd[2] * (Sd[3] * Sd[7] - Sd[4] * Sd[6]);
        long double c1 = -8 * dS, c2 = -2 * f, l = lam[k];
        long double P0 = ((l * l + c2) * l + c1) * l + c0, P1 = (4 * l * l + 2 * c2) * l + c1, P2 = 12 * l * l + 2 * c2;
        long double rho = fmaxl(fabsl(l), 2 * sqrtl(f));
        if (rho == 0) rho = 1;
        out[3 * k + 0] = (double)(fabsl(P0 - p0) / (rho * rho * rho * rho));
        out[3 * k + 1] = (double)(fabsl(P1 - p1) / (rho * rho * rho));
        out[3 * k + 2] = (double)(fabsl(P2 - p2) / (rho * rho));
    }
}

// decision of the FP32 stage and, beside it, lambda_max of the same (float) covariance in FP64
int hm_q32_excluded(const float* S, float lam, double* lam_max) {
    double Sd[9];
    for (int q = 0; q < 9; q++) Sd[q] = S[q];
    double qv[4], gap;
    *lam_max = tsc::key_top_eigen(tsc::key_matrix(Sd), qv, &gap);
    return tsc::quartic32_excluded(S, frob32(S), lam) ? 1 : 0;
}

// the T = S^T S form of the same stage (component-sequential screen): decision + FP64 lambda_max
int hm_q32t_excluded(const float* S, float lam, double* lam_max) {
    double Sd[9];
    for (int q = 0; q < 9; q++) Sd[q] = S[q];
    double qv[4], gap;
    *lam_max = tsc::key_top_eigen(tsc::key_matrix(Sd), qv, &gap);
    float t[6];
    tsc::quartic32_T_from_rows(S, t);
    return tsc::quartic32_T_excluded(t, lam) ? 1 : 0;
}

// the same with a SCALED column side (rmsd_screen.cu, ScFrame): the accumulators hold S^ = S diag(t) (here: the float
// products s_ab * t_b, rounded — the MMA's own sums of FP16 products are exact to within the operand bound), T^ is
// accumulated from its rows, un-scaled entry by entry with the float constants 1 / (t_b t_c) as the kernel does, and the
// quartic runs on that T.  lam_max is that of the UNscaled float covariance in FP64.
int hm_q32t_excluded_scaled(const float* S, const double* t3, float lam, double* lam_max) {
    double Sd[9];
    for (int q = 0; q < 9; q++) Sd[q] = S[q];
    double qv[4], gap;
    *lam_max = tsc::key_top_eigen(tsc::key_matrix(Sd), qv, &gap);
    float Ss[9], t[6];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) Ss[3 * a + b] = (float)((double)S[3 * a + b] * t3[b]);
    tsc::quartic32_T_from_rows(Ss, t);
    const int bb[6] = {0, 1, 2, 0, 0, 1}, cc[6] = {0, 1, 2, 1, 2, 2};
    for (int q = 0; q < 6; q++) t[q] = tsc::OpsF32::mul(t[q], (float)(1.0 / (t3[bb[q]] * t3[cc[q]])));
    return tsc::quartic32_T_excluded(t, lam) ? 1 : 0;
}

// d = the bound on |det S| the T form uses (before the 1 + 4u scale), next to |det S| in long double, and f
void hm_q32t_det(const float* S, int n, double* out) {
    for (int k = 0; k < n; k++) {
        const float* s = S + 9 * k;
        float t[6], f, c0, da;
        tsc::quartic32_T_from_rows(s, t);
        tsc::quartic32_T_coeffs<tsc::OpsF32>(t, f, c0, da);
        long double Sd[9];
        for (int q = 0; q < 9; q++) Sd[q] = s[q];
        long double dS = Sd[0] * (Sd[4] * Sd[8] - Sd[5] * Sd[7]) - Sd[1] * (Sd[3] * Sd[8] - Sd[5] * Sd[6])
                       + Sd[2] * (Sd[3] * Sd[7] - Sd[4] * Sd[6]);
        out[3 * k + 0] = (double)(da >= 1.17549435e-38f ? sqrtf(da) : 0.f);
        out[3 * k + 1] = (double)fabsl(dS);
        out[3 * k + 2] = (double)f;
    }
}
}
