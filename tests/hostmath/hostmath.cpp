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
}
