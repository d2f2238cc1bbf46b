// tfd_moi.cu — the two cheap similarity filters that run right before the RMSD prune in
// Embedder.similarity_refining (tscode/embedder.py:1325-1352):
//
//   prune_conformers_tfd (numba_functions.py:142-231): float32 torsion fingerprints (dihedral of every
//     quadruplet, algebra.py:24-57), pair test  sum_q | |a_q - b_q| wrapped to [0, 180] | < thresh
//     (numba_functions.py:241-256), grouping loop identical to prune_conformers_rmsd_rot_corr's;
//   prune_by_moment_of_inertia (optimization_methods.py:327-358): principal moments of inertia of the heavy
//     atoms (algebra.py:166-187), pair test  all(|I_i - I_j| / I_i < max_deviation), first match per row
//     (algebra.py:189-203), connected components.
//
// Both loops stop at the first similar later structure of a row, so — exactly like the stateless rot_corr
// path (rotcorr.cu) — all the host replay needs is  first_hit[i] = min{ j > i : similar(i, j) }.
// One warp per row, 32 partners per step, early exit; the per-structure features are tiny (Q floats / 3
// doubles) and stay L2-resident.  Integer / FP32-compare work, HBM traffic negligible: reported as time.
#include "tsc_common.cuh"

namespace tsc {

// algebra.py:24-57 (Praxeolitic formula), FP64, degrees; stored as float32 like the reference's tf_mat
__global__ void __launch_bounds__(256) tfd_fingerprint_kernel(const double* __restrict__ S, int64_t N, int A,
                                                              const int32_t* __restrict__ quads, int Q,
                                                              float* __restrict__ tf) {
    const int64_t total = N * Q;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / Q;
        const int q = (int)(e % Q);
        const double* X = S + i * (int64_t)A * 3;
        const double* p0 = X + 3 * quads[4 * q], *p1 = X + 3 * quads[4 * q + 1];
        const double* p2 = X + 3 * quads[4 * q + 2], *p3 = X + 3 * quads[4 * q + 3];
        const double b0[3] = {-1.0 * (p1[0] - p0[0]), -1.0 * (p1[1] - p0[1]), -1.0 * (p1[2] - p0[2])};
        double b1[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        const double b2[3] = {p3[0] - p2[0], p3[1] - p2[1], p3[2] - p2[2]};
        const double n1 = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
        b1[0] /= n1; b1[1] /= n1; b1[2] /= n1;
        const double d0 = b0[0] * b1[0] + b0[1] * b1[1] + b0[2] * b1[2];
        const double d2 = b2[0] * b1[0] + b2[1] * b1[1] + b2[2] * b1[2];
        const double v[3] = {b0[0] - d0 * b1[0], b0[1] - d0 * b1[1], b0[2] - d0 * b1[2]};
        const double w[3] = {b2[0] - d2 * b1[0], b2[1] - d2 * b1[1], b2[2] - d2 * b1[2]};
        const double x = v[0] * w[0] + v[1] * w[1] + v[2] * w[2];
        const double c[3] = {b1[1] * v[2] - b1[2] * v[1], b1[2] * v[0] - b1[0] * v[2], b1[0] * v[1] - b1[1] * v[0]};
        const double y = c[0] * w[0] + c[1] * w[1] + c[2] * w[2];
        tf[e] = (float)(atan2(y, x) * (180.0 / 3.14159265358979323846));
    }
}

// numba_functions.py:241-256: deltas = |a - b| in float32; wrapped and summed in float64 (the reference's
// `deltas - (deltas > 180) * 360` promotes to float64); every term is a float32 value, so the sum is exact in
// FP64 whatever the order.
__device__ __forceinline__ double tfd_sum(const float* __restrict__ a, const float* __restrict__ b, int Q) {
    double s = 0.0;
    for (int q = 0; q < Q; q++) {
        const float d = fabsf(a[q] - b[q]);
        s += fabs((double)d - (d > 180.0f ? 360.0 : 0.0));
    }
    return s;
}

__global__ void __launch_bounds__(256) tfd_scan_kernel(const float* __restrict__ tf, int64_t N, int Q, double thresh,
                                                       int32_t* __restrict__ first_hit, unsigned long long* near_count) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long near = 0;
    for (int64_t i = warp_g; i < N; i += nwarps) {
        const float* a = tf + i * Q;
        int32_t hit = (int32_t)N;
        for (int64_t j0 = i + 1; j0 < N; j0 += 32) {
            const int64_t j = j0 + lane;
            bool sim = false;
            if (j < N) {
                const double s = tfd_sum(a, tf + j * Q, Q);
                sim = s < thresh;
                near += fabs(s - thresh) < 1e-4;
            }
            const uint32_t b = __ballot_sync(0xffffffffu, sim);
            if (b) { hit = (int32_t)(j0 + __ffs(b) - 1); break; }
        }
        if (lane == 0) first_hit[i] = hit;
    }
    if (near_count && near) atomicAdd(near_count, near);
}

// ---- moments of inertia -----------------------------------------------------------------------
// algebra.py:166-187: centre of mass removed, I_ab = sum_n m_n (|r_n|^2 delta_ab - r_na r_nb), principal
// moments ordered by absolute value (ascending; they are non-negative).  The reference diagonalises with
// numpy's general eig and B^-1 A B; here a cyclic Jacobi on the symmetric 3x3 — same values to ~1e-13 relative.
__device__ __forceinline__ void jacobi3(double A[3][3], double ev[3]) {
    for (int sweep = 0; sweep < 30; sweep++) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        const double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
        if (off <= 1e-32 * diag) break;
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int r = p + 1; r < 3; r++) {
                const double apr = A[p][r];
                if (apr == 0.0) continue;
                const double theta = (A[r][r] - A[p][p]) / (2.0 * apr);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                const double c = 1.0 / sqrt(fma(t, t, 1.0)), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double akp = A[k][p], akr = A[k][r];
                    A[k][p] = c * akp - s * akr;
                    A[k][r] = s * akp + c * akr;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double apk = A[p][k], ark = A[r][k];
                    A[p][k] = c * apk - s * ark;
                    A[r][k] = s * apk + c * ark;
                }
            }
    }
    double a = A[0][0], b = A[1][1], c = A[2][2], t;
    if (fabs(a) > fabs(b)) { t = a; a = b; b = t; }
    if (fabs(b) > fabs(c)) { t = b; b = c; c = t; }
    if (fabs(a) > fabs(b)) { t = a; a = b; b = t; }
    ev[0] = a; ev[1] = b; ev[2] = c;
}

__global__ void __launch_bounds__(256) moi_moments_kernel(const double* __restrict__ S, int64_t N, int A,
                                                          const int32_t* __restrict__ heavy_idx, int M,
                                                          const double* __restrict__ masses, double* __restrict__ moments) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp_g; i < N; i += nwarps) {
        const double* X = S + i * (int64_t)A * 3;
        double mt = 0, cx = 0, cy = 0, cz = 0;
        for (int m = lane; m < M; m += 32) {
            const double* r = X + 3 * heavy_idx[m];
            const double w = masses[m];
            mt += w; cx = fma(r[0], w, cx); cy = fma(r[1], w, cy); cz = fma(r[2], w, cz);
        }
        mt = warp_sum(mt); cx = warp_sum(cx) / mt; cy = warp_sum(cy) / mt; cz = warp_sum(cz) / mt;
        double xx = 0, yy = 0, zz = 0, xy = 0, xz = 0, yz = 0;
        for (int m = lane; m < M; m += 32) {
            const double* r = X + 3 * heavy_idx[m];
            const double w = masses[m], x = r[0] - cx, y = r[1] - cy, z = r[2] - cz;
            xx = fma(w, y * y + z * z, xx); yy = fma(w, x * x + z * z, yy); zz = fma(w, x * x + y * y, zz);
            xy = fma(-w, x * y, xy); xz = fma(-w, x * z, xz); yz = fma(-w, y * z, yz);
        }
        xx = warp_sum(xx); yy = warp_sum(yy); zz = warp_sum(zz); xy = warp_sum(xy); xz = warp_sum(xz); yz = warp_sum(yz);
        if (lane == 0) {
            double T[3][3] = {{xx, xy, xz}, {xy, yy, yz}, {xz, yz, zz}}, ev[3];
            jacobi3(T, ev);
            moments[3 * i] = ev[0]; moments[3 * i + 1] = ev[1]; moments[3 * i + 2] = ev[2];
        }
    }
}

// algebra.py:189-203: first j > i with all(|I_i - I_j| / I_i < max_deviation)
__global__ void __launch_bounds__(256) moi_scan_kernel(const double* __restrict__ moments, int64_t N, double max_dev,
                                                       int32_t* __restrict__ first_hit, unsigned long long* near_count) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long near = 0;
    for (int64_t i = warp_g; i < N; i += nwarps) {
        const double a0 = moments[3 * i], a1 = moments[3 * i + 1], a2 = moments[3 * i + 2];
        int32_t hit = (int32_t)N;
        for (int64_t j0 = i + 1; j0 < N; j0 += 32) {
            const int64_t j = j0 + lane;
            bool sim = false;
            if (j < N) {
                const double r0 = fabs(a0 - moments[3 * j]) / a0, r1 = fabs(a1 - moments[3 * j + 1]) / a1,
                             r2 = fabs(a2 - moments[3 * j + 2]) / a2;
                sim = (r0 < max_dev) && (r1 < max_dev) && (r2 < max_dev);
                const double worst = fmax(r0, fmax(r1, r2));
                near += fabs(worst - max_dev) < 1e-9 * max_dev;
            }
            const uint32_t b = __ballot_sync(0xffffffffu, sim);
            if (b) { hit = (int32_t)(j0 + __ffs(b) - 1); break; }
        }
        if (lane == 0) first_hit[i] = hit;
    }
    if (near_count && near) atomicAdd(near_count, near);
}

// ---- constraint scores of embedded poses -----------------------------------------------------------
// _score_embed_poses (numba_functions.py:273-288): scores[j] = sum_i | |x_i1 - x_i2| - d_i |, accumulated in
// float32;  fitness_check (optimization_methods.py:544-557): error = sum_i ( |x_a - x_b| - target_i ) over the
// constraints whose target is not None (signed, float64), verdict error < threshold.  One thread per structure;
// the norm is evaluated with separately rounded multiplies and adds like numba's norm_of (algebra.py:90-96).
__global__ void __launch_bounds__(256) constraint_score_kernel(const double* __restrict__ S, int64_t P, int A,
                                                               const int32_t* __restrict__ cons, const double* __restrict__ targets,
                                                               int K, int per_pose, float* __restrict__ score_abs32,
                                                               double* __restrict__ error_signed) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const double* X = S + p * (int64_t)A * 3;
        const int32_t* c = cons + (per_pose ? p * K * 2 : 0);
        const double* tg = targets + (per_pose ? p * K : 0);
        float s32 = 0.0f;
        double err = 0.0;
        for (int k = 0; k < K; k++) {
            const double t = tg[k];
            const double* a = X + 3 * c[2 * k], *b = X + 3 * c[2 * k + 1];
            const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
            const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            if (t == t) {                       // NaN marks a constraint without target (None)
                s32 = (float)((double)s32 + fabs(d - t));          // float32 accumulator += float64 term
                err += d - t;
            }
        }
        if (score_abs32) score_abs32[p] = s32;
        if (error_signed) error_signed[p] = err;
    }
}

}  // namespace tsc

extern "C" int tsc_tfd_fingerprints(const double* S, int64_t N, int32_t A, const int32_t* quads, int32_t Q, float* tf,
                                    void* stream) {
    if (N <= 0 || Q <= 0) return 0;
    int64_t blocks = (N * Q + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tsc::tfd_fingerprint_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, N, A, quads, Q, tf);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_tfd_scan(const float* tf, int64_t N, int32_t Q, double thresh, int32_t* first_hit,
                            uint64_t* near_count, void* stream) {
    if (N <= 0) return 0;
    int64_t blocks = (N + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tsc::tfd_scan_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        tf, N, Q, thresh, first_hit, reinterpret_cast<unsigned long long*>(near_count));
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_moi_moments(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M,
                               const double* masses, double* moments, void* stream) {
    if (N <= 0 || M <= 0) return 0;
    int64_t blocks = (N + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tsc::moi_moments_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, N, A, heavy_idx, M, masses, moments);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_moi_scan(const double* moments, int64_t N, double max_deviation, int32_t* first_hit,
                            uint64_t* near_count, void* stream) {
    if (N <= 0) return 0;
    int64_t blocks = (N + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tsc::moi_scan_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        moments, N, max_deviation, first_hit, reinterpret_cast<unsigned long long*>(near_count));
    TSC_CHECK_LAUNCH();
    return 0;
}

// scores of P structures S (P, A, 3) against K distance constraints: cons (K, 2) int32 atom pairs and targets (K)
// shared by all structures (per_pose = 0) or (P, K, 2) / (P, K) per structure (per_pose = 1); a NaN target skips the
// constraint.  score_abs32 (P) float32 = _score_embed_poses; error_signed (P) float64 = fitness_check's error.
extern "C" int tsc_constraint_scores(const double* S, int64_t P, int32_t A, const int32_t* cons, const double* targets,
                                     int32_t K, int32_t per_pose, float* score_abs32, double* error_signed, void* stream) {
    if (P <= 0) return 0;
    int64_t blocks = (P + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tsc::constraint_score_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, P, A, cons, targets, K, per_pose,
                                                                                   score_abs32, error_signed);
    TSC_CHECK_LAUNCH();
    return 0;
}
