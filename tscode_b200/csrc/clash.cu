// clash.cu — fused rigid-body pose transform + inter-fragment clash screen.
//
// Reference per pose (tscode/embeds.py:116-118, 713-714, 841-842):
//     pose = get_embed(mols, conf_ids)              # concat((R_k @ X_k.T).T + t_k), embeds.py:961-969
//     ok   = compenetration_check(pose, ids, thresh, max_clashes)      # numba_functions.py:59-105
// compenetration_check counts atom pairs of different fragments closer than `thresh`
// (all_dists, algebra.py:98-157, then `< thresh`) over (m2,m1) [, (m3,m2), (m1,m3)] and
// passes the pose iff the count is <= max_clashes; the early returns between the three
// blocks cannot change the verdict because counts only grow.
//
// Here one warp owns one pose: fragment conformers are read from the (L2-resident) fragment
// library, rotated + translated on the fly into a per-warp shared-memory SoA image, and the
// pair tests run from there — the pose is never materialised in HBM.  Per pose the kernel
// reads F*(9+3) doubles + F ints and writes one verdict byte.
//     d < thresh  is evaluated as  d^2 < t2  with t2 the exact image of the threshold under
// correctly-rounded sqrt (computed on the host), so no sqrt is needed and the comparison is
// the same predicate.  A warp stops as soon as count > max_clashes.
//
// Roofline: FP64 FMA pipe (6 FP64 instructions per atom pair); bytes are negligible.
#include "tsc_common.cuh"
#include "tsc_math.cuh"

namespace tsc {

constexpr int CLASH_WARPS = 8;

struct ClashThresh {
    double t2;        // d < thresh  <=>  d2 < t2
    double near_lo2;  // (thresh - 1e-9)^2   (REPORT mode only)
    double near_hi2;  // (thresh + 1e-9)^2
};

// count pairs (a in A-set, b in B-set) with d2 < t2; stop once total > max_clashes.
// Register-tiled: every lane keeps two atoms of the B set in registers (64 per sweep) and walks the A set, whose
// coordinates are shared-memory broadcasts — 3 loads + 2 x 6 FP64 instructions per 64 pairs.  (First version:
// lanes strided over the flattened (a, b) pair index with a ballot per 32 pairs; ncu: 1 554 warp instructions per
// pose on C2, issue slots 73 % busy.)  The per-lane hit counts are summed over the warp (REDUX) every 8 A atoms
// for the early exit.  Same predicate per pair, so the verdict is unchanged.
template <bool REPORT>
__device__ __forceinline__ bool count_block(const double* xa, const double* ya, const double* za, int na,
                                            const double* xb, const double* yb, const double* zb, int nb,
                                            const ClashThresh& th, long long max_clashes, long long& count,
                                            unsigned long long& near, int lane) {
    for (int b0 = 0; b0 < nb; b0 += 64) {
        const int b1 = b0 + lane, b2 = b0 + 32 + lane;
        const bool v1 = b1 < nb, v2 = b2 < nb;
        // absent atoms sit at 1e200: d2 overflows to +inf, never below any threshold
        const double x1 = v1 ? xb[b1] : 1e200, y1 = v1 ? yb[b1] : 1e200, z1 = v1 ? zb[b1] : 1e200;
        const double x2 = v2 ? xb[b2] : 1e200, y2 = v2 ? yb[b2] : 1e200, z2 = v2 ? zb[b2] : 1e200;
        auto pair2 = [&](int a, int& cnt) {
            const double ax = xa[a], ay = ya[a], az = za[a];
            {
                const double dx = ax - x1, dy = ay - y1, dz = az - z1;
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                cnt += d2 < th.t2;
                if (REPORT) near += (d2 > th.near_lo2 && d2 < th.near_hi2);
            }
            {
                const double dx = ax - x2, dy = ay - y2, dz = az - z2;
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                cnt += d2 < th.t2;
                if (REPORT) near += (d2 > th.near_lo2 && d2 < th.near_hi2);
            }
        };
        int a = 0;
        for (; a + 8 <= na; a += 8) {                 // 8 A atoms per early-exit check, no per-atom branches
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < 8; u++) pair2(a + u, cnt);
            count += __reduce_add_sync(0xffffffffu, cnt);
            if (!REPORT && count > max_clashes) return true;
        }
        if (a < na) {
            int cnt = 0;
            for (; a < na; a++) pair2(a, cnt);
            count += __reduce_add_sync(0xffffffffu, cnt);
            if (!REPORT && count > max_clashes) return true;
        }
    }
    return count > max_clashes;
}

template <bool REPORT>
__device__ __forceinline__ int verdict_fragments(const double* sx, const double* sy, const double* sz,
                                                 const int* off, const int* n, int F, const ClashThresh& th,
                                                 long long max_clashes, unsigned long long& near, int lane) {
    long long count = 0;
    bool over;
    // (m2, m1)
    over = count_block<REPORT>(sx + off[1], sy + off[1], sz + off[1], n[1], sx + off[0], sy + off[0], sz + off[0],
                               n[0], th, max_clashes, count, near, lane);
    if (over && !REPORT) return 0;
    if (F == 3) {
        over = count_block<REPORT>(sx + off[2], sy + off[2], sz + off[2], n[2], sx + off[1], sy + off[1],
                                   sz + off[1], n[1], th, max_clashes, count, near, lane);
        if (over && !REPORT) return 0;
        over = count_block<REPORT>(sx + off[0], sy + off[0], sz + off[0], n[0], sx + off[2], sy + off[2],
                                   sz + off[2], n[2], th, max_clashes, count, near, lane);
    }
    return count > max_clashes ? 0 : 1;
}

// ids=None branch: count_clashes over the full symmetric matrix, (d < 0.5) & (d > 0); each
// close pair is counted twice (numba_functions.py:49-56).
__device__ __forceinline__ int verdict_intramolecular(const double* sx, const double* sy, const double* sz, int A,
                                                      double t2_half, long long max_clashes, int lane) {
    long long count = 0;
    for (int a = 0; a + 1 < A; a++) {
        for (int b0 = a + 1; b0 < A; b0 += 32) {
            const int b = b0 + lane;
            bool hit = false;
            if (b < A) {
                const double dx = sx[a] - sx[b], dy = sy[a] - sy[b], dz = sz[a] - sz[b];
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                hit = (d2 < t2_half) && (d2 > 0.0);
            }
            count += 2 * __popc(__ballot_sync(0xffffffffu, hit));
            if (count > max_clashes) return 0;
        }
    }
    return 1;
}

template <bool REPORT, int F>       // F (2 or 3) is a template parameter: fragment loops unroll, offsets stay in registers
__global__ void __launch_bounds__(CLASH_WARPS * 32) embed_clash_kernel(
    const double* __restrict__ frag_lib, const int64_t* __restrict__ frag_off, const int32_t* __restrict__ n_atoms,
    const int32_t* __restrict__ conf, const double* __restrict__ R, const double* __restrict__ t, int64_t P,
    int A_total, ClashThresh th, long long max_clashes, uint8_t* __restrict__ verdict, unsigned long long* near_out) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sx = smem + (size_t)warp * 3 * A_total;
    double* sy = sx + A_total;
    double* sz = sy + A_total;
    int off[3] = {0, 0, 0}, n[3] = {0, 0, 0};
    const double* Xbase[3] = {nullptr, nullptr, nullptr};
    {
        int o = 0;
#pragma unroll
        for (int k = 0; k < F; k++) { n[k] = n_atoms[k]; off[k] = o; o += n[k]; Xbase[k] = frag_lib + frag_off[k]; }
    }
    unsigned long long near = 0;
    const int64_t warp_g = (int64_t)blockIdx.x * CLASH_WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * CLASH_WARPS;
    // (requesting the next pose's parameters ahead — lanes fetching R ++ t together, read back with shuffles — was
    // measured slower, 0.180 vs 0.166 ms on C2: the kernel is bound by issue slots, not by that round trip)
    for (int64_t p = warp_g; p < P; p += nwarps) {
#pragma unroll
        for (int k = 0; k < F; k++) {
            const double* X = Xbase[k] + (int64_t)conf[p * F + k] * 3 * n[k];
            const double* r = R + (p * F + k) * 9;
            const double* tt = t + (p * F + k) * 3;
            const double r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4], r5 = r[5], r6 = r[6], r7 = r[7],
                         r8 = r[8], t0 = tt[0], t1 = tt[1], t2 = tt[2];
            for (int a = lane; a < n[k]; a += 32) {
                const double x = X[3 * a], y = X[3 * a + 1], z = X[3 * a + 2];
                sx[off[k] + a] = fma(r2, z, fma(r1, y, r0 * x)) + t0;
                sy[off[k] + a] = fma(r5, z, fma(r4, y, r3 * x)) + t1;
                sz[off[k] + a] = fma(r8, z, fma(r7, y, r6 * x)) + t2;
            }
        }
        __syncwarp();
        const int v = verdict_fragments<REPORT>(sx, sy, sz, off, n, F, th, max_clashes, near, lane);
        if (lane == 0) verdict[p] = (uint8_t)v;
        __syncwarp();
    }
    if (REPORT) {
        near = __reduce_add_sync(0xffffffffu, (unsigned)near);
        if (lane == 0 && near && near_out) atomicAdd(near_out, near);
    }
}

template <bool REPORT>
__global__ void __launch_bounds__(CLASH_WARPS * 32) clash_structs_kernel(
    const double* __restrict__ S, int64_t P, int A, const int32_t* __restrict__ ids, int F, ClashThresh th,
    double t2_half, long long max_clashes, uint8_t* __restrict__ verdict, unsigned long long* near_out) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sx = smem + (size_t)warp * 3 * A;
    double* sy = sx + A;
    double* sz = sy + A;
    int off[3] = {0, 0, 0}, n[3] = {0, 0, 0};
    if (F == 2) {             // only ids[0] is read (numba_functions.py:77-78)
        n[0] = ids[0]; n[1] = A - n[0]; off[1] = n[0];
    } else if (F == 3) {
        n[0] = ids[0]; n[1] = ids[1]; n[2] = A - n[0] - n[1]; off[1] = n[0]; off[2] = n[0] + n[1];
    }
    unsigned long long near = 0;
    const int64_t warp_g = (int64_t)blockIdx.x * CLASH_WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * CLASH_WARPS;
    for (int64_t p = warp_g; p < P; p += nwarps) {
        const double* X = S + p * (int64_t)A * 3;
        for (int e = lane; e < 3 * A; e += 32) {          // coalesced AoS read, SoA scatter
            const double v = X[e];
            const int a = e / 3, c = e - 3 * a;
            (c == 0 ? sx : c == 1 ? sy : sz)[a] = v;
        }
        __syncwarp();
        int v;
        if (F == 0) v = verdict_intramolecular(sx, sy, sz, A, t2_half, max_clashes, lane);
        else v = verdict_fragments<REPORT>(sx, sy, sz, off, n, F, th, max_clashes, near, lane);
        if (lane == 0) verdict[p] = (uint8_t)v;
        __syncwarp();
    }
    if (REPORT) {
        near = __reduce_add_sync(0xffffffffu, (unsigned)near);
        if (lane == 0 && near && near_out) atomicAdd(near_out, near);
    }
}

// get_embed for a selected set of poses: S_out[q] = pose keep_idx[q]   (embeds.py:961-969)
__global__ void __launch_bounds__(256) embed_gather_kernel(const double* __restrict__ frag_lib,
                                                           const int64_t* __restrict__ frag_off,
                                                           const int32_t* __restrict__ n_atoms, int F,
                                                           const int32_t* __restrict__ conf,
                                                           const double* __restrict__ R, const double* __restrict__ t,
                                                           const int64_t* __restrict__ keep_idx, int64_t n_keep,
                                                           int A_total, double* __restrict__ S_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t warp_g = (int64_t)blockIdx.x * 8 + warp;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t q = warp_g; q < n_keep; q += nwarps) {
        const int64_t p = keep_idx ? keep_idx[q] : q;
        int o = 0;
        for (int k = 0; k < F; k++) {
            const int nk = n_atoms[k];
            const double* X = frag_lib + frag_off[k] + (int64_t)conf[p * F + k] * 3 * nk;
            const double* r = R + (p * F + k) * 9;
            const double* tt = t + (p * F + k) * 3;
            for (int a = lane; a < nk; a += 32) {
                const double x = X[3 * a], y = X[3 * a + 1], z = X[3 * a + 2];
                double* d = S_out + (q * A_total + o + a) * 3;
                d[0] = fma(r[2], z, fma(r[1], y, r[0] * x)) + tt[0];
                d[1] = fma(r[5], z, fma(r[4], y, r[3] * x)) + tt[1];
                d[2] = fma(r[8], z, fma(r[7], y, r[6] * x)) + tt[2];
            }
            o += nk;
        }
    }
}

static int clash_grid(int64_t P) {
    int64_t blocks = (P + CLASH_WARPS - 1) / CLASH_WARPS;
    const int64_t cap = 148 * 8;         // 8 resident CTAs of 8 warps per SM
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// ------------------------------------------------------------------------------------------
// On-device pose parameters of the string embed (tscode/embeds.py:91-114): the reference builds, per pose and on
// the host, mol2.rotation = rot_mat_from_pointer(ref_vec, angle) @ rotation_matrix_from_vectors(mol_vec, -ref_vec)
// (the first factor only if angle != 0) and mol2.position = p1 - rotation @ p2, 15-60 us of numpy small-array
// overhead each, then ships nothing: it screens the pose immediately.  Here the whole pose space
//   conformers (c1, c2) x reactive centres (ai1, ai2) x systematic angles, in the reference's loop order,
// is generated by one kernel into the (conf, R, t) arrays the fused clash screen consumes, so that only the
// small per-conformer tables of centres and orbital vectors cross PCIe.
//   c1t/v1t: (n_conf1, n_c1, 3) centres / orbital vectors of molecule 1, c2t/v2t likewise for molecule 2;
//   sin_half/cos_half/nonzero (n_ang): host-evaluated sin, cos of angle/2 (algebra.py:337-341) and angle != 0;
//   flip (9): rot_mat_from_pointer([0,0,1], 180) as the host evaluates it (antiparallel case, utils.py:203-205).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__global__ void __launch_bounds__(256) string_embed_params_kernel(
    const double* __restrict__ c1t, const double* __restrict__ v1t, const double* __restrict__ c2t,
    const double* __restrict__ v2t, int n_conf1, int n_conf2, int n_c1, int n_c2, const double* __restrict__ sin_half,
    const double* __restrict__ cos_half, const uint8_t* __restrict__ nonzero, int n_ang, const double* __restrict__ flip,
    int64_t P, int32_t* __restrict__ conf, double* __restrict__ R, double* __restrict__ t) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = p;
        const int k = (int)(r % n_ang); r /= n_ang;
        const int ai2 = (int)(r % n_c2); r /= n_c2;
        const int ai1 = (int)(r % n_c1); r /= n_c1;
        const int c2 = (int)(r % n_conf2); r /= n_conf2;
        const int c1 = (int)r;
        const double* p1 = c1t + ((int64_t)c1 * n_c1 + ai1) * 3;
        const double* rv = v1t + ((int64_t)c1 * n_c1 + ai1) * 3;        // ref_vec
        const double* p2 = c2t + ((int64_t)c2 * n_c2 + ai2) * 3;
        const double* mv = v2t + ((int64_t)c2 * n_c2 + ai2) * 3;        // mol_vec
        // rotation_matrix_from_vectors(mol_vec, -ref_vec)                (utils.py:183-208)
        // norms and cross product with separately rounded multiplies and adds, as numba / numpy evaluate them: for
        // (anti)parallel vectors the cross product is pure rounding noise that Rodrigues' formula then divides by,
        // so an FMA-contracted cross product would give a different (equally arbitrary) matrix than the reference
        const double na = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(mv[0], mv[0]), __dmul_rn(mv[1], mv[1])), __dmul_rn(mv[2], mv[2])));
        const double nb = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(rv[0], rv[0]), __dmul_rn(rv[1], rv[1])), __dmul_rn(rv[2], rv[2])));
        const double a0 = mv[0] / na, a1 = mv[1] / na, a2 = mv[2] / na;
        const double b0 = -rv[0] / nb, b1 = -rv[1] / nb, b2 = -rv[2] / nb;
        const double vx = __dsub_rn(__dmul_rn(a1, b2), __dmul_rn(a2, b1));
        const double vy = __dsub_rn(__dmul_rn(a2, b0), __dmul_rn(a0, b2));
        const double vz = __dsub_rn(__dmul_rn(a0, b1), __dmul_rn(a1, b0));
        const double s = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
        double Rv[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (s != 0.0) {
            const double c = a0 * b0 + a1 * b1 + a2 * b2;
            const double f = (1.0 - c) / (s * s);
            const double K[9] = {0, -vz, vy, vz, 0, -vx, -vy, vx, 0};
            double K2[9];
            mat3_mul(K, K, K2);
#pragma unroll
            for (int e = 0; e < 9; e++) Rv[e] = (Rv[e] + K[e]) + K2[e] * f;
        } else {
            const double sx = a0 + b0, sy = a1 + b1, sz = a2 + b2;
            if (sqrt(__dadd_rn(__dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy)), __dmul_rn(sz, sz))) == 0.0) {
#pragma unroll
                for (int e = 0; e < 9; e++) Rv[e] = flip[e];
            }
        }
        double R2[9];
        if (nonzero[k]) {        // delta_rot = rot_mat_from_pointer(ref_vec, angle) (algebra.py:325-344)
            const double u0 = rv[0] / nb, u1 = rv[1] / nb, u2 = rv[2] / nb;
            const double q1 = sin_half[k] * u0, q2 = sin_half[k] * u1, q3 = sin_half[k] * u2, q0 = cos_half[k];
            const double D[9] = {2 * (q0 * q0 + q1 * q1) - 1, 2 * (q1 * q2 - q0 * q3),     2 * (q1 * q3 + q0 * q2),
                                 2 * (q1 * q2 + q0 * q3),     2 * (q0 * q0 + q2 * q2) - 1, 2 * (q2 * q3 - q0 * q1),
                                 2 * (q1 * q3 - q0 * q2),     2 * (q2 * q3 + q0 * q1),     2 * (q0 * q0 + q3 * q3) - 1};
            mat3_mul(D, Rv, R2);
        } else {
#pragma unroll
            for (int e = 0; e < 9; e++) R2[e] = Rv[e];
        }
        double* Ro = R + p * 18;
        double* to = t + p * 6;
#pragma unroll
        for (int e = 0; e < 9; e++) { Ro[e] = (e % 4 == 0) ? 1.0 : 0.0; Ro[9 + e] = R2[e]; }
        to[0] = 0.0; to[1] = 0.0; to[2] = 0.0;
        to[3] = p1[0] - (R2[0] * p2[0] + R2[1] * p2[1] + R2[2] * p2[2]);
        to[4] = p1[1] - (R2[3] * p2[0] + R2[4] * p2[1] + R2[5] * p2[2]);
        to[5] = p1[2] - (R2[6] * p2[0] + R2[7] * p2[1] + R2[8] * p2[2]);
        conf[2 * p] = c1; conf[2 * p + 1] = c2;
    }
}

// ------------------------------------------------------------------------------------------
// On-device pose parameters of the cyclical embeds (tscode/embeds.py:657-709).  For one group (a combination of
// conformers, pivots and polygon orientation) and molecule i the reference computes, per pose,
//     A   = align_vec_pair([end - start, direction_i], [pivot_i, mol_direction_i])        # 3x3 SVD, :682
//     axis = A @ (rc0 - rc1)  or  A @ pivot_i;  Sr = rot_mat_from_pointer(axis, angle_i)   # :688-696
//     rotation = Sr @ A;  position = A @ apm - Sr @ (A @ apm) + mean(vec_pair) - A @ meanpoint   # :698-708
// Only Sr depends on the pose (its angle), so a first kernel does the per-(group, molecule) part once — the
// alignment is the optimal rotation of a two-vector correlation, obtained here as the top eigenvector of Horn's
// key matrix instead of an SVD with reflection fix (same rotation whenever it is unique) — and a second kernel
// expands groups x angle combinations into the (R, t) arrays of the fused clash screen.
// ------------------------------------------------------------------------------------------
__global__ void cyclical_group_kernel(const double* __restrict__ ref2, const double* __restrict__ tgt2,
                                      const double* __restrict__ axis_src, const double* __restrict__ apm,
                                      const double* __restrict__ vmean, const double* __restrict__ pmean, int64_t n,
                                      double* __restrict__ A_out, double* __restrict__ axis_out,
                                      double* __restrict__ cor_out, double* __restrict__ pos_out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const double* rf = ref2 + e * 6;
        const double* tg = tgt2 + e * 6;
        // cross-covariance taking tgt onto ref: S[a][b] = sum_j tgt[j][a] * ref[j][b]  (= B^T of algebra.py:266-272)
        double S[9];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) S[3 * a + b] = tg[a] * rf[b] + tg[3 + a] * rf[3 + b];
        double A[9];
        kabsch_rot_from_cov(S, A, nullptr, nullptr);
        const double* u = axis_src + e * 3, *c = apm + e * 3, *vm = vmean + e * 3, *pm = pmean + e * 3;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            axis_out[e * 3 + r] = A[3 * r] * u[0] + A[3 * r + 1] * u[1] + A[3 * r + 2] * u[2];
            cor_out[e * 3 + r] = A[3 * r] * c[0] + A[3 * r + 1] * c[1] + A[3 * r + 2] * c[2];
            pos_out[e * 3 + r] = vm[r] - (A[3 * r] * pm[0] + A[3 * r + 1] * pm[1] + A[3 * r + 2] * pm[2]);
        }
#pragma unroll
        for (int q = 0; q < 9; q++) A_out[e * 9 + q] = A[q];
    }
}

__global__ void __launch_bounds__(256) cyclical_pose_kernel(
    const double* __restrict__ A, const double* __restrict__ axis, const double* __restrict__ cor,
    const double* __restrict__ pos, const int32_t* __restrict__ gconf, const int32_t* __restrict__ combos,
    const double* __restrict__ sin_half, const double* __restrict__ cos_half, int64_t G, int F, int64_t C,
    int32_t* __restrict__ conf, double* __restrict__ R, double* __restrict__ t) {
    const int64_t total = G * C * F;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % F);
        const int64_t pose = e / F, g = pose / C, c = pose % C;
        const int64_t gi = g * F + i;
        const int k = combos[c * F + i];
        const double* ax = axis + gi * 3;
        const double nrm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);        // rot_mat_from_pointer, algebra.py:325-344
        const double q1 = sin_half[k] * (ax[0] / nrm), q2 = sin_half[k] * (ax[1] / nrm), q3 = sin_half[k] * (ax[2] / nrm),
                     q0 = cos_half[k];
        const double Sr[9] = {2 * (q0 * q0 + q1 * q1) - 1, 2 * (q1 * q2 - q0 * q3),     2 * (q1 * q3 + q0 * q2),
                              2 * (q1 * q2 + q0 * q3),     2 * (q0 * q0 + q2 * q2) - 1, 2 * (q2 * q3 - q0 * q1),
                              2 * (q1 * q3 - q0 * q2),     2 * (q2 * q3 + q0 * q1),     2 * (q0 * q0 + q3 * q3) - 1};
        double Rm[9];
        mat3_mul(Sr, A + gi * 9, Rm);
        const double* cr = cor + gi * 3;
        const double* ps = pos + gi * 3;
        double* Ro = R + e * 9;
        double* to = t + e * 3;
#pragma unroll
        for (int q = 0; q < 9; q++) Ro[q] = Rm[q];
#pragma unroll
        for (int r = 0; r < 3; r++)
            to[r] = (cr[r] - (Sr[3 * r] * cr[0] + Sr[3 * r + 1] * cr[1] + Sr[3 * r + 2] * cr[2])) + ps[r];
        conf[e] = gconf[gi];
    }
}

}  // namespace tsc

extern "C" int tsc_embed_clash(const double* frag_lib, const int64_t* frag_off, const int32_t* n_atoms, int32_t F,
                               int32_t A_total, const int32_t* conf, const double* R, const double* t, int64_t P,
                               double t2, double thresh, int64_t max_clashes, uint8_t* verdict,
                               uint64_t* near_count, void* stream) {
    using namespace tsc;
    if (P <= 0) return 0;
    if (F != 2 && F != 3) return (int)cudaErrorInvalidValue;
    ClashThresh th{t2, (thresh - 1e-9) * (thresh - 1e-9), (thresh + 1e-9) * (thresh + 1e-9)};
    const size_t smem = (size_t)CLASH_WARPS * 3 * A_total * sizeof(double);
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    cudaError_t e;
    if (near_count) {
        auto kern = F == 2 ? embed_clash_kernel<true, 2> : embed_clash_kernel<true, 3>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<clash_grid(P), CLASH_WARPS * 32, smem, (cudaStream_t)stream>>>(
            frag_lib, frag_off, n_atoms, conf, R, t, P, A_total, th, max_clashes, verdict,
            reinterpret_cast<unsigned long long*>(near_count));
    } else {
        auto kern = F == 2 ? embed_clash_kernel<false, 2> : embed_clash_kernel<false, 3>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<clash_grid(P), CLASH_WARPS * 32, smem, (cudaStream_t)stream>>>(
            frag_lib, frag_off, n_atoms, conf, R, t, P, A_total, th, max_clashes, verdict, nullptr);
    }
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_clash_structs(const double* S, int64_t P, int32_t A, const int32_t* ids, int32_t F, double t2,
                                 double thresh, double t2_half, int64_t max_clashes, uint8_t* verdict,
                                 uint64_t* near_count, void* stream) {
    using namespace tsc;
    if (P <= 0) return 0;
    if (F != 0 && F != 2 && F != 3) return (int)cudaErrorInvalidValue;
    ClashThresh th{t2, (thresh - 1e-9) * (thresh - 1e-9), (thresh + 1e-9) * (thresh + 1e-9)};
    const size_t smem = (size_t)CLASH_WARPS * 3 * A * sizeof(double);
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    cudaError_t e;
    if (near_count && F != 0) {
        e = cudaFuncSetAttribute(clash_structs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        clash_structs_kernel<true><<<clash_grid(P), CLASH_WARPS * 32, smem, (cudaStream_t)stream>>>(
            S, P, A, ids, F, th, t2_half, max_clashes, verdict, reinterpret_cast<unsigned long long*>(near_count));
    } else {
        e = cudaFuncSetAttribute(clash_structs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        clash_structs_kernel<false><<<clash_grid(P), CLASH_WARPS * 32, smem, (cudaStream_t)stream>>>(
            S, P, A, ids, F, th, t2_half, max_clashes, verdict, nullptr);
    }
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_embed_gather(const double* frag_lib, const int64_t* frag_off, const int32_t* n_atoms, int32_t F,
                                int32_t A_total, const int32_t* conf, const double* R, const double* t,
                                const int64_t* keep_idx, int64_t n_keep, double* S_out, void* stream) {
    using namespace tsc;
    if (n_keep <= 0) return 0;
    int64_t blocks = (n_keep + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    embed_gather_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(frag_lib, frag_off, n_atoms, F, conf, R,
                                                                           t, keep_idx, n_keep, A_total, S_out);
    TSC_CHECK_LAUNCH();
    return 0;
}

// Pose parameters of the whole string-embed pose space, P = n_conf1 * n_conf2 * n_c1 * n_c2 * n_ang poses in the
// reference's loop order (embeds.py:91-114): conf (P, 2) int32, R (P, 2, 3, 3), t (P, 2, 3), ready for
// tsc_embed_clash / tsc_embed_gather.
extern "C" int tsc_string_embed_params(const double* centers1, const double* vecs1, const double* centers2,
                                       const double* vecs2, int32_t n_conf1, int32_t n_conf2, int32_t n_c1, int32_t n_c2,
                                       const double* sin_half, const double* cos_half, const uint8_t* nonzero,
                                       int32_t n_ang, const double* flip, int32_t* conf, double* R, double* t,
                                       void* stream) {
    const int64_t P = (int64_t)n_conf1 * n_conf2 * n_c1 * n_c2 * n_ang;
    if (P <= 0) return 0;
    int64_t blocks = (P + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tsc::string_embed_params_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        centers1, vecs1, centers2, vecs2, n_conf1, n_conf2, n_c1, n_c2, sin_half, cos_half, nonzero, n_ang, flip, P, conf,
        R, t);
    TSC_CHECK_LAUNCH();
    return 0;
}

// Pose parameters of cyclical embeds (embeds.py:657-709) for G groups x C angle combinations x F molecules.
//   per (group, molecule), (G, F, ...) arrays: ref2 (2, 3) = [end - start, direction], tgt2 (2, 3) = [pivot,
//   mol_direction], axis_src (3) = rc0 - rc1 or pivot, apm (3) = atomic_pivot_mean, vmean (3) = mean(vec_pair),
//   pmean (3) = pivot.meanpoint, gconf int32 = conformer index;  combos (C, F) int32 = index of every molecule's
//   angle in the table;  sin_half / cos_half: host-evaluated sin, cos of angle/2.
//   scratch: G*F*18 doubles.  Out: conf (G*C, F) int32, R (G*C, F, 3, 3), t (G*C, F, 3); pose index = g*C + c.
extern "C" int tsc_cyclical_embed_params(const double* ref2, const double* tgt2, const double* axis_src, const double* apm,
                                         const double* vmean, const double* pmean, const int32_t* gconf, int64_t G,
                                         int32_t F, const int32_t* combos, int64_t C, const double* sin_half,
                                         const double* cos_half, double* scratch, int32_t* conf, double* R, double* t,
                                         void* stream) {
    using namespace tsc;
    if (G <= 0 || C <= 0 || F <= 0) return 0;
    const int64_t n = G * F;
    double* A = scratch, *axis = A + n * 9, *cor = axis + n * 3, *pos = cor + n * 3;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t b1 = (n + 127) / 128;
    if (b1 > 148 * 16) b1 = 148 * 16;
    cyclical_group_kernel<<<(unsigned)b1, 128, 0, st>>>(ref2, tgt2, axis_src, apm, vmean, pmean, n, A, axis, cor, pos);
    TSC_CHECK_LAUNCH();
    int64_t b2 = (G * C * F + 255) / 256;
    if (b2 > 148 * 32) b2 = 148 * 32;
    cyclical_pose_kernel<<<(unsigned)b2, 256, 0, st>>>(A, axis, cor, pos, gconf, combos, sin_half, cos_half, G, F, C, conf, R, t);
    TSC_CHECK_LAUNCH();
    return 0;
}
