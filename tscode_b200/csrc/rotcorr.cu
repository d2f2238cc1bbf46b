// rotcorr.cu — rotor-corrected RMSD for prune_conformers_rmsd_rot_corr.
//
// Reference per pair (tscode/torsion_module.py:953-1011, rotationally_corrected_rmsd):
//   for every "dummy" (locally symmetric) rotor t: try each angle of its n-fold set, rotating the
//   rotor's moving atoms of the second structure about the i2->i3 bond (utils.py:389-414,
//   rotation matrix from algebra.py:284-344), keep the angle with the smallest local Kabsch RMSD
//   over the heavy atoms of the rotor's sub-graph (:982-999, strict <); then apply every best
//   rotation in torsion order (:1004-1008) and return the heavy-atom Kabsch RMSD (:1011; the
//   structures were centred on the all-atom centroid beforehand, :1023; kabsch_rmsd of rmsd==1.4
//   is rotation-only).  Similarity is the single test rmsd < max_rmsd (:1118).
//
// Everything graph-derived (torsion quadruplets, angle sets, rotation masks, sub-graph node
// lists) is pair-independent and arrives precomputed.  The evaluation here is STATELESS (always
// from the centred input): the reference's in-place mutation is reproduced on the host by
// tracking rotor states (tscode_b200/torsion_module.py), which needs the per-pair best-angle
// codes this kernel can emit.
//
// One warp per pair, lanes over atoms, both structures staged in per-warp shared memory.
// FP64-pipe bound (sum_t n_t + 1 eigen-solves per pair); reported as pairs/s.
#include "tsc_common.cuh"
#include "tsc_math.cuh"

namespace tsc {

constexpr int RC_WARPS = 8;
constexpr int RC_MAX_T = 20;                // rotors per molecule: 3 bits of best-angle code each in a 64-bit word
constexpr int RC_MAX_ANG = 6;

struct RotCorrParams {
    const double* Sc;            // (N, A, 3) centred
    int64_t N;
    int A;
    const uint8_t* heavy;        // (A)
    int T;
    const int32_t* i2;           // (T)
    const int32_t* i3;           // (T)
    const int32_t* n_ang;        // (T)
    const double* sin_half;      // (T, 6)
    const double* cos_half;      // (T, 6)
    const uint8_t* rot_mask;     // (T, A)
    const uint8_t* node_mask;    // (T, A)
    int64_t row_begin, row_end;
    int64_t row_stride;          // forward scan: rows row_begin, row_begin + row_stride, ... (row sharding over ranks)
    double max_rmsd;
    uint32_t* sim_bits;          // (N, Wb)
    int64_t Wb;
    uint64_t* codes;             // (N, N) or null: best angle index of rotor t at bits [3t, 3t + 3)
    double* rmsd_out;            // (N, N) or null
    unsigned long long* near_count;
};

// algebra.py:325-344 + :284-323 — rotation about `axis` (not normalised) given sin/cos of half the angle
__device__ __forceinline__ void rot_from_axis(double ax, double ay, double az, double sh, double ch, double R[9]) {
    const double n = sqrt(ax * ax + ay * ay + az * az);
    ax /= n; ay /= n; az /= n;
    const double q1 = sh * ax, q2 = sh * ay, q3 = sh * az, q0 = ch;
    R[0] = 2 * (q0 * q0 + q1 * q1) - 1; R[1] = 2 * (q1 * q2 - q0 * q3);     R[2] = 2 * (q1 * q3 + q0 * q2);
    R[3] = 2 * (q1 * q2 + q0 * q3);     R[4] = 2 * (q0 * q0 + q2 * q2) - 1; R[5] = 2 * (q2 * q3 - q0 * q1);
    R[6] = 2 * (q1 * q3 - q0 * q2);     R[7] = 2 * (q2 * q3 + q0 * q1);     R[8] = 2 * (q0 * q0 + q3 * q3) - 1;
}

__device__ __forceinline__ void rot_point(const double R[9], double cx, double cy, double cz, double& x, double& y,
                                          double& z) {
    const double dx = x - cx, dy = y - cy, dz = z - cz;
    x = (R[0] * dx + R[1] * dy + R[2] * dz) + cx;
    y = (R[3] * dx + R[4] * dy + R[5] * dz) + cy;
    z = (R[6] * dx + R[7] * dy + R[8] * dz) + cz;
}

// One pair, both structures already staged in this warp's shared memory (r* = first structure,
// c* = second structure, which is left rotated to the best rotor angles on return — the
// reference's in-place mutation).  All lanes return the same rmsd / code.
__device__ __forceinline__ double rotcorr_eval(const RotCorrParams& p, const double* rx, const double* ry,
                                               const double* rz, double* cx, double* cy, double* cz, int lane,
                                               uint64_t& code_out) {
    const int A = p.A;
    uint64_t code = 0;
    int best_idx[RC_MAX_T];
    // ---- search phase: every rotor on the unmodified second structure (:982-999) ----
    for (int t = 0; t < p.T; t++) {
        const int a2 = p.i2[t], a3 = p.i3[t];
        const double ox = cx[a3], oy = cy[a3], oz = cz[a3];
        const double ax = cx[a2] - ox, ay = cy[a2] - oy, az = cz[a2] - oz;
        const uint8_t* rm = p.rot_mask + (size_t)t * A;
        const uint8_t* nm = p.node_mask + (size_t)t * A;
        double best = 1e10;
        int bi = 0;
        for (int k = 0; k < p.n_ang[t]; k++) {
            double R[9];
            rot_from_axis(ax, ay, az, p.sin_half[t * RC_MAX_ANG + k], p.cos_half[t * RC_MAX_ANG + k], R);
            double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, G = 0.0, cnt = 0.0;
            for (int a = lane; a < A; a += 32) {
                if (!nm[a]) continue;
                const double px = rx[a], py = ry[a], pz = rz[a];
                double qx = cx[a], qy = cy[a], qz = cz[a];
                if (rm[a]) rot_point(R, ox, oy, oz, qx, qy, qz);
                S[0] = fma(px, qx, S[0]); S[1] = fma(px, qy, S[1]); S[2] = fma(px, qz, S[2]);
                S[3] = fma(py, qx, S[3]); S[4] = fma(py, qy, S[4]); S[5] = fma(py, qz, S[5]);
                S[6] = fma(pz, qx, S[6]); S[7] = fma(pz, qy, S[7]); S[8] = fma(pz, qz, S[8]);
                G += px * px + py * py + pz * pz + qx * qx + qy * qy + qz * qz;
                cnt += 1.0;
            }
#pragma unroll
            for (int c = 0; c < 9; c++) S[c] = warp_sum(S[c]);
            G = warp_sum(G);
            cnt = warp_sum(cnt);
            double q[4];
            const double lam = key_top_eigen(key_matrix(S), q, nullptr);
            const double local = sqrt(fmax(G - 2.0 * lam, 0.0) / cnt);
            if (local < best) { best = local; bi = k; }
        }
        best_idx[t] = bi;
        code |= (uint64_t)bi << (3 * t);
    }
    // ---- apply phase: best rotations in torsion order, each about the CURRENT axis (:1004-1008) ----
    for (int t = 0; t < p.T; t++) {
        const int k = best_idx[t];
        if (k == 0 && p.sin_half[t * RC_MAX_ANG] == 0.0) continue;      // angle 0: identity
        const int a2 = p.i2[t], a3 = p.i3[t];
        const double ox = cx[a3], oy = cy[a3], oz = cz[a3];
        double R[9];
        rot_from_axis(cx[a2] - ox, cy[a2] - oy, cz[a2] - oz, p.sin_half[t * RC_MAX_ANG + k],
                      p.cos_half[t * RC_MAX_ANG + k], R);
        __syncwarp();
        const uint8_t* rm = p.rot_mask + (size_t)t * A;
        for (int a = lane; a < A; a += 32)
            if (rm[a]) {
                double x = cx[a], y = cy[a], z = cz[a];
                rot_point(R, ox, oy, oz, x, y, z);
                cx[a] = x; cy[a] = y; cz[a] = z;
            }
        __syncwarp();
    }
    // ---- global heavy-atom Kabsch RMSD, explicit rotation and differences (:1011) ----
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, cnt = 0.0;
    for (int a = lane; a < A; a += 32) {
        if (!p.heavy[a]) continue;
        const double px = rx[a], py = ry[a], pz = rz[a], qx = cx[a], qy = cy[a], qz = cz[a];
        S[0] = fma(px, qx, S[0]); S[1] = fma(px, qy, S[1]); S[2] = fma(px, qz, S[2]);
        S[3] = fma(py, qx, S[3]); S[4] = fma(py, qy, S[4]); S[5] = fma(py, qz, S[5]);
        S[6] = fma(pz, qx, S[6]); S[7] = fma(pz, qy, S[7]); S[8] = fma(pz, qz, S[8]);
        cnt += 1.0;
    }
#pragma unroll
    for (int c = 0; c < 9; c++) S[c] = warp_sum(S[c]);
    cnt = warp_sum(cnt);
    double R[9];
    kabsch_rot_from_cov(S, R, nullptr, nullptr);
    double ss = 0.0;
    for (int a = lane; a < A; a += 32) {
        if (!p.heavy[a]) continue;
        const double px = rx[a], py = ry[a], pz = rz[a];
        const double dx = fma(R[0], px, fma(R[1], py, R[2] * pz)) - cx[a];
        const double dy = fma(R[3], px, fma(R[4], py, R[5] * pz)) - cy[a];
        const double dz = fma(R[6], px, fma(R[7], py, R[8] * pz)) - cz[a];
        ss += fma(dx, dx, fma(dy, dy, dz * dz));
    }
    ss = warp_sum(ss);
    const double rmsd = sqrt(ss / cnt);
    code_out = code;
    return rmsd;
}

// Lane-parallel form of the search phase.  The (rotor, angle) combinations of a pair are independent, so
// instead of running them one after the other with the 32 lanes splitting the atoms — which makes every
// lane repeat each of the sum_t n_t eigen-solves (~2.5k FP64 instructions each: 94 % of the kernel) — lane c
// takes combination c: it walks all atoms (shared-memory broadcast reads), rotates the moving ones of its
// rotor by its angle, accumulates its own covariance and solves its own eigenproblem.  The best angle per
// rotor is then a warp minimum over the lanes of that rotor (lowest angle index on exact ties, as the
// reference's strict `<` in ascending order, torsion_module.py:994).  Apply phase and global RMSD unchanged.
//   abits[a]: bit t = atom a moves with rotor t (rotation mask), bit 32 + t = atom a belongs to rotor t's
//   heavy-atom sub-graph (staged in shared memory by the caller).
__device__ __forceinline__ double rotcorr_eval2(const RotCorrParams& p, const uint64_t* __restrict__ abits,
                                                const double* rx, const double* ry, const double* rz, double* cx,
                                                double* cy, double* cz, int lane, uint64_t& code_out) {
    const int A = p.A;
    int best_idx[RC_MAX_T];
#pragma unroll
    for (int t = 0; t < RC_MAX_T; t++) best_idx[t] = 0;
    // passes of whole rotors: as many consecutive rotors as fit in 32 lanes (n_ang <= 6, so at least five)
    for (int t_begin = 0; t_begin < p.T;) {
        int t_end = t_begin, ncomb = 0;
        while (t_end < p.T && ncomb + p.n_ang[t_end] <= 32) ncomb += p.n_ang[t_end++];
        int tc = -1, kc = 0;
        {   // lane -> (rotor, angle) of this pass
            int acc = 0;
            for (int t = t_begin; t < t_end; t++) {
                const int n = p.n_ang[t];
                if (tc < 0 && lane < acc + n) { tc = t; kc = lane - acc; }
                acc += n;
            }
        }
        double local = 1e300;
        if (tc >= 0) {
            const int a2 = p.i2[tc], a3 = p.i3[tc];
            const double ox = cx[a3], oy = cy[a3], oz = cz[a3];
            double R[9];
            rot_from_axis(cx[a2] - ox, cy[a2] - oy, cz[a2] - oz, p.sin_half[tc * RC_MAX_ANG + kc],
                          p.cos_half[tc * RC_MAX_ANG + kc], R);
            double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, G = 0.0, cnt = 0.0;
            const uint64_t rbit = 1ull << tc, nbit = 1ull << (32 + tc);
            for (int a = 0; a < A; a++) {
                const uint64_t ab = abits[a];
                if (!(ab & nbit)) continue;
                const double px = rx[a], py = ry[a], pz = rz[a];
                double qx = cx[a], qy = cy[a], qz = cz[a];
                if (ab & rbit) rot_point(R, ox, oy, oz, qx, qy, qz);
                S[0] = fma(px, qx, S[0]); S[1] = fma(px, qy, S[1]); S[2] = fma(px, qz, S[2]);
                S[3] = fma(py, qx, S[3]); S[4] = fma(py, qy, S[4]); S[5] = fma(py, qz, S[5]);
                S[6] = fma(pz, qx, S[6]); S[7] = fma(pz, qy, S[7]); S[8] = fma(pz, qz, S[8]);
                G += px * px + py * py + pz * pz + qx * qx + qy * qy + qz * qz;
                cnt += 1.0;
            }
            double q[4];
            const double lam = key_top_eigen(key_matrix(S), q, nullptr);
            local = sqrt(fmax(G - 2.0 * lam, 0.0) / cnt);
        }
        __syncwarp();
        // per rotor: minimum over its lanes, lowest angle index on ties
        for (int t = t_begin; t < t_end; t++) {
            const double v = (tc == t) ? local : 1e300;
            double m = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
            const uint32_t who = __ballot_sync(0xffffffffu, tc == t && v == m);
            const int src = who ? __ffs(who) - 1 : 0;
            const int kb = __shfl_sync(0xffffffffu, kc, src);
            if (who) best_idx[t] = kb;
        }
        t_begin = t_end;
    }
    uint64_t code = 0;
    for (int t = 0; t < p.T; t++) code |= (uint64_t)best_idx[t] << (3 * t);
    // ---- apply phase: best rotations in torsion order, each about the CURRENT axis (:1004-1008) ----
    for (int t = 0; t < p.T; t++) {
        const int k = best_idx[t];
        if (k == 0 && p.sin_half[t * RC_MAX_ANG] == 0.0) continue;      // angle 0: identity
        const int a2 = p.i2[t], a3 = p.i3[t];
        const double ox = cx[a3], oy = cy[a3], oz = cz[a3];
        double R[9];
        rot_from_axis(cx[a2] - ox, cy[a2] - oy, cz[a2] - oz, p.sin_half[t * RC_MAX_ANG + k],
                      p.cos_half[t * RC_MAX_ANG + k], R);
        __syncwarp();
        for (int a = lane; a < A; a += 32)
            if (abits[a] & (1ull << t)) {
                double x = cx[a], y = cy[a], z = cz[a];
                rot_point(R, ox, oy, oz, x, y, z);
                cx[a] = x; cy[a] = y; cz[a] = z;
            }
        __syncwarp();
    }
    // ---- global heavy-atom Kabsch RMSD, explicit rotation and differences (:1011) ----
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, cnt = 0.0;
    for (int a = lane; a < A; a += 32) {
        if (!p.heavy[a]) continue;
        const double px = rx[a], py = ry[a], pz = rz[a], qx = cx[a], qy = cy[a], qz = cz[a];
        S[0] = fma(px, qx, S[0]); S[1] = fma(px, qy, S[1]); S[2] = fma(px, qz, S[2]);
        S[3] = fma(py, qx, S[3]); S[4] = fma(py, qy, S[4]); S[5] = fma(py, qz, S[5]);
        S[6] = fma(pz, qx, S[6]); S[7] = fma(pz, qy, S[7]); S[8] = fma(pz, qz, S[8]);
        cnt += 1.0;
    }
#pragma unroll
    for (int c = 0; c < 9; c++) S[c] = warp_sum(S[c]);
    cnt = warp_sum(cnt);
    double R[9];
    kabsch_rot_from_cov(S, R, nullptr, nullptr);
    double ss = 0.0;
    for (int a = lane; a < A; a += 32) {
        if (!p.heavy[a]) continue;
        const double px = rx[a], py = ry[a], pz = rz[a];
        const double dx = fma(R[0], px, fma(R[1], py, R[2] * pz)) - cx[a];
        const double dy = fma(R[3], px, fma(R[4], py, R[5] * pz)) - cy[a];
        const double dz = fma(R[6], px, fma(R[7], py, R[8] * pz)) - cz[a];
        ss += fma(dx, dx, fma(dy, dy, dz * dz));
    }
    ss = warp_sum(ss);
    code_out = code;
    return sqrt(ss / cnt);
}

// rotation / sub-graph masks of every atom as one word (see rotcorr_eval2), into shared memory
__device__ __forceinline__ void stage_abits(const RotCorrParams& p, uint64_t* abits) {
    for (int a = threadIdx.x; a < p.A; a += blockDim.x) {
        uint64_t w = 0;
        for (int t = 0; t < p.T; t++) {
            if (p.rot_mask[(size_t)t * p.A + a]) w |= 1ull << t;
            if (p.node_mask[(size_t)t * p.A + a]) w |= 1ull << (32 + t);
        }
        abits[a] = w;
    }
}

// ------------------------------------------------------------------------------------------
// Forward scan (stateless mode, large N).  The grouping loop (torsion_module.py:1098-1125) walks, for
// every first structure i, the later structures j in order and stops at the first similar one; pairs it
// found dissimilar are cached and never looked at again.  With a stateless pair function the only facts
// the loop ever uses about row i are therefore  first_hit[i] = min{ j > i : rmsd(i, j) < max_rmsd }  and
// the best-angle codes of the pairs (i, j <= first_hit[i]) it mutates on the way.  One CTA per row
// (rows handed out through an atomic counter), its 8 warps evaluating 8 consecutive j at a time until a
// batch contains a hit: a redundant ensemble needs a few dozen pairs per row instead of N - i.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RC_WARPS * 32) rotcorr_scan_kernel(const RotCorrParams p, int32_t* __restrict__ first_hit,
                                                                     int32_t* __restrict__ row_counter) {
    extern __shared__ double smem[];
    __shared__ int s_row, s_hit;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int A = p.A;
    double* rx = smem;                                      // first structure, shared by the warps
    double* ry = rx + A; double* rz = ry + A;
    double* cx = rz + A + (size_t)warp * 3 * A;             // second structure, per warp
    double* cy = cx + A; double* cz = cy + A;
    uint64_t* abits = reinterpret_cast<uint64_t*>(smem + (size_t)(3 + 3 * RC_WARPS) * A);
    stage_abits(p, abits);
    unsigned long long near = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { s_row = atomicAdd(row_counter, 1); s_hit = 0x7fffffff; }
        __syncthreads();
        const int64_t i = p.row_begin + (int64_t)s_row * p.row_stride;
        if (i >= p.row_end) break;
        const double* Pi = p.Sc + i * (int64_t)A * 3;
        for (int a = threadIdx.x; a < A; a += blockDim.x) { rx[a] = Pi[3 * a]; ry[a] = Pi[3 * a + 1]; rz[a] = Pi[3 * a + 2]; }
        __syncthreads();
        for (int64_t j0 = i + 1; j0 < p.N; j0 += RC_WARPS) {
            const int64_t j = j0 + warp;
            if (j < p.N) {
                const double* Pj = p.Sc + j * (int64_t)A * 3;
                for (int a = lane; a < A; a += 32) { cx[a] = Pj[3 * a]; cy[a] = Pj[3 * a + 1]; cz[a] = Pj[3 * a + 2]; }
                __syncwarp();
                uint64_t code;
                const double rmsd = rotcorr_eval2(p, abits, rx, ry, rz, cx, cy, cz, lane, code);
                if (lane == 0) {
                    if (p.codes) p.codes[i * p.N + j] = code;
                    if (p.rmsd_out) p.rmsd_out[i * p.N + j] = rmsd;
                    if (rmsd < p.max_rmsd) atomicMin(&s_hit, (int)j);
                    near += fabs(rmsd - p.max_rmsd) < 1e-6;
                }
                __syncwarp();
            }
            __syncthreads();
            if (s_hit != 0x7fffffff) break;
        }
        if (threadIdx.x == 0) first_hit[i] = (s_hit == 0x7fffffff) ? (int32_t)p.N : s_hit;
    }
    if (lane == 0 && near && p.near_count) atomicAdd(p.near_count, near);
}

__global__ void __launch_bounds__(RC_WARPS * 32) rotcorr_pairs_kernel(const RotCorrParams p) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int A = p.A;
    double* rx = smem + (size_t)warp * 6 * A;
    double* ry = rx + A; double* rz = ry + A;
    double* cx = rz + A; double* cy = cx + A; double* cz = cy + A;
    uint64_t* abits = reinterpret_cast<uint64_t*>(smem + (size_t)RC_WARPS * 6 * A);
    stage_abits(p, abits);
    __syncthreads();
    const int64_t warp_g = (int64_t)blockIdx.x * RC_WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * RC_WARPS;
    const int64_t nrows = p.row_end - p.row_begin;
    unsigned long long near = 0;
    for (int64_t idx = warp_g; idx < nrows * p.N; idx += nwarps) {
        const int64_t i = p.row_begin + idx / p.N, j = idx % p.N;
        if (j <= i) continue;
        const double* Pi = p.Sc + i * (int64_t)A * 3;
        const double* Pj = p.Sc + j * (int64_t)A * 3;
        __syncwarp();
        for (int a = lane; a < A; a += 32) {
            rx[a] = Pi[3 * a]; ry[a] = Pi[3 * a + 1]; rz[a] = Pi[3 * a + 2];
            cx[a] = Pj[3 * a]; cy[a] = Pj[3 * a + 1]; cz[a] = Pj[3 * a + 2];
        }
        __syncwarp();
        uint64_t code;
        const double rmsd = rotcorr_eval2(p, abits, rx, ry, rz, cx, cy, cz, lane, code);
        if (lane == 0) {
            if (rmsd < p.max_rmsd) atomicOr(&p.sim_bits[i * p.Wb + (j >> 5)], 1u << (j & 31));
            if (p.codes) p.codes[i * p.N + j] = code;
            if (p.rmsd_out) p.rmsd_out[i * p.N + j] = rmsd;
            near += fabs(rmsd - p.max_rmsd) < 1e-6;
        }
    }
    if (lane == 0 && near && p.near_count) atomicAdd(p.near_count, near);
}


// Stateful mode (exact emulation of the reference's in-place mutation): one row of the grouping
// loop.  cur (N, A, 3) holds the CURRENT (mutated) structures; for the first structure i and a
// list of second structures js[k] this evaluates every pair from the current coordinates and
// stages the mutated copy of js[k]; the host finds the first similar k and commits the staged
// copies of everything visited (tsc_rotcorr_commit).
__global__ void __launch_bounds__(RC_WARPS * 32) rotcorr_row_kernel(const RotCorrParams p, int64_t i,
                                                                    const int32_t* __restrict__ js, int n,
                                                                    double* __restrict__ rmsd, uint64_t* __restrict__ codes,
                                                                    double* __restrict__ staged) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int A = p.A;
    double* rx = smem + (size_t)warp * 6 * A;
    double* ry = rx + A; double* rz = ry + A;
    double* cx = rz + A; double* cy = cx + A; double* cz = cy + A;
    const double* Pi = p.Sc + i * (int64_t)A * 3;
    for (int k = blockIdx.x * RC_WARPS + warp; k < n; k += gridDim.x * RC_WARPS) {
        const double* Pj = p.Sc + (int64_t)js[k] * A * 3;
        __syncwarp();
        for (int a = lane; a < A; a += 32) {
            rx[a] = Pi[3 * a]; ry[a] = Pi[3 * a + 1]; rz[a] = Pi[3 * a + 2];
            cx[a] = Pj[3 * a]; cy[a] = Pj[3 * a + 1]; cz[a] = Pj[3 * a + 2];
        }
        __syncwarp();
        uint64_t code;
        const double r = rotcorr_eval(p, rx, ry, rz, cx, cy, cz, lane, code);
        __syncwarp();
        double* O = staged + (int64_t)k * A * 3;
        for (int a = lane; a < A; a += 32) { O[3 * a] = cx[a]; O[3 * a + 1] = cy[a]; O[3 * a + 2] = cz[a]; }
        if (lane == 0) { rmsd[k] = r; codes[k] = code; }
    }
}

__global__ void rotcorr_commit_kernel(double* __restrict__ cur, const double* __restrict__ staged,
                                      const int32_t* __restrict__ js, int n_accept, int A3) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)n_accept * A3;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e / A3), o = (int)(e % A3);
        cur[(int64_t)js[k] * A3 + o] = staged[e];
    }
}

// Returned structures: rotor t of structure idx[s] rotated by its accumulated state angle
// (sin/cos of half the angle per (s, t)), in torsion order, each about the current axis.
__global__ void __launch_bounds__(RC_WARPS * 32) rotcorr_apply_kernel(
    const double* __restrict__ Sc, int64_t n, int A, const int64_t* __restrict__ idx, int T,
    const int32_t* __restrict__ i2, const int32_t* __restrict__ i3, const double* __restrict__ sin_half,
    const double* __restrict__ cos_half, const uint8_t* __restrict__ rot_mask, double* __restrict__ out) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* cx = smem + (size_t)warp * 3 * A;
    double* cy = cx + A; double* cz = cy + A;
    for (int64_t s = (int64_t)blockIdx.x * RC_WARPS + warp; s < n; s += (int64_t)gridDim.x * RC_WARPS) {
        const double* P = Sc + idx[s] * (int64_t)A * 3;
        __syncwarp();
        for (int a = lane; a < A; a += 32) { cx[a] = P[3 * a]; cy[a] = P[3 * a + 1]; cz[a] = P[3 * a + 2]; }
        __syncwarp();
        for (int t = 0; t < T; t++) {
            const double sh = sin_half[s * T + t], ch = cos_half[s * T + t];
            if (sh == 0.0 && ch == 1.0) continue;
            const int a2 = i2[t], a3 = i3[t];
            const double ox = cx[a3], oy = cy[a3], oz = cz[a3];
            double R[9];
            rot_from_axis(cx[a2] - ox, cy[a2] - oy, cz[a2] - oz, sh, ch, R);
            __syncwarp();
            const uint8_t* rm = rot_mask + (size_t)t * A;
            for (int a = lane; a < A; a += 32)
                if (rm[a]) {
                    double x = cx[a], y = cy[a], z = cz[a];
                    rot_point(R, ox, oy, oz, x, y, z);
                    cx[a] = x; cy[a] = y; cz[a] = z;
                }
            __syncwarp();
        }
        double* O = out + s * (int64_t)A * 3;
        for (int a = lane; a < A; a += 32) { O[3 * a] = cx[a]; O[3 * a + 1] = cy[a]; O[3 * a + 2] = cz[a]; }
    }
}

}  // namespace tsc

extern "C" int tsc_rotcorr_pairs(const double* Sc, int64_t N, int32_t A, const uint8_t* heavy, int32_t T,
                                 const int32_t* tor_i2, const int32_t* tor_i3, const int32_t* n_ang,
                                 const double* sin_half, const double* cos_half, const uint8_t* rot_mask,
                                 const uint8_t* node_mask, int64_t row_begin, int64_t row_end, double max_rmsd,
                                 uint32_t* sim_bits, uint64_t* codes, double* rmsd_out, uint64_t* near_count,
                                 void* stream) {
    using namespace tsc;
    if (N <= 1 || row_end <= row_begin) return 0;
    if (T < 0 || T > RC_MAX_T) return (int)cudaErrorInvalidValue;
    RotCorrParams p{Sc, N, A, heavy, T, tor_i2, tor_i3, n_ang, sin_half, cos_half, rot_mask, node_mask, row_begin,
                    row_end, 1, max_rmsd, sim_bits, (N + 31) / 32, codes, rmsd_out,
                    reinterpret_cast<unsigned long long*>(near_count)};
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sim_bits + row_begin * p.Wb, 0, (size_t)(row_end - row_begin) * p.Wb * 4, st);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)RC_WARPS * 6 * A * sizeof(double) + (size_t)A * 8;
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    e = cudaFuncSetAttribute(rotcorr_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int64_t work = (row_end - row_begin) * N;
    int64_t blocks = (work + RC_WARPS - 1) / RC_WARPS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    rotcorr_pairs_kernel<<<(unsigned)blocks, RC_WARPS * 32, smem, st>>>(p);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_rotcorr_apply(const double* Sc, int64_t n, int32_t A, const int64_t* idx, int32_t T,
                                 const int32_t* tor_i2, const int32_t* tor_i3, const double* sin_half,
                                 const double* cos_half, const uint8_t* rot_mask, double* out, void* stream) {
    using namespace tsc;
    if (n <= 0) return 0;
    const size_t smem = (size_t)RC_WARPS * 3 * A * sizeof(double);
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(rotcorr_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int64_t blocks = (n + RC_WARPS - 1) / RC_WARPS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    rotcorr_apply_kernel<<<(unsigned)blocks, RC_WARPS * 32, smem, (cudaStream_t)stream>>>(
        Sc, n, A, idx, T, tor_i2, tor_i3, sin_half, cos_half, rot_mask, out);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_rotcorr_row(const double* cur, int64_t N, int32_t A, const uint8_t* heavy, int32_t T,
                               const int32_t* tor_i2, const int32_t* tor_i3, const int32_t* n_ang,
                               const double* sin_half, const double* cos_half, const uint8_t* rot_mask,
                               const uint8_t* node_mask, int64_t i, const int32_t* js, int32_t n, double* rmsd,
                               uint64_t* codes, double* staged, void* stream) {
    using namespace tsc;
    if (n <= 0) return 0;
    if (T < 0 || T > RC_MAX_T) return (int)cudaErrorInvalidValue;
    RotCorrParams p{cur, N, A, heavy, T, tor_i2, tor_i3, n_ang, sin_half, cos_half, rot_mask, node_mask, 0, 0, 1,
                    0.0, nullptr, 0, nullptr, nullptr, nullptr};
    const size_t smem = (size_t)RC_WARPS * 6 * A * sizeof(double);
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(rotcorr_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int blocks = (n + RC_WARPS - 1) / RC_WARPS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    rotcorr_row_kernel<<<blocks, RC_WARPS * 32, smem, (cudaStream_t)stream>>>(p, i, js, n, rmsd, codes, staged);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_rotcorr_commit(double* cur, const double* staged, const int32_t* js, int32_t n_accept, int32_t A,
                                  void* stream) {
    if (n_accept <= 0) return 0;
    const int64_t total = (int64_t)n_accept * A * 3;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    tsc::rotcorr_commit_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cur, staged, js, n_accept, A * 3);
    TSC_CHECK_LAUNCH();
    return 0;
}

// Forward scan of rows row_begin, row_begin + row_stride, ... < row_end (row_stride = number of ranks when the rows are
// dealt round-robin): first_hit[i] = first j > i with rmsd(i, j) < max_rmsd (N if none);
// codes / rmsd_out (dense (N, N), may be NULL) are written for the pairs evaluated on the way (at least every
// j <= first_hit[i]).  row_counter: one int32 of scratch.
extern "C" int tsc_rotcorr_scan(const double* Sc, int64_t N, int32_t A, const uint8_t* heavy, int32_t T,
                                const int32_t* tor_i2, const int32_t* tor_i3, const int32_t* n_ang,
                                const double* sin_half, const double* cos_half, const uint8_t* rot_mask,
                                const uint8_t* node_mask, int64_t row_begin, int64_t row_end, int64_t row_stride,
                                double max_rmsd, int32_t* first_hit, uint64_t* codes, double* rmsd_out,
                                uint64_t* near_count, int32_t* row_counter, void* stream) {
    using namespace tsc;
    if (N <= 0 || row_end <= row_begin) return 0;
    if (T < 0 || T > RC_MAX_T || row_stride < 1) return (int)cudaErrorInvalidValue;
    RotCorrParams p{Sc, N, A, heavy, T, tor_i2, tor_i3, n_ang, sin_half, cos_half, rot_mask, node_mask, row_begin,
                    row_end, row_stride, max_rmsd, nullptr, 0, codes, rmsd_out,
                    reinterpret_cast<unsigned long long*>(near_count)};
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(row_counter, 0, 4, st);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)(3 + 3 * RC_WARPS) * A * sizeof(double) + (size_t)A * 8;
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    e = cudaFuncSetAttribute(rotcorr_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int64_t blocks = (row_end - row_begin + row_stride - 1) / row_stride;
    if (blocks > 148 * 4) blocks = 148 * 4;
    rotcorr_scan_kernel<<<(unsigned)blocks, RC_WARPS * 32, smem, st>>>(p, first_hit, row_counter);
    TSC_CHECK_LAUNCH();
    return 0;
}
