// rmsd_verify.cu — exact re-evaluation of screened pairs, and the batched rmsd_and_max entry.
//
// For every bit the screen (rmsd_sim.cu) left set, a warp redoes what the reference does for
// that pair (tscode/rmsd_pruning.py:6-41): cross-covariance, optimal proper rotation (Horn key
// matrix + Jacobi instead of LAPACK gesdd + reflection fix — same rotation whenever it is
// unique), p rotated explicitly, explicit differences, rmsd = sqrt(sum/M), max deviation = max
// row norm; the bit survives iff  rmsd < thr  and  maxdev < 2*thr  (:75, :95; both strict).
// This is also what gives RMSD values their 1e-9 A accuracy near zero, where the closed form
// G - 2*lambda cancels.
//
// Warps scan the similarity rows they own, batch up to 32 candidates, and evaluate a batch with
// lanes splitting the atoms of each pair for the two passes over the coordinates while the
// eigen-solves of the whole batch run one per lane (round-1 first version solved every
// candidate redundantly on all 32 lanes: 2.6 ms for C3's 250 k candidates).
#include "tsc_common.cuh"
#include "tsc_math.cuh"

namespace tsc {

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct PairEval {
    double rmsd, maxdev, gap, lam;
};

// p/q accessors differ between the packed layout and plain AoS arrays
struct PackedView {
    const double* base;
    int64_t nb_pad;
    __device__ __forceinline__ void load(int64_t i, int m, double& x, double& y, double& z) const {
        const int64_t o = packed_index(i, m, 0, nb_pad);
        x = base[o]; y = base[o + CB * KS]; z = base[o + 2 * CB * KS];
    }
};
struct AosView {
    const double* base;     // (n, M, 3)
    int M;
    __device__ __forceinline__ void load(int64_t i, int m, double& x, double& y, double& z) const {
        const double* a = base + (i * M + m) * 3;
        x = a[0]; y = a[1]; z = a[2];
    }
};

// Warp-cooperative rmsd_and_max.  All lanes return the same values.
template <class VP, class VQ>
__device__ __forceinline__ PairEval eval_pair(const VP& P, int64_t i, const VQ& Q, int64_t j, int M, int lane) {
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int m = lane; m < M; m += 32) {
        double px, py, pz, qx, qy, qz;
        P.load(i, m, px, py, pz);
        Q.load(j, m, qx, qy, qz);
        S[0] = fma(px, qx, S[0]); S[1] = fma(px, qy, S[1]); S[2] = fma(px, qz, S[2]);
        S[3] = fma(py, qx, S[3]); S[4] = fma(py, qy, S[4]); S[5] = fma(py, qz, S[5]);
        S[6] = fma(pz, qx, S[6]); S[7] = fma(pz, qy, S[7]); S[8] = fma(pz, qz, S[8]);
    }
#pragma unroll
    for (int c = 0; c < 9; c++) S[c] = warp_sum(S[c]);
    double R[9];
    PairEval ev;
    kabsch_rot_from_cov(S, R, &ev.lam, &ev.gap);
    double ss = 0.0, mx = 0.0;
    for (int m = lane; m < M; m += 32) {
        double px, py, pz, qx, qy, qz;
        P.load(i, m, px, py, pz);
        Q.load(j, m, qx, qy, qz);
        const double dx = fma(R[0], px, fma(R[1], py, R[2] * pz)) - qx;
        const double dy = fma(R[3], px, fma(R[4], py, R[5] * pz)) - qy;
        const double dz = fma(R[6], px, fma(R[7], py, R[8] * pz)) - qz;
        const double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
        ss += d2;
        mx = fmax(mx, d2);
    }
    ss = warp_sum(ss);
    mx = warp_max(mx);
    ev.rmsd = sqrt(ss / (double)M);
    ev.maxdev = sqrt(mx);
    return ev;
}

// Warp-cooperative pieces of eval_pair, split so that the (lane-redundant, ~2.5k FP64 instruction)
// eigen-solve of up to 32 candidates can run one per lane instead of 32 times per candidate.
template <class V>
__device__ __forceinline__ void pair_cov(const V& P, int64_t i, int64_t j, int M, int lane, double S[9]) {
#pragma unroll
    for (int c = 0; c < 9; c++) S[c] = 0.0;
    for (int m = lane; m < M; m += 32) {
        double px, py, pz, qx, qy, qz;
        P.load(i, m, px, py, pz);
        P.load(j, m, qx, qy, qz);
        S[0] = fma(px, qx, S[0]); S[1] = fma(px, qy, S[1]); S[2] = fma(px, qz, S[2]);
        S[3] = fma(py, qx, S[3]); S[4] = fma(py, qy, S[4]); S[5] = fma(py, qz, S[5]);
        S[6] = fma(pz, qx, S[6]); S[7] = fma(pz, qy, S[7]); S[8] = fma(pz, qz, S[8]);
    }
#pragma unroll
    for (int c = 0; c < 9; c++) S[c] = warp_sum(S[c]);
}
template <class V>
__device__ __forceinline__ void pair_diff(const V& P, int64_t i, int64_t j, int M, int lane, const double R[9],
                                          double& rmsd, double& maxdev) {
    double ss = 0.0, mx = 0.0;
    for (int m = lane; m < M; m += 32) {
        double px, py, pz, qx, qy, qz;
        P.load(i, m, px, py, pz);
        P.load(j, m, qx, qy, qz);
        const double dx = fma(R[0], px, fma(R[1], py, R[2] * pz)) - qx;
        const double dy = fma(R[3], px, fma(R[4], py, R[5] * pz)) - qy;
        const double dz = fma(R[6], px, fma(R[7], py, R[8] * pz)) - qz;
        const double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
        ss += d2;
        mx = fmax(mx, d2);
    }
    ss = warp_sum(ss);
    mx = warp_max(mx);
    rmsd = sqrt(ss / (double)M);
    maxdev = sqrt(mx);
}

// stats: [0] candidates examined, [1] confirmed similar, [2] within 1e-6 A of a threshold,
//        [3] degenerate top eigenvalue (rotation not unique)
// Each warp scans the similarity rows it owns and batches up to 32 candidates (row, j) in shared
// memory; a batch is evaluated in three phases: cooperative covariances (lane k keeps the one of
// candidate k), one eigen-solve per lane, cooperative explicit differences with the rotation
// broadcast from lane k.  Bits that fail are cleared with atomicAnd.
constexpr int VF_WARPS = 8;
__global__ void __launch_bounds__(VF_WARPS * 32) rmsd_verify_kernel(const double* __restrict__ packed, int64_t N, int M,
                                                                   int64_t nb_pad, const int32_t* __restrict__ row_blocks,
                                                                   int n_rb, double thr, uint32_t* sim_bits, int64_t W,
                                                                   unsigned long long* stats, int2* pair_list,
                                                                   int64_t pair_stride, const int2* cand,
                                                                   int64_t cand_stride) {
    // a valid candidate list means rmsd_verify_list_kernel has done the work already
    if (cand && cand[0].x >= 0 && (int64_t)cand[0].x <= cand_stride - 1) return;
    __shared__ int32_t s_row[VF_WARPS][32], s_i[VF_WARPS][32], s_j[VF_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t warp_g = (int64_t)blockIdx.x * VF_WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * VF_WARPS;
    const PackedView V{packed, nb_pad};
    const double thr2 = 2.0 * thr;
    unsigned long long n_cand = 0, n_ok = 0, n_near = 0, n_deg = 0;
    int count = 0;

    auto flush = [&]() {
        if (count == 0) return;
        __syncwarp();
        double Sk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < count; k++) {                       // phase A
            double S[9];
            pair_cov(V, s_i[warp][k], s_j[warp][k], M, lane, S);
            if (lane == k) {
#pragma unroll
                for (int c = 0; c < 9; c++) Sk[c] = S[c];
            }
        }
        double Rk[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, lam = 1.0, gap = 1.0;
        if (lane < count) kabsch_rot_from_cov(Sk, Rk, &lam, &gap);          // phase B: one solve per lane
        __syncwarp();
        uint32_t okmask = 0;
        for (int k = 0; k < count; k++) {                       // phase C
            double R[9];
#pragma unroll
            for (int c = 0; c < 9; c++) R[c] = __shfl_sync(0xffffffffu, Rk[c], k);
            const double lk = __shfl_sync(0xffffffffu, lam, k), gk = __shfl_sync(0xffffffffu, gap, k);
            double rmsd, maxdev;
            const int64_t i = s_i[warp][k], j = s_j[warp][k];
            pair_diff(V, i, j, M, lane, R, rmsd, maxdev);
            const bool ok = (rmsd < thr) && (maxdev < thr2);
            n_cand++;
            n_ok += ok;
            n_near += (fabs(rmsd - thr) < 1e-6) || ((rmsd < thr) && fabs(maxdev - thr2) < 1e-6);
            n_deg += ok && (gk < 1e-9 * fabs(lk));
            if (!ok && lane == 0) atomicAnd(&sim_bits[(int64_t)s_row[warp][k] * W + (j >> 5)], ~(1u << (j & 31)));
            okmask |= (ok ? 1u : 0u) << k;
        }
        if (pair_list && okmask) {           // confirmed pairs of the batch -> (i, j) list, one atomic per batch
            int64_t base = 0;
            if (lane == 0) base = list_reserve(&pair_list[0].x, __popc(okmask), pair_stride - 1);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((okmask >> lane) & 1u) {
                const int64_t slot = base + __popc(okmask & ((1u << lane) - 1u));
                if (slot < pair_stride - 1) pair_list[1 + slot] = make_int2(s_i[warp][lane], s_j[warp][lane]);
            }
        }
        count = 0;
        __syncwarp();
    };

    for (int64_t row = warp_g; row < (int64_t)n_rb * CB; row += nwarps) {
        const int64_t ib = row_blocks[row / CB];
        const int64_t i = ib * CB + (row % CB);
        if (i >= N) continue;
        const uint32_t* rw = sim_bits + row * W;
        for (int64_t w0 = ib; w0 < W; w0 += 32) {
            const int64_t w = w0 + lane;
            const uint32_t word = (w < W) ? rw[w] : 0u;
            uint32_t pending = __ballot_sync(0xffffffffu, word != 0u);
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1;
                uint32_t bits = __shfl_sync(0xffffffffu, word, src);
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (lane == 0) {
                        s_row[warp][count] = (int32_t)row;
                        s_i[warp][count] = (int32_t)i;
                        s_j[warp][count] = (int32_t)((w0 + src) * 32 + b);
                    }
                    if (++count == 32) flush();
                }
            }
        }
    }
    flush();
    if (lane == 0 && stats) {
        if (n_cand) atomicAdd(&stats[0], n_cand);
        if (n_ok) atomicAdd(&stats[1], n_ok);
        if (n_near) atomicAdd(&stats[2], n_near);
        if (n_deg) atomicAdd(&stats[3], n_deg);
    }
}

// ------------------------------------------------------------------------------------------
// Verification from a CANDIDATE LIST (the tcgen05 screens append (local row, j) of every bit they
// set).  The bit-row scan above costs a dependent L2 round trip per 32 words of a 1564-word row and
// leaves most lanes idle during the eigen-solves (a row has ~5 candidates); here a warp takes 32
// list entries at a time: covariances and explicit differences are computed 8 candidates per pass by
// groups of 4 lanes (one 20-atom slab of the packed layout per lane and step, 16-byte loads of
// contiguous runs), the 32 eigen-solves run one per lane.  Same arithmetic per pair as eval_pair up
// to the order of the atom sums.  Runs only if the list did not overflow (header count <= capacity);
// otherwise rmsd_verify_kernel does the work from the bit rows.
// ------------------------------------------------------------------------------------------
// FULL: a whole 20-atom slab with the loop unrolled, so that the 60 (cov) / 120 (diff) loads of the slab are
// independent instructions the scheduler can keep in flight together (the rolled loop paid one L2 round trip
// per two atoms).
template <bool FULL>
__device__ __forceinline__ void slab_cov(const double* __restrict__ P, const double* __restrict__ Q, int n, double S[9]) {
    // P, Q: x run of this conformer's slab; y at + CB*KS, z at + 2*CB*KS
    const int nn = FULL ? KS : n;
#pragma unroll
    for (int k = 0; k < (FULL ? KS : 1); k += 2) {
        if (!FULL) {
#pragma unroll 1
            for (int kk = 0; kk < nn; kk += 2) {
                const double2 px = *reinterpret_cast<const double2*>(P + kk), py = *reinterpret_cast<const double2*>(P + CB * KS + kk),
                              pz = *reinterpret_cast<const double2*>(P + 2 * CB * KS + kk);
                const double2 qx = *reinterpret_cast<const double2*>(Q + kk), qy = *reinterpret_cast<const double2*>(Q + CB * KS + kk),
                              qz = *reinterpret_cast<const double2*>(Q + 2 * CB * KS + kk);
                S[0] = fma(px.x, qx.x, S[0]); S[1] = fma(px.x, qy.x, S[1]); S[2] = fma(px.x, qz.x, S[2]);
                S[3] = fma(py.x, qx.x, S[3]); S[4] = fma(py.x, qy.x, S[4]); S[5] = fma(py.x, qz.x, S[5]);
                S[6] = fma(pz.x, qx.x, S[6]); S[7] = fma(pz.x, qy.x, S[7]); S[8] = fma(pz.x, qz.x, S[8]);
                S[0] = fma(px.y, qx.y, S[0]); S[1] = fma(px.y, qy.y, S[1]); S[2] = fma(px.y, qz.y, S[2]);
                S[3] = fma(py.y, qx.y, S[3]); S[4] = fma(py.y, qy.y, S[4]); S[5] = fma(py.y, qz.y, S[5]);
                S[6] = fma(pz.y, qx.y, S[6]); S[7] = fma(pz.y, qy.y, S[7]); S[8] = fma(pz.y, qz.y, S[8]);
            }
        } else {
            const double2 px = *reinterpret_cast<const double2*>(P + k), py = *reinterpret_cast<const double2*>(P + CB * KS + k),
                          pz = *reinterpret_cast<const double2*>(P + 2 * CB * KS + k);
            const double2 qx = *reinterpret_cast<const double2*>(Q + k), qy = *reinterpret_cast<const double2*>(Q + CB * KS + k),
                          qz = *reinterpret_cast<const double2*>(Q + 2 * CB * KS + k);
            S[0] = fma(px.x, qx.x, S[0]); S[1] = fma(px.x, qy.x, S[1]); S[2] = fma(px.x, qz.x, S[2]);
            S[3] = fma(py.x, qx.x, S[3]); S[4] = fma(py.x, qy.x, S[4]); S[5] = fma(py.x, qz.x, S[5]);
            S[6] = fma(pz.x, qx.x, S[6]); S[7] = fma(pz.x, qy.x, S[7]); S[8] = fma(pz.x, qz.x, S[8]);
            S[0] = fma(px.y, qx.y, S[0]); S[1] = fma(px.y, qy.y, S[1]); S[2] = fma(px.y, qz.y, S[2]);
            S[3] = fma(py.y, qx.y, S[3]); S[4] = fma(py.y, qy.y, S[4]); S[5] = fma(py.y, qz.y, S[5]);
            S[6] = fma(pz.y, qx.y, S[6]); S[7] = fma(pz.y, qy.y, S[7]); S[8] = fma(pz.y, qz.y, S[8]);
        }
    }
}
template <bool FULL>
__device__ __forceinline__ void slab_diff(const double* __restrict__ P, const double* __restrict__ Q, int n,
                                          const double R[9], double& ss, double& mx) {
    if (FULL) {
#pragma unroll
        for (int k = 0; k < KS; k += 2) {
            const double2 px = *reinterpret_cast<const double2*>(P + k), py = *reinterpret_cast<const double2*>(P + CB * KS + k),
                          pz = *reinterpret_cast<const double2*>(P + 2 * CB * KS + k);
            const double2 qx = *reinterpret_cast<const double2*>(Q + k), qy = *reinterpret_cast<const double2*>(Q + CB * KS + k),
                          qz = *reinterpret_cast<const double2*>(Q + 2 * CB * KS + k);
            {
                const double dx = fma(R[0], px.x, fma(R[1], py.x, R[2] * pz.x)) - qx.x;
                const double dy = fma(R[3], px.x, fma(R[4], py.x, R[5] * pz.x)) - qy.x;
                const double dz = fma(R[6], px.x, fma(R[7], py.x, R[8] * pz.x)) - qz.x;
                const double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
                ss += d2; mx = fmax(mx, d2);
            }
            {
                const double dx = fma(R[0], px.y, fma(R[1], py.y, R[2] * pz.y)) - qx.y;
                const double dy = fma(R[3], px.y, fma(R[4], py.y, R[5] * pz.y)) - qy.y;
                const double dz = fma(R[6], px.y, fma(R[7], py.y, R[8] * pz.y)) - qz.y;
                const double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
                ss += d2; mx = fmax(mx, d2);
            }
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < n; k++) {
            const double px = P[k], py = P[CB * KS + k], pz = P[2 * CB * KS + k];
            const double dx = fma(R[0], px, fma(R[1], py, R[2] * pz)) - Q[k];
            const double dy = fma(R[3], px, fma(R[4], py, R[5] * pz)) - Q[CB * KS + k];
            const double dz = fma(R[6], px, fma(R[7], py, R[8] * pz)) - Q[2 * CB * KS + k];
            const double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
            ss += d2;
            mx = fmax(mx, d2);
        }
    }
}

__global__ void __launch_bounds__(VF_WARPS * 32) rmsd_verify_list_kernel(
    const double* __restrict__ packed, int64_t N, int M, int64_t nb_pad, const int32_t* __restrict__ row_blocks,
    double thr, uint32_t* sim_bits, int64_t W, unsigned long long* stats, const int2* __restrict__ cand,
    int64_t cand_stride, int2* pair_list, int64_t pair_stride, const int32_t* __restrict__ progress) {
    const int64_t n_all = cand[0].x;
    if (n_all <= 0 || n_all > cand_stride - 1) return;
    const int64_t done = progress ? progress[0] : 0;
    const int64_t n = n_all - done;
    if (n <= 0) return;
    cand += done;
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nslab = num_slabs(M);
    const int grp = lane >> 2, sub = lane & 3;
    const double thr2 = 2.0 * thr;
    unsigned long long n_cand = 0, n_ok = 0, n_near = 0, n_deg = 0;
    // candidates per warp and step: 32 (one eigen-solve per lane) when the list keeps every warp busy; with a short list
    // (several ranks: 1/world of the candidates each) fewer, in multiples of the 8 a pass handles, so that the work
    // spreads over all warps instead of queueing three passes + solve + three passes behind each other in a few
    const int64_t per_warp = (n + nwarps - 1) / nwarps;
    const int bs = per_warp >= 32 ? 32 : (int)((per_warp + 7) & ~int64_t(7));
    for (int64_t b0 = gwarp * bs; b0 < n; b0 += nwarps * bs) {
        const int count = (int)((n - b0 < bs) ? n - b0 : bs);
        int32_t lrow = 0, j = 0, i = 0;
        if (lane < count) {
            const int2 e = cand[1 + b0 + lane];
            lrow = e.x; j = e.y;
            i = row_blocks[lrow / CB] * CB + (lrow % CB);
        }
        // ---- phase A: covariances, 8 candidates per pass, 4 lanes (slabs) per candidate ----
        double Sk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int pass = 0; pass * 8 < count; pass++) {
            const int c = pass * 8 + grp;
            const int64_t ci = __shfl_sync(0xffffffffu, i, c), cj = __shfl_sync(0xffffffffu, j, c);
            double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (c < count)
                for (int sl = sub; sl < nslab; sl += 4) {
                    const int na = (M - sl * KS < KS) ? M - sl * KS : KS;
                    const double* P = packed + packed_index(ci, sl * KS, 0, nb_pad);
                    const double* Q = packed + packed_index(cj, sl * KS, 0, nb_pad);
                    if (na == KS) slab_cov<true>(P, Q, KS, S);
                    else slab_cov<false>(P, Q, (na + 1) & ~1, S);      // padding atoms are zero
                }
#pragma unroll
            for (int q = 0; q < 9; q++) {
                S[q] += __shfl_xor_sync(0xffffffffu, S[q], 1);
                S[q] += __shfl_xor_sync(0xffffffffu, S[q], 2);
                const double v = __shfl_sync(0xffffffffu, S[q], 4 * (lane & 7));
                if ((lane >> 3) == pass) Sk[q] = v;
            }
        }
        // ---- phase B: one eigen-solve per lane ----
        double Rk[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, lam = 1.0, gap = 1.0;
        if (lane < count) kabsch_rot_from_cov(Sk, Rk, &lam, &gap);
        // ---- phase C: explicit rotation + differences ----
        uint32_t okmask = 0;
        for (int pass = 0; pass * 8 < count; pass++) {
            const int c = pass * 8 + grp;
            const int64_t ci = __shfl_sync(0xffffffffu, i, c), cj = __shfl_sync(0xffffffffu, j, c);
            const int32_t crow = __shfl_sync(0xffffffffu, lrow, c);
            double R[9];
#pragma unroll
            for (int q = 0; q < 9; q++) R[q] = __shfl_sync(0xffffffffu, Rk[q], c);
            const double lk = __shfl_sync(0xffffffffu, lam, c), gk = __shfl_sync(0xffffffffu, gap, c);
            double ss = 0.0, mx = 0.0;
            if (c < count)
                for (int sl = sub; sl < nslab; sl += 4) {
                    const int na = (M - sl * KS < KS) ? M - sl * KS : KS;
                    const double* P = packed + packed_index(ci, sl * KS, 0, nb_pad);
                    const double* Q = packed + packed_index(cj, sl * KS, 0, nb_pad);
                    if (na == KS) slab_diff<true>(P, Q, KS, R, ss, mx);
                    else slab_diff<false>(P, Q, na, R, ss, mx);
                }
            ss += __shfl_xor_sync(0xffffffffu, ss, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const double rmsd = sqrt(ss / (double)M), maxdev = sqrt(mx);
            const bool ok = (rmsd < thr) && (maxdev < thr2);
            const bool mine = (c < count) && sub == 0;
            if (mine) {
                n_cand++;
                n_ok += ok;
                n_near += (fabs(rmsd - thr) < 1e-6) || ((rmsd < thr) && fabs(maxdev - thr2) < 1e-6);
                n_deg += ok && (gk < 1e-9 * fabs(lk));
                if (!ok) atomicAnd(&sim_bits[(int64_t)crow * W + (cj >> 5)], ~(1u << (cj & 31)));
            }
            const uint32_t okb = __ballot_sync(0xffffffffu, mine && ok);      // bit 4g set <=> candidate pass*8+g confirmed
#pragma unroll
            for (int g = 0; g < 8; g++) okmask |= ((okb >> (4 * g)) & 1u) << (pass * 8 + g);
        }
        if (pair_list && okmask) {
            int64_t base = 0;
            if (lane == 0) base = list_reserve(&pair_list[0].x, __popc(okmask), pair_stride - 1);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((okmask >> lane) & 1u) {
                const int64_t slot = base + __popc(okmask & ((1u << lane) - 1u));
                if (slot < pair_stride - 1) pair_list[1 + slot] = make_int2(i, j);
            }
        }
    }
    n_cand = warp_sum_u64(n_cand); n_ok = warp_sum_u64(n_ok); n_near = warp_sum_u64(n_near); n_deg = warp_sum_u64(n_deg);
    if (lane == 0 && stats) {
        if (n_cand) atomicAdd(&stats[0], n_cand);
        if (n_ok) atomicAdd(&stats[1], n_ok);
        if (n_near) atomicAdd(&stats[2], n_near);
        if (n_deg) atomicAdd(&stats[3], n_deg);
    }
}

// ------------------------------------------------------------------------------------------
// The same verification with COALESCED loads (default).  ncu on the kernel above (profiles/r02_ncu_kernels.md): L1/TEX
// throughput 89 %, FP64 pipe 16 % — every lane of a load instruction reads its own 160-byte run, 32 different cache
// lines per instruction, and the L1 processes one line per cycle: 2 x 240 line-cycles per candidate, which is exactly
// the 0.45 ms the kernel takes on C3.  Here the 32 lanes of a warp walk ONE candidate's atoms together (lane = atom:
// the x / y / z runs of a 20-atom slab are contiguous in the packed layout, so a load instruction touches 2-3 lines),
// the nine covariance sums are reduced with a halving butterfly (lanes exchange the half of the values they do not
// keep: 4 + 2 + 1 + 2 shuffles for eight values instead of 8 x 5) and parked in shared memory for lane k, the
// eigen-solves still run one per lane for the whole batch.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

// R = ceil(M / 32) rounds of the atom loop as a compile-time constant (1..4; 0 = any M, rolled loop): with the rounds
// unrolled all 6 R loads of a candidate are issued before the first use, one L2 round trip per candidate and phase
// instead of R.
template <int R>
__global__ void __launch_bounds__(VF_WARPS * 32) rmsd_verify_list_coop_kernel(
    const double* __restrict__ packed, int64_t N, int M, int64_t nb_pad, const int32_t* __restrict__ row_blocks,
    double thr, uint32_t* sim_bits, int64_t W, unsigned long long* stats, const int2* __restrict__ cand,
    int64_t cand_stride, int2* pair_list, int64_t pair_stride, const int32_t* __restrict__ progress) {
    // entries [progress[0], count) of the list: a caller that verifies while the screen is still producing (the
    // pipelined upload) passes the number of entries earlier calls have dealt with (tsc_rmsd_verify_incr)
    const int64_t n_all = cand[0].x;
    if (n_all <= 0 || n_all > cand_stride - 1) return;
    const int64_t done = progress ? progress[0] : 0;
    const int64_t n = n_all - done;
    if (n <= 0) return;
    cand += done;
    __shared__ double s_cov[VF_WARPS][32][9];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double thr2 = 2.0 * thr;
    const int64_t comp = (int64_t)CB * KS, slab_stride = nb_pad * 3 * comp;
    unsigned long long n_cand = 0, n_ok = 0, n_near = 0, n_deg = 0;
    const int64_t per_warp = (n + nwarps - 1) / nwarps;
    const int bs = per_warp >= 32 ? 32 : (int)per_warp;           // short list: spread it over all warps
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    int64_t offs[R > 0 ? R : 1];                                  // this lane's atom of every round inside a conformer
#pragma unroll
    for (int r = 0; r < (R > 0 ? R : 1); r++) {
        const int m = lane + 32 * r;
        offs[r] = (int64_t)(m / KS) * slab_stride + (m % KS);
    }
    for (int64_t b0 = gwarp * bs; b0 < n; b0 += nwarps * bs) {
        const int count = (int)((n - b0 < bs) ? n - b0 : bs);
        int32_t lrow = 0, j = 0, i = 0;
        int64_t oi = 0, oj = 0;                                   // offset of the conformer inside a slab image
        if (lane < count) {
            const int2 e = cand[1 + b0 + lane];
            lrow = e.x; j = e.y;
            i = row_blocks[lrow / CB] * CB + (lrow % CB);
            oi = (int64_t)(i / CB) * 3 * comp + (int64_t)(i % CB) * KS;
            oj = (int64_t)(j / CB) * 3 * comp + (int64_t)(j % CB) * KS;
        }
        // ---- phase A: covariances, one candidate at a time, lane = atom.  (Measured and rejected: loading candidate
        // k + 1 into a second register set while k is reduced — 196 registers, half the resident warps — and
        // prefetching its lines into L1 — 0.36 against 0.33 ms on C3: the extra instructions cost more.) ----
        constexpr int RR = R > 0 ? R : 1;
        auto load_pair = [&](int k, double (&c)[RR][6], bool zero_q) {
            const double* P = packed + __shfl_sync(0xffffffffu, oi, k);
            const double* Q = packed + __shfl_sync(0xffffffffu, oj, k);
#pragma unroll
            for (int r = 0; r < RR; r++) {
                const bool in = lane + 32 * r < M;
                const int64_t o = in ? offs[r] : 0;                   // (atom 0: always there)
                c[r][0] = P[o]; c[r][1] = P[o + comp]; c[r][2] = P[o + 2 * comp];
                c[r][3] = Q[o]; c[r][4] = Q[o + comp]; c[r][5] = Q[o + 2 * comp];
                if (!in) {
                    c[r][0] = c[r][1] = c[r][2] = 0.0;
                    if (zero_q) c[r][3] = c[r][4] = c[r][5] = 0.0;
                }
            }
        };
        auto reduce_store = [&](int k, double (&v)[9]) {
            // halving butterfly over v[0..7]: after the three exchanges lane L holds entry 4 b4 + 2 b3 + b2
            double w4[4], w2[2], w1;
#pragma unroll
            for (int q = 0; q < 4; q++) w4[q] = (b4 ? v[q + 4] : v[q]) + shfl_xor_d(b4 ? v[q] : v[q + 4], 16);
#pragma unroll
            for (int q = 0; q < 2; q++) w2[q] = (b3 ? w4[q + 2] : w4[q]) + shfl_xor_d(b3 ? w4[q] : w4[q + 2], 8);
            w1 = (b2 ? w2[1] : w2[0]) + shfl_xor_d(b2 ? w2[0] : w2[1], 4);
            w1 += shfl_xor_d(w1, 2);
            w1 += shfl_xor_d(w1, 1);
            double v8 = v[8];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v8 += shfl_xor_d(v8, o);
            if ((lane & 3) == 0) s_cov[warp][k][lane >> 2] = w1;
            if (lane == 0) s_cov[warp][k][8] = v8;
        };
        auto cov_of = [&](int k, const double (&c)[RR][6]) {
            double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int r = 0; r < RR; r++) {
                v[0] = fma(c[r][0], c[r][3], v[0]); v[1] = fma(c[r][0], c[r][4], v[1]); v[2] = fma(c[r][0], c[r][5], v[2]);
                v[3] = fma(c[r][1], c[r][3], v[3]); v[4] = fma(c[r][1], c[r][4], v[4]); v[5] = fma(c[r][1], c[r][5], v[5]);
                v[6] = fma(c[r][2], c[r][3], v[6]); v[7] = fma(c[r][2], c[r][4], v[7]); v[8] = fma(c[r][2], c[r][5], v[8]);
            }
            reduce_store(k, v);
        };
        if (R > 0) {
            double c0[RR][6];
            for (int k = 0; k < count; k++) {
                load_pair(k, c0, false);
                cov_of(k, c0);
            }
        } else {
            for (int k = 0; k < count; k++) {
                const double* P = packed + __shfl_sync(0xffffffffu, oi, k);
                const double* Q = packed + __shfl_sync(0xffffffffu, oj, k);
                double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                for (int m = lane; m < M; m += 32) {
                    const int64_t o = (int64_t)(m / KS) * slab_stride + (m % KS);
                    const double px = P[o], py = P[o + comp], pz = P[o + 2 * comp];
                    const double qx = Q[o], qy = Q[o + comp], qz = Q[o + 2 * comp];
                    v[0] = fma(px, qx, v[0]); v[1] = fma(px, qy, v[1]); v[2] = fma(px, qz, v[2]);
                    v[3] = fma(py, qx, v[3]); v[4] = fma(py, qy, v[4]); v[5] = fma(py, qz, v[5]);
                    v[6] = fma(pz, qx, v[6]); v[7] = fma(pz, qy, v[7]); v[8] = fma(pz, qz, v[8]);
                }
                reduce_store(k, v);
            }
        }
        __syncwarp();
        // ---- phase B: one eigen-solve per lane ----
        double Rk[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, lam = 1.0, gap = 1.0;
        if (lane < count) {
            double Sk[9];
#pragma unroll
            for (int q = 0; q < 9; q++) Sk[q] = s_cov[warp][lane][q];
            kabsch_rot_from_cov(Sk, Rk, &lam, &gap);
        }
        __syncwarp();
        // ---- phase C: explicit rotation + differences, one candidate at a time ----
        double my_ss = 0.0, my_mx = 0.0;                          // lane k: sum and maximum of candidate k
        auto finish = [&](int k, double ss, double mx) {
            // two values, one exchange: the lower half-warp finishes the sum, the upper one the maximum; lane k keeps
            // both (the divisions and square roots of the whole batch then run once, one candidate per lane)
            const double send = b4 ? ss : mx, got = shfl_xor_d(send, 16);
            double r = b4 ? fmax(mx, got) : ss + got;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const double t = shfl_xor_d(r, o);
                r = b4 ? fmax(r, t) : r + t;
            }
            const double s_all = __shfl_sync(0xffffffffu, r, 0), m_all = __shfl_sync(0xffffffffu, r, 16);
            if (lane == k) { my_ss = s_all; my_mx = m_all; }
        };
        auto diff_of = [&](int k, const double (&c)[RR][6]) {
            double Rm[9];
#pragma unroll
            for (int q = 0; q < 9; q++) Rm[q] = __shfl_sync(0xffffffffu, Rk[q], k);
            double ss = 0.0, mx = 0.0;
#pragma unroll
            for (int r = 0; r < RR; r++) {
                const double dx = fma(Rm[0], c[r][0], fma(Rm[1], c[r][1], Rm[2] * c[r][2])) - c[r][3];
                const double dy = fma(Rm[3], c[r][0], fma(Rm[4], c[r][1], Rm[5] * c[r][2])) - c[r][4];
                const double dz = fma(Rm[6], c[r][0], fma(Rm[7], c[r][1], Rm[8] * c[r][2])) - c[r][5];
                const double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
                ss += d2;
                mx = fmax(mx, d2);
            }
            finish(k, ss, mx);
        };
        if (R > 0) {
            double c0[RR][6];
            for (int k = 0; k < count; k++) {
                load_pair(k, c0, true);
                diff_of(k, c0);
            }
        } else {
            for (int k = 0; k < count; k++) {
                const double* P = packed + __shfl_sync(0xffffffffu, oi, k);
                const double* Q = packed + __shfl_sync(0xffffffffu, oj, k);
                double Rm[9];
#pragma unroll
                for (int q = 0; q < 9; q++) Rm[q] = __shfl_sync(0xffffffffu, Rk[q], k);
                double ss = 0.0, mx = 0.0;
                for (int m = lane; m < M; m += 32) {
                    const int64_t o = (int64_t)(m / KS) * slab_stride + (m % KS);
                    const double px = P[o], py = P[o + comp], pz = P[o + 2 * comp];
                    const double dx = fma(Rm[0], px, fma(Rm[1], py, Rm[2] * pz)) - Q[o];
                    const double dy = fma(Rm[3], px, fma(Rm[4], py, Rm[5] * pz)) - Q[o + comp];
                    const double dz = fma(Rm[6], px, fma(Rm[7], py, Rm[8] * pz)) - Q[o + 2 * comp];
                    const double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
                    ss += d2;
                    mx = fmax(mx, d2);
                }
                finish(k, ss, mx);
            }
        }
        bool ok = false;
        if (lane < count) {
            const double rmsd = sqrt(my_ss / (double)M), maxdev = sqrt(my_mx);
            ok = (rmsd < thr) && (maxdev < thr2);
            n_cand++;
            n_ok += ok;
            n_near += (fabs(rmsd - thr) < 1e-6) || ((rmsd < thr) && fabs(maxdev - thr2) < 1e-6);
            n_deg += ok && (gap < 1e-9 * fabs(lam));
            if (!ok) atomicAnd(&sim_bits[(int64_t)lrow * W + (j >> 5)], ~(1u << (j & 31)));
        }
        const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
        if (pair_list && okmask) {
            int64_t base = 0;
            if (lane == 0) base = list_reserve(&pair_list[0].x, __popc(okmask), pair_stride - 1);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((okmask >> lane) & 1u) {
                const int64_t slot = base + __popc(okmask & ((1u << lane) - 1u));
                if (slot < pair_stride - 1) pair_list[1 + slot] = make_int2(i, j);
            }
        }
        __syncwarp();
    }
    n_cand = warp_sum_u64(n_cand); n_ok = warp_sum_u64(n_ok); n_near = warp_sum_u64(n_near); n_deg = warp_sum_u64(n_deg);
    if (lane == 0 && stats) {
        if (n_cand) atomicAdd(&stats[0], n_cand);
        if (n_ok) atomicAdd(&stats[1], n_ok);
        if (n_near) atomicAdd(&stats[2], n_near);
        if (n_deg) atomicAdd(&stats[3], n_deg);
    }
}

// Batched rmsd_and_max_numba on explicit AoS pairs: P, Q are (n, M, 3); one warp per pair.
__global__ void __launch_bounds__(256) rmsd_pairs_kernel(const double* __restrict__ P, const double* __restrict__ Q,
                                                         int64_t n, int M, int64_t q_stride_is_zero,
                                                         double* __restrict__ rmsd, double* __restrict__ maxdev) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const AosView VP{P, M}, VQ{Q, M};
    for (int64_t k = warp_g; k < n; k += nwarps) {
        // q_stride_is_zero: compare one reference P[0] against every Q[k] (the _rmsd_similarity shape)
        const PairEval ev = eval_pair(VP, q_stride_is_zero ? 0 : k, VQ, k, M, lane);
        if (lane == 0) { rmsd[k] = ev.rmsd; maxdev[k] = ev.maxdev; }
    }
}

// ------------------------------------------------------------------------------------------
// Group-local de-duplication of generated poses (the step right after the clash test inside the cyclical
// embeds, tscode/embeds.py:714-718 / 842-846):
//     if compenetration_check(pose): if not _rmsd_similarity(pose, angular_poses, rmsd_thr=1): keep, append
// i.e. within one group (same conformers, pairing and orientation; <= (steps+1)^F poses) a pose is kept iff it
// passed the clash test and is not similar (all atoms, rmsd < thr and max deviation < 2 thr,
// rmsd_pruning.py:208-224) to any pose of the group kept BEFORE it.  Groups are independent; inside a group the
// pair similarities do not depend on the greedy state, so they are computed for all pairs of clash-passing
// poses at once (one warp per pair) and the greedy pass itself is a tiny sequential kernel, one thread per group.
//   S (P, A, 3) poses;  pair k = (pi[k], pj[k]) with pj[k] earlier than pi[k] in the same group;  sim[k] out.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rmsd_pairs_idx_kernel(const double* __restrict__ S, const int32_t* __restrict__ pi,
                                                             const int32_t* __restrict__ pj, int64_t n, int M, double thr,
                                                             uint8_t* __restrict__ sim) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const AosView V{S, M};
    for (int64_t k = warp_g; k < n; k += nwarps) {
        // ref = the new pose (rotated onto the kept one), as _rmsd_similarity(ref=pose, structures=kept) does
        const PairEval ev = eval_pair(V, pi[k], V, pj[k], M, lane);
        if (lane == 0) sim[k] = (uint8_t)((ev.rmsd < thr) && (ev.maxdev < 2.0 * thr));
    }
}

// one thread per group: members [g_begin[g], g_begin[g+1]) of `order` (pose indices of the clash-passing poses in
// generation order); pair_base[m] = index of the first pair of member m, its pairs being (m, earlier member 0..r-1)
__global__ void group_greedy_kernel(const int32_t* __restrict__ g_begin, int32_t n_groups, const int32_t* __restrict__ order,
                                    const int64_t* __restrict__ pair_base, const uint8_t* __restrict__ sim,
                                    uint8_t* __restrict__ keep) {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += gridDim.x * blockDim.x) {
        const int b = g_begin[g], e = g_begin[g + 1];
        for (int m = b; m < e; m++) {
            const uint8_t* row = sim + pair_base[m];
            bool dup = false;
            for (int q = b; q < m && !dup; q++) dup = row[q - b] && keep[order[q]];
            keep[order[m]] = (uint8_t)!dup;
        }
    }
}

}  // namespace tsc

namespace tsc {
__global__ void verify_progress_kernel(const int32_t* cand_header, int64_t cand_stride, int32_t* progress) {
    const int64_t n = cand_header[0];
    if (n >= 0 && n <= cand_stride - 1) progress[0] = (int32_t)n;       // (an overflowed list is never advanced)
}
}  // namespace tsc

static int verify_impl(const double* packed, int64_t N, int32_t M, const int32_t* row_blocks, int32_t n_rb, double thr,
                       uint32_t* sim_bits, uint64_t* stats, int32_t* pair_list, int64_t pair_stride,
                       const int32_t* cand_list, int64_t cand_stride, int32_t* progress, int32_t final_call, void* stream);

extern "C" int tsc_rmsd_verify(const double* packed, int64_t N, int32_t M, const int32_t* row_blocks,
                               int32_t n_rb, double thr, uint32_t* sim_bits, uint64_t* stats, int32_t* pair_list,
                               int64_t pair_stride, const int32_t* cand_list, int64_t cand_stride, void* stream) {
    return verify_impl(packed, N, M, row_blocks, n_rb, thr, sim_bits, stats, pair_list, pair_stride, cand_list, cand_stride,
                       nullptr, 1, stream);
}

// Incremental form for callers that interleave screen launches and verification on one stream (the pipelined upload
// of rmsd_pruning.py): verifies the candidate-list entries appended since the previous call and then advances
// progress[0] (device int32, zero it before the first screen launch).  final_call != 0 additionally runs the bit-row
// scan that takes over when the list has overflowed (it re-examines every set bit, so it must run once, at the end).
extern "C" int tsc_rmsd_verify_incr(const double* packed, int64_t N, int32_t M, const int32_t* row_blocks,
                                    int32_t n_rb, double thr, uint32_t* sim_bits, uint64_t* stats, int32_t* pair_list,
                                    int64_t pair_stride, const int32_t* cand_list, int64_t cand_stride,
                                    int32_t* progress, int32_t final_call, void* stream) {
    if (!cand_list || !progress) return (int)cudaErrorInvalidValue;
    return verify_impl(packed, N, M, row_blocks, n_rb, thr, sim_bits, stats, pair_list, pair_stride, cand_list, cand_stride,
                       progress, final_call, stream);
}

static int verify_impl(const double* packed, int64_t N, int32_t M, const int32_t* row_blocks, int32_t n_rb, double thr,
                       uint32_t* sim_bits, uint64_t* stats, int32_t* pair_list, int64_t pair_stride,
                       const int32_t* cand_list, int64_t cand_stride, int32_t* progress, int32_t final_call, void* stream) {
    using namespace tsc;
    if (N <= 0 || n_rb <= 0) return 0;
    const int64_t nb_pad = num_blocks_padded(N);
    if (cand_list) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int rounds = (M + 31) / 32;
        auto kern = rounds == 1 ? rmsd_verify_list_coop_kernel<1> : rounds == 2 ? rmsd_verify_list_coop_kernel<2>
                  : rounds == 3 ? rmsd_verify_list_coop_kernel<3> : rounds == 4 ? rmsd_verify_list_coop_kernel<4>
                  : rounds == 5 ? rmsd_verify_list_coop_kernel<5> : rmsd_verify_list_coop_kernel<0>;
#ifdef TSC_VERIFY_PER_LANE
        kern = rmsd_verify_list_kernel;                           // the previous form: 4 lanes per candidate, per-lane runs
#endif
        kern<<<sms * 4, VF_WARPS * 32, 0, (cudaStream_t)stream>>>(
            packed, N, M, nb_pad, row_blocks, thr, sim_bits, nb_pad, reinterpret_cast<unsigned long long*>(stats),
            reinterpret_cast<const int2*>(cand_list), cand_stride, reinterpret_cast<int2*>(pair_list), pair_stride,
            progress);
        TSC_CHECK_LAUNCH();
        if (progress) {
            verify_progress_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(cand_list, cand_stride, progress);
            TSC_CHECK_LAUNCH();
        }
    }
    if (!final_call) return 0;
    int64_t rows = (int64_t)n_rb * CB;
    int64_t blocks = (rows + 4 * VF_WARPS - 1) / (4 * VF_WARPS);       // ~4 rows per warp: fuller batches
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    rmsd_verify_kernel<<<(unsigned)blocks, VF_WARPS * 32, 0, (cudaStream_t)stream>>>(
        packed, N, M, nb_pad, row_blocks, n_rb, thr, sim_bits, nb_pad, reinterpret_cast<unsigned long long*>(stats),
        reinterpret_cast<int2*>(pair_list), pair_stride, reinterpret_cast<const int2*>(cand_list), cand_stride);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_rmsd_pairs(const double* P, const double* Q, int64_t n, int32_t M, int32_t broadcast_p,
                              double* rmsd, double* maxdev, void* stream) {
    using namespace tsc;
    if (n <= 0) return 0;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    rmsd_pairs_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(P, Q, n, M, broadcast_p, rmsd, maxdev);
    TSC_CHECK_LAUNCH();
    return 0;
}

// Similarity of explicit index pairs of one pose array: sim[k] = rmsd_and_max(S[pi[k]], S[pj[k]]) passes
// (rmsd < thr and maxdev < 2 thr), all M atoms.
extern "C" int tsc_rmsd_pairs_idx(const double* S, const int32_t* pi, const int32_t* pj, int64_t n, int32_t M,
                                  double thr, uint8_t* sim, void* stream) {
    using namespace tsc;
    if (n <= 0) return 0;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    rmsd_pairs_idx_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, pi, pj, n, M, thr, sim);
    TSC_CHECK_LAUNCH();
    return 0;
}

// Greedy pass of the group-local de-duplication (embeds.py:714-718): keep[order[m]] = 1 iff member m is not similar
// to an earlier KEPT member of its group.  keep must be zero for poses that are not members.
extern "C" int tsc_group_greedy(const int32_t* g_begin, int32_t n_groups, const int32_t* order, const int64_t* pair_base,
                                const uint8_t* sim, uint8_t* keep, void* stream) {
    if (n_groups <= 0) return 0;
    int blocks = (n_groups + 127) / 128;
    if (blocks > 148 * 8) blocks = 148 * 8;
    tsc::group_greedy_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(g_begin, n_groups, order, pair_base, sim, keep);
    TSC_CHECK_LAUNCH();
    return 0;
}
