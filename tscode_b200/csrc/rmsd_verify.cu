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
                                                                   int64_t pair_stride) {
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
            int base = 0;
            if (lane == 0) base = atomicAdd(&pair_list[0].x, __popc(okmask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((okmask >> lane) & 1u) {
                const int64_t slot = (int64_t)base + __popc(okmask & ((1u << lane) - 1u));
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

}  // namespace tsc

extern "C" int tsc_rmsd_verify(const double* packed, int64_t N, int32_t M, const int32_t* row_blocks,
                               int32_t n_rb, double thr, uint32_t* sim_bits, uint64_t* stats, int32_t* pair_list,
                               int64_t pair_stride, void* stream) {
    using namespace tsc;
    if (N <= 0 || n_rb <= 0) return 0;
    const int64_t nb_pad = num_blocks_padded(N);
    int64_t rows = (int64_t)n_rb * CB;
    int64_t blocks = (rows + 4 * VF_WARPS - 1) / (4 * VF_WARPS);       // ~4 rows per warp: fuller batches
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    rmsd_verify_kernel<<<(unsigned)blocks, VF_WARPS * 32, 0, (cudaStream_t)stream>>>(
        packed, N, M, nb_pad, row_blocks, n_rb, thr, sim_bits, nb_pad, reinterpret_cast<unsigned long long*>(stats),
        reinterpret_cast<int2*>(pair_list), pair_stride);
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
