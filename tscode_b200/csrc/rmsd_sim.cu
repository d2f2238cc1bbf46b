// rmsd_sim.cu — all-pairs Kabsch similarity screen: the hot kernel of prune_conformers_rmsd.
//
// What the reference does per pair (tscode/rmsd_pruning.py:6-41, called from :70) is a 3xM.Mx3
// cross-covariance, a 3x3 SVD, a rotation and an explicit difference.  Here the cross-
// covariances of a 32 x 64 tile of conformer pairs are one dense contraction
//     C[(i,a),(j,b)] = sum_m P[(i,a),m] * Q[m,(j,b)]         (3*32 x M) . (M x 3*64)
// run either on the FP64 tensor cores (DMMA.8x8x4, variant 0) or on the FP64 FMA pipe
// (variant 1), followed per pair by a closed-form, iteration-free screen on the quartic of
// Horn's key matrix (tsc_math.cuh: Budan-Fourier sign test at the threshold eigenvalue).
// Pairs that survive the screen get their bit set in sim_bits; rmsd_verify.cu then re-evaluates
// exactly those pairs the way the reference does (explicit rotation, RMSD and max deviation)
// and clears the bits that fail.  Bits are final only after verification.
//
// Structure (one persistent CTA per SM, 12 warps = 3 warpgroups, registers re-split with setmaxnreg):
//   warp 8, lane 0 : producer (warpgroup 2 shrinks to 40 registers/thread) — walks this CTA's
//                    tiles a full ring ahead of the math, streaming (slab, block) chunks of the
//                    packed ensemble into a 4-stage shared-memory ring with 1-D bulk TMA
//                    (cp.async.bulk -> UBLKCP), completion on "full" mbarriers; with the last
//                    slab of a tile it also delivers the tile descriptor and the 32 + 64 squared
//                    norms the epilogue needs, so consumers never touch global memory for input
//   warps 0..7     : consumers (warpgroups 0-1 grow to 232 registers/thread; 72 FP64
//                    accumulators each) — each owns a 16 x 16 pair sub-tile, waits on "full",
//                    issues the MMAs/FMAs from shared memory, releases the stage on an "empty"
//                    mbarrier, and after the last slab runs the screen and writes 16 result bits
//                    per row straight to HBM.
//   (round-1 first version had thread 0 double as producer two slabs ahead: ncu showed 4.9 % of
//   warp samples in the full-barrier wait and 2.4 % on the epilogue's G loads, profiles/r01_*.)
// No __syncthreads() after setup; HBM traffic is 1 bit per pair out, operands come from L2.
//
// Roofline: FP64 pipe (tensor or FMA).  Algorithmic work 18*M flop per pair.
#include "tsc_common.cuh"
#include "tsc_math.cuh"

namespace tsc {

constexpr int SIM_NSTAGE = 4;
constexpr int SIM_DATA_D = 3 * CHUNK_D;                 // I chunk + 2 J chunks, doubles
constexpr int SIM_DATA_BYTES = SIM_DATA_D * 8;          // 46080
constexpr int SIM_G_D = CB + 2 * CB;                    // G of the 32 rows and 64 columns (last slab only)
constexpr int SIM_G_BYTES = SIM_G_D * 8;                // 768
constexpr int SIM_STAGE_D = SIM_DATA_D + SIM_G_D + 2;   // + int4 tile descriptor
constexpr int SIM_STAGE_BYTES = SIM_STAGE_D * 8;        // 46864
constexpr int SIM_CONSUMER_WARPS = 8;
constexpr int SIM_THREADS = (SIM_CONSUMER_WARPS + 4) * 32;   // + producer warpgroup
constexpr int SIM_REGS_CONSUMER = 232;
constexpr int SIM_REGS_PRODUCER = 40;
constexpr size_t SIM_SMEM_BYTES = (size_t)SIM_NSTAGE * SIM_STAGE_BYTES + 2 * SIM_NSTAGE * sizeof(uint64_t);

struct SimParams {
    const double* packed;
    const double* G;
    const int4* tiles;        // (ib, jp, lb, unused): I block ib, J blocks 2jp,2jp+1, local row block lb
    int64_t n_tiles;
    int64_t N;
    int64_t nb_pad;
    int nslab;
    int M;
    double e_thr;             // M * thr^2 * (1 + 1e-6)
    double g_eps;             // 1e-10: margin proportional to G_i + G_j
    uint16_t* sim_bits16;     // sim_bits viewed as halfwords; row stride 2*W halfwords
    int64_t W;                // words per row (= nb_pad)
};

// ---- per-pair screen shared by both variants ----------------------------------------------
__device__ __forceinline__ uint32_t screen_bit(const double S[9], double Gi, double Gj, int64_t i, int64_t j,
                                               const SimParams& p) {
    if (!(j > i) || j >= p.N) return 0u;
    const double Gs = Gi + Gj;
    return screen_candidate(S, Gs, fma(p.g_eps, Gs, p.e_thr)) ? 1u : 0u;
}

// ---- variant 0: FP64 tensor cores ------------------------------------------------------------
// warp tile 16 (i) x 16 (j) = 2 x 2 DMMA tiles x 9 (a,b) component pairs
struct ConsumerDMMA {
    double acc[2][2][9][2];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int v = 0; v < 2; v++)
#pragma unroll
                for (int c = 0; c < 9; c++) acc[u][v][c][0] = acc[u][v][c][1] = 0.0;
    }
    // Ibase/Jbase: this warp's first conformer row inside the stage's I / J image
    __device__ __forceinline__ void slab(const double* __restrict__ Iw, const double* __restrict__ Jw, int lane) {
        const int fo = (lane >> 2) * KS + (lane & 3);
#pragma unroll
        for (int ks = 0; ks < KS / 4; ks++) {
            double a[2][3], b[2][3];
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    a[u][c] = Iw[c * (CB * KS) + u * 8 * KS + ks * 4 + fo];
                    b[u][c] = Jw[c * (CB * KS) + u * 8 * KS + ks * 4 + fo];
                }
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int v = 0; v < 2; v++)
#pragma unroll
                    for (int ca = 0; ca < 3; ca++)
#pragma unroll
                        for (int cb = 0; cb < 3; cb++)
                            dmma884(acc[u][v][3 * ca + cb][0], acc[u][v][3 * ca + cb][1], a[u][ca], b[v][cb]);
        }
    }
    // i0/j0: global conformer index of the warp tile's first row / column
    __device__ __forceinline__ void epilogue(int64_t i0, int64_t j0, int64_t row0, int lane, const SimParams& p,
                                             const double* __restrict__ gI, const double* __restrict__ gJ) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int64_t i = i0 + u * 8 + (lane >> 2);
            const double Gi = gI[u * 8 + (lane >> 2)];
            uint32_t bits = 0;
#pragma unroll
            for (int v = 0; v < 2; v++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int jc = v * 8 + 2 * (lane & 3) + e;
                    const int64_t j = j0 + jc;
                    double S[9];
#pragma unroll
                    for (int c = 0; c < 9; c++) S[c] = acc[u][v][c][e];
                    bits |= screen_bit(S, Gi, gJ[jc], i, j, p) << jc;
                }
            bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
            bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
            if ((lane & 3) == 0 && i < p.N)
                p.sim_bits16[(row0 + u * 8 + (lane >> 2)) * (2 * p.W) + (j0 >> 4)] = (uint16_t)bits;
        }
    }
};

// ---- variant 1: FP64 FMA pipe ------------------------------------------------------------------
// warp tile 16 x 16; thread owns rows g + 4r (g = lane>>3, r = 0..3) and columns h + 8s
// (h = lane&7, s = 0..1); coordinates are read two atoms at a time as 16-byte vectors.
struct ConsumerFMA {
    double acc[4][2][9];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int c = 0; c < 9; c++) acc[r][s][c] = 0.0;
    }
    __device__ __forceinline__ void slab(const double* __restrict__ Iw, const double* __restrict__ Jw, int lane) {
        const int g = lane >> 3, h = lane & 7;
#pragma unroll 2
        for (int k2 = 0; k2 < KS / 2; k2++) {
            double2 a[4][3], b[2][3];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 3; c++)
                    a[r][c] = *reinterpret_cast<const double2*>(Iw + c * (CB * KS) + (g + 4 * r) * KS + 2 * k2);
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int c = 0; c < 3; c++)
                    b[s][c] = *reinterpret_cast<const double2*>(Jw + c * (CB * KS) + (h + 8 * s) * KS + 2 * k2);
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int s = 0; s < 2; s++)
#pragma unroll
                    for (int ca = 0; ca < 3; ca++)
#pragma unroll
                        for (int cb = 0; cb < 3; cb++) {
                            double t = fma(a[r][ca].x, b[s][cb].x, acc[r][s][3 * ca + cb]);
                            acc[r][s][3 * ca + cb] = fma(a[r][ca].y, b[s][cb].y, t);
                        }
        }
    }
    __device__ __forceinline__ void epilogue(int64_t i0, int64_t j0, int64_t row0, int lane, const SimParams& p,
                                             const double* __restrict__ gI, const double* __restrict__ gJ) {
        const int g = lane >> 3, h = lane & 7;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int64_t i = i0 + g + 4 * r;
            const double Gi = gI[g + 4 * r];
            uint32_t bits = 0;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int jc = h + 8 * s;
                const int64_t j = j0 + jc;
                bits |= screen_bit(acc[r][s], Gi, gJ[jc], i, j, p) << jc;
            }
            bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
            bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
            bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
            if (h == 0 && i < p.N) p.sim_bits16[(row0 + g + 4 * r) * (2 * p.W) + (j0 >> 4)] = (uint16_t)bits;
        }
    }
};

template <class Consumer>
__global__ void __launch_bounds__(SIM_THREADS, 1) rmsd_sim_kernel(const SimParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)SIM_NSTAGE * SIM_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + SIM_NSTAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SIM_NSTAGE; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], SIM_CONSUMER_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    int stage = 0;
    uint32_t phase = 0;
    if (warp >= SIM_CONSUMER_WARPS) {
        // ===== producer warpgroup =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SIM_REGS_PRODUCER));
        if (warp == SIM_CONSUMER_WARPS && lane == 0) {
            for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                const int4 tl = p.tiles[t];
                for (int s = 0; s < p.nslab; s++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    double* dst = stages + (size_t)stage * SIM_STAGE_D;
                    const double* srcI = p.packed + ((int64_t)s * p.nb_pad + tl.x) * CHUNK_D;
                    const double* srcJ = p.packed + ((int64_t)s * p.nb_pad + 2 * (int64_t)tl.y) * CHUNK_D;
                    const bool last = (s == p.nslab - 1);
                    if (last) *reinterpret_cast<int4*>(dst + SIM_DATA_D + SIM_G_D) = tl;
                    mbar_arrive_expect_tx(&full_bar[stage], SIM_DATA_BYTES + (last ? SIM_G_BYTES : 0));
                    bulk_g2s(dst, srcI, CHUNK_BYTES, &full_bar[stage]);
                    bulk_g2s(dst + CHUNK_D, srcJ, 2 * CHUNK_BYTES, &full_bar[stage]);
                    if (last) {
                        bulk_g2s(dst + SIM_DATA_D, p.G + (int64_t)tl.x * CB, CB * 8, &full_bar[stage]);
                        bulk_g2s(dst + SIM_DATA_D + CB, p.G + (int64_t)tl.y * 2 * CB, 2 * CB * 8, &full_bar[stage]);
                    }
                    if (++stage == SIM_NSTAGE) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ===== consumer warpgroups =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SIM_REGS_CONSUMER));
    const int ihalf = warp & 1;            // rows 16*ihalf .. +15 of the 32-row I block
    const int jquart = warp >> 1;          // columns 16*jquart .. +15 of the 64-column J pair
    const int woffI = ihalf * 16 * KS;
    const int woffJ = CHUNK_D * (1 + (jquart >> 1)) + (jquart & 1) * 16 * KS;
    Consumer cons;
    for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        cons.zero();
        for (int s = 0; s < p.nslab - 1; s++) {
            mbar_wait(&full_bar[stage], phase);
            const double* st = stages + (size_t)stage * SIM_STAGE_D;
            cons.slab(st + woffI, st + woffJ, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == SIM_NSTAGE) { stage = 0; phase ^= 1u; }
        }
        // last slab: its stage also carries G and the tile descriptor; release it after the epilogue
        mbar_wait(&full_bar[stage], phase);
        const double* st = stages + (size_t)stage * SIM_STAGE_D;
        cons.slab(st + woffI, st + woffJ, lane);
        const int4 tl = *reinterpret_cast<const int4*>(st + SIM_DATA_D + SIM_G_D);
        cons.epilogue((int64_t)tl.x * CB + ihalf * 16, (int64_t)tl.y * 2 * CB + jquart * 16,
                      (int64_t)tl.z * CB + ihalf * 16, lane, p, st + SIM_DATA_D + ihalf * 16,
                      st + SIM_DATA_D + CB + jquart * 16);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == SIM_NSTAGE) { stage = 0; phase ^= 1u; }
    }
}

}  // namespace tsc

extern "C" int tsc_rmsd_sim_tiles(const double* packed, const double* G, int64_t N, int32_t M,
                                  const int32_t* tiles, int64_t n_tiles, double thr, uint32_t* sim_bits,
                                  int32_t variant, int32_t grid_ctas, void* stream) {
    using namespace tsc;
    if (n_tiles <= 0 || N <= 0) return 0;
    SimParams p;
    p.packed = packed;
    p.G = G;
    p.tiles = reinterpret_cast<const int4*>(tiles);
    p.n_tiles = n_tiles;
    p.N = N;
    p.nb_pad = num_blocks_padded(N);
    p.nslab = num_slabs(M);
    p.M = M;
    p.e_thr = (double)M * thr * thr * (1.0 + 1e-6);
    p.g_eps = 1e-10;
    p.sim_bits16 = reinterpret_cast<uint16_t*>(sim_bits);
    p.W = p.nb_pad;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = grid_ctas > 0 ? grid_ctas : sms;
    if ((int64_t)grid > n_tiles) grid = (int)n_tiles;
    cudaError_t e;
    if (variant == 0) {
        e = cudaFuncSetAttribute(rmsd_sim_kernel<ConsumerDMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SIM_SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        rmsd_sim_kernel<ConsumerDMMA><<<grid, SIM_THREADS, SIM_SMEM_BYTES, (cudaStream_t)stream>>>(p);
    } else if (variant == 1) {
        e = cudaFuncSetAttribute(rmsd_sim_kernel<ConsumerFMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SIM_SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        rmsd_sim_kernel<ConsumerFMA><<<grid, SIM_THREADS, SIM_SMEM_BYTES, (cudaStream_t)stream>>>(p);
    } else {
        return (int)cudaErrorInvalidValue;
    }
    TSC_CHECK_LAUNCH();
    return 0;
}
