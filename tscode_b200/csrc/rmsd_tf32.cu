// rmsd_tf32.cu — tcgen05 / TMEM pre-screen of the all-pairs Kabsch similarity (variant 2).
//
// Idea: almost every pair of a conformer ensemble is FAR from the RMSD threshold, and a pair can
// be *proved* dissimilar from an approximate cross-covariance as long as the approximation error
// is bounded.  So the 3x3 covariances are first computed on the 5th-generation tensor cores in
// TF32 (operands pre-rounded to TF32 by pack_tf32, FP32 accumulation in TMEM), ~25x the FP64
// rate, and the screen of rmsd_sim.cu (Budan-Fourier sign test on the key-matrix quartic,
// tsc_math.cuh) is run in FP64 on the approximate S with the threshold eigenvalue lowered by a
// rigorous bound on the approximation error:
//     |S~_ab - S_ab| <= eps * sum_m |p_ma q_mb| <= eps * sqrt(G_i^a G_j^b)        (Cauchy-Schwarz)
//     => ||S~ - S||_F <= eps * sqrt(G_i G_j),   lambda_max(S) <= lambda_max(S~) + sqrt(3) ||S~ - S||_F
// with eps = 1.05e-3 covering two TF32 roundings (2 * 2^-11), the FP64->FP32 conversion and a
// generous allowance (80 * 2^-22 per term) for the tensor core's FP32 accumulation.  Pairs that
// cannot be excluded get their bit set, exactly like the FP64 variants, and rmsd_verify.cu
// re-evaluates them exactly in FP64 — so the final bits are identical; only the number of
// candidates differs (pairs within ~0.1 A above the threshold become candidates too).
//
// GEMM shape.  For a panel of 128 conformers i and a tile of 16 conformers j, three MMAs per
// 8-atom K block compute  D_a[i, (b, j)] = sum_m A_a[i, m] * B[(b, j), m]   (a, b in {x,y,z}):
// M = 128 (TMEM lanes = rows i), N = 48, K = 8, kind::tf32.  With the three D_a side by side in
// TMEM, thread i of the epilogue reads all nine entries of pair (i, j) from its own lane.
//
// Roles (one persistent CTA per SM, 10 warps):
//   warp 0 lane 0 : producer — bulk-TMA (cp.async.bulk) of the A panel (stationary, 1536*Mp B)
//                   and of a ring of B tiles (192*Mp B each), operands laid out in HBM by pack_tf32
//                   in the canonical no-swizzle K-major core-matrix order, so no tensor map is needed
//   warp 1        : TMEM allocation; lane 0 issues tcgen05.mma and tcgen05.commit
//   warps 2..9    : epilogue — two groups of 4 warps alternate over tiles: tcgen05.ld the 16 x 9
//                   accumulators of their row, release the TMEM buffer, run the FP64 screen,
//                   store 16 bits per row
// Pipelines: A full/empty, B ring full/empty, 3 TMEM accumulator buffers full/empty.
#include <cuda_fp16.h>
#include "tf32_common.cuh"

namespace tsc {

// pack: FP64 AoS -> TF32-rounded FP32 operand images + exact G, sqrt(G)
__global__ void __launch_bounds__(256) pack_tf32_kernel(const double* __restrict__ S, int64_t N, int A,
                                                        const int32_t* __restrict__ heavy_idx, int M, int Mp,
                                                        int64_t n_rows_pad, float* __restrict__ PA,
                                                        float* __restrict__ PB, float* __restrict__ PR,
                                                        double* __restrict__ G, double* __restrict__ sG,
                                                        float* __restrict__ CT) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + warp;          // one warp per conformer (incl. padding rows)
    if (i >= n_rows_pad) return;
    const bool live = i < N;
    const double* src = S + (live ? i : 0) * (int64_t)A * 3;
    const int nkc = Mp / 4;
    const int64_t panel = i / TF_ROWS, r = i % TF_ROWS;
    const int64_t jt = i / TF_J, jj = i % TF_J;
    double g = 0.0;
    for (int m = lane; m < Mp; m += 32) {
        double x = 0.0, y = 0.0, z = 0.0;
        if (live && m < M) {
            const double* a = src + (int64_t)heavy_idx[m] * 3;
            x = a[0]; y = a[1]; z = a[2];
            g = fma(x, x, fma(y, y, fma(z, z, g)));
        }
        const float fx = to_tf32((float)x), fy = to_tf32((float)y), fz = to_tf32((float)z);
        const int kc = m >> 2, e = m & 3;
        float* pa = PA + (((panel * 3) * nkc + kc) * TF_ROWS + r) * 4 + e;
        pa[0] = fx;
        pa[(int64_t)nkc * TF_ROWS * 4] = fy;
        pa[(int64_t)2 * nkc * TF_ROWS * 4] = fz;
        float* pb = PB + ((jt * nkc + kc) * TF_N + jj) * 4 + e;
        pb[0] = fx;
        pb[TF_J * 4] = fy;
        pb[2 * TF_J * 4] = fz;
        float* pr = PR + (size_t)i * 3 * Mp + m;          // row-major image for the TMEM-resident operand
        pr[0] = fx;
        pr[Mp] = fy;
        pr[2 * Mp] = fz;
    }
    g = warp_sum(g);
    if (lane == 0) {
        const double sg = sqrt(g);
        G[i] = g; sG[i] = sg;
        if (CT) {       // FP32 column terms of the Samuelson fast path, directed roundings (tf32_common.cuh)
            CT[jt * 32 + jj] = __double2float_rd(0.5 * (1.0 - 1e-10) * g);
            CT[jt * 32 + 16 + jj] = __double2float_ru(sg);
        }
    }
}

// pack, FP16 form: the same three images with 8 FP16 values per 16-byte K chunk (kind::f16, K = 16 per MMA;
// atoms padded to a multiple of 16).  FP16 keeps TF32's 10-bit mantissa (relative rounding error 2^-11) but
// has a 5-bit exponent: |x| >= 65520 becomes inf (the pair then fails every exclusion test, becomes a
// candidate and is decided exactly by the verify kernel), and values below the smallest normal 2^-14 are
// set to ZERO here, explicitly, so that no subnormal ever reaches the tensor core; their absolute error
// (<= 2^-14 each) enters the bound through sqrt(G)' = sqrt(G) + alpha sqrt(T), T = number of zeroed
// coordinates of the conformer, alpha = 2^-14 (1 + 2^-10) / eps:
//   ||S~ - S||_F <= 2 * 2^-11 (1 + 2^-11) sqrt(G_i G_j) + 2^-14 (1 + 2^-11) (sqrt(T_i G_j) + sqrt(G_i T_j))
//                <= eps sqrt(G_i)' sqrt(G_j)'                      (eps = 1.05e-3 as for TF32)
__global__ void __launch_bounds__(256) pack_f16_kernel(const double* __restrict__ S, int64_t N, int A,
                                                       const int32_t* __restrict__ heavy_idx, int M, int Mp,
                                                       int64_t n_rows_pad, __half* __restrict__ PA,
                                                       __half* __restrict__ PB, __half* __restrict__ PR,
                                                       double* __restrict__ G, double* __restrict__ sG,
                                                       float* __restrict__ CT, int64_t row_begin) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = row_begin + (int64_t)blockIdx.x * 8 + warp;      // one warp per conformer (incl. padding rows)
    if (i >= n_rows_pad) return;
    const bool live = i < N;
    const double* src = S + (live ? i : 0) * (int64_t)A * 3;
    const int nkc = Mp / 8;
    const int64_t panel = i / TF_ROWS, r = i % TF_ROWS;
    const int64_t jt = i / TF_J, jj = i % TF_J;
    double g = 0.0;
    int tiny = 0;
    const double fmin_normal = 6.103515625e-05;                 // 2^-14
    for (int m = lane; m < Mp; m += 32) {
        double x = 0.0, y = 0.0, z = 0.0;
        if (live && m < M) {
            const double* a = src + (int64_t)heavy_idx[m] * 3;
            x = a[0]; y = a[1]; z = a[2];
            g = fma(x, x, fma(y, y, fma(z, z, g)));
        }
        tiny += (x != 0.0 && fabs(x) < fmin_normal) + (y != 0.0 && fabs(y) < fmin_normal) +
                (z != 0.0 && fabs(z) < fmin_normal);
        const __half hx = fabs(x) < fmin_normal ? __float2half_rn(0.f) : __double2half(x);
        const __half hy = fabs(y) < fmin_normal ? __float2half_rn(0.f) : __double2half(y);
        const __half hz = fabs(z) < fmin_normal ? __float2half_rn(0.f) : __double2half(z);
        const int kc = m >> 3, e = m & 7;
        __half* pa = PA + (((panel * 3) * nkc + kc) * TF_ROWS + r) * 8 + e;
        pa[0] = hx;
        pa[(int64_t)nkc * TF_ROWS * 8] = hy;
        pa[(int64_t)2 * nkc * TF_ROWS * 8] = hz;
        __half* pb = PB + ((jt * nkc + kc) * TF_N + jj) * 8 + e;
        pb[0] = hx;
        pb[TF_J * 8] = hy;
        pb[2 * TF_J * 8] = hz;
        __half* pr = PR + (size_t)i * 3 * Mp + m;           // row-major image for the TMEM-resident operand
        pr[0] = hx;
        pr[Mp] = hy;
        pr[2 * Mp] = hz;
    }
    g = warp_sum(g);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tiny += __shfl_xor_sync(0xffffffffu, tiny, o);
    if (lane == 0) {
        const double alpha = 6.103515625e-05 * (1.0 + 9.765625e-4) / TF_EPS;
        const double sg = sqrt(g) + alpha * sqrt((double)tiny);
        G[i] = g; sG[i] = sg;
        CT[jt * 32 + jj] = __double2float_rd(0.5 * (1.0 - 1e-10) * g);
        CT[jt * 32 + 16 + jj] = __double2float_ru(sg);
    }
}

template <int NGROUPS, int STEP>   // NGROUPS epilogue groups of 4 warps; STEP columns per TMEM load round (4 or 8)
__global__ void __launch_bounds__((2 + 4 * NGROUPS) * 32, 1) rmsd_tf32_kernel(const TfParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nkc = p.Mp / 4;
    const uint32_t a_bytes = 3u * nkc * TF_ROWS * 16u;        // 1536 * Mp
    const uint32_t b_bytes = (uint32_t)nkc * TF_N * 16u;      // 192 * Mp
    float* smA = reinterpret_cast<float*>(smem_raw);
    unsigned char* smB = smem_raw + a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + a_bytes + (size_t)p.nb_stages * b_bytes);
    uint64_t* a_full = bars;                   // 1
    uint64_t* a_empty = bars + 1;              // 1
    uint64_t* b_full = bars + 2;               // nb_stages
    uint64_t* b_empty = b_full + TF_MAX_BSTAGES;
    uint64_t* t_full = b_empty + TF_MAX_BSTAGES;      // TF_NACC
    uint64_t* t_empty = t_full + TF_NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + TF_NACC);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < p.nb_stages; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int t = 0; t < TF_NACC; t++) { mbar_init(&t_full[t], 1); mbar_init(&t_empty[t], 4); }   // 4 warps per group
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TF_TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            int bs = 0; uint32_t bph = 0, aph = 0;
            for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
                const int4 w = p.items[it];
                mbar_wait(a_empty, aph ^ 1u);
                mbar_arrive_expect_tx(a_full, a_bytes);
                {   // the panel image is contiguous; copy it in <= 32 KB pieces
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.PA) + (size_t)w.x * a_bytes;
                    for (uint32_t off = 0; off < a_bytes; off += 32768u) {
                        const uint32_t n = (a_bytes - off < 32768u) ? (a_bytes - off) : 32768u;
                        bulk_g2s(reinterpret_cast<unsigned char*>(smA) + off, src + off, n, a_full);
                    }
                }
                aph ^= 1u;
                for (int t = 0; t < w.z; t++) {
                    mbar_wait(&b_empty[bs], bph ^ 1u);
                    mbar_arrive_expect_tx(&b_full[bs], b_bytes);
                    bulk_g2s(smB + (size_t)bs * b_bytes,
                             reinterpret_cast<const unsigned char*>(p.PB) + (size_t)(w.y + t) * b_bytes, b_bytes,
                             &b_full[bs]);
                    if (++bs == p.nb_stages) { bs = 0; bph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(TF_ROWS, TF_N);
            const uint32_t a_addr = smem_u32(smA);
            const uint32_t a_lbo = TF_ROWS * 16u, b_lbo = TF_N * 16u;      // next 16-byte K chunk
            const uint32_t a_comp = (uint32_t)nkc * TF_ROWS * 16u;         // next component image
            int bs = 0, acc = 0; uint32_t bph = 0, aph = 0, tph = 0;
            for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
                const int4 w = p.items[it];
                mbar_wait(a_full, aph);
                aph ^= 1u;
                for (int t = 0; t < w.z; t++) {
                    mbar_wait(&b_full[bs], bph);
                    mbar_wait(&t_empty[acc], tph ^ 1u);
                    tcgen05_fence_after();
                    const uint32_t b_addr = smem_u32(smB + (size_t)bs * b_bytes);
                    const uint32_t d0 = tmem_base + (uint32_t)acc * TF_ACC_COLS;
                    for (int kb = 0; kb < p.Mp / 8; kb++) {
                        const uint64_t bd = umma_desc_kmajor(b_addr + (uint32_t)kb * 2u * b_lbo, b_lbo, 128u);
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            const uint64_t ad =
                                umma_desc_kmajor(a_addr + (uint32_t)a * a_comp + (uint32_t)kb * 2u * a_lbo, a_lbo, 128u);
                            umma_tf32_ss(d0 + (uint32_t)a * TF_N, ad, bd, idesc, kb > 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&b_empty[bs]);        // smem stage reusable once these MMAs retire
                    umma_commit(&t_full[acc]);        // accumulators ready for the epilogue
                    if (++bs == p.nb_stages) { bs = 0; bph ^= 1u; }
                    if (++acc == TF_NACC) { acc = 0; tph ^= 1u; }
                }
                umma_commit(a_empty);                 // A panel reusable once every MMA of the item retired
            }
        }
    } else {
        // ===================== epilogue (8 warps, two groups alternating over tiles) =====================
        // Per pair:  lam = threshold eigenvalue lowered by the TF32 error bound (3 FP64 ops);
        //   fast path (FP32, no conversion): Samuelson's bound lambda_max <= sqrt(3) ||S~||_F, i.e. the
        //     pair is excluded if 3 * sum(S~^2) <= lam^2  — decided per half-tile with one warp vote;
        //   full path (FP64, branch-free): Budan-Fourier sign test on the key-matrix quartic at lam.
        const int ew = warp - 2;                      // 0 .. 4*NGROUPS-1
        const int grp = ew >> 2;                      // this group takes tiles with tile index % NGROUPS == grp
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
        const int row_in_panel = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        int acc = 0; uint32_t tph = 0;
        int64_t tile_seq = 0;                         // running tile counter of this CTA
        for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
            const int4 w = p.items[it];
            const int64_t i = (int64_t)w.x * TF_ROWS + row_in_panel;
            const TfRow row = tf32_row_consts(p.G[i], p.sG[i], p.e_thr);
            uint16_t* out_row = p.sim_bits16 + ((int64_t)w.w * CB + row_in_panel) * (2 * p.W);
            for (int t = 0; t < w.z; t++, tile_seq++) {
                const bool mine = ((int)(tile_seq % NGROUPS) == grp);
                if (mine) {
                    const int64_t j0 = (int64_t)(w.y + t) * TF_J;
                    // lanes 0..15 fetch G[j0+lane], lanes 16..31 sqrt(G)[j0+lane-16]; broadcast by shuffle later
                    const float gvf = tf32_col_term(p.G, p.sG, j0, lane);
                    mbar_wait(&t_full[acc], tph);
                    tcgen05_fence_after();
                    const uint32_t d0 = tmem_base + lane_addr + (uint32_t)acc * TF_ACC_COLS;
                    const uint32_t bits = tf32_epilogue_tile<STEP>(d0, gvf, row, p.G, p.sG, i, j0, p.N, lane, &t_empty[acc]);
                    if (i < p.N && (j0 >> 4) < 2 * p.W) out_row[j0 >> 4] = (uint16_t)bits;
                }
                if (++acc == TF_NACC) { acc = 0; tph ^= 1u; }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TF_TMEM_COLS);
}

}  // namespace tsc

extern "C" int64_t tsc_tf32_pa_floats(int64_t N, int32_t M) {
    const int64_t Mp = (M + 7) / 8 * 8, npanel = (N + tsc::TF_ROWS - 1) / tsc::TF_ROWS;
    return npanel * 3 * Mp * tsc::TF_ROWS;
}
extern "C" int64_t tsc_tf32_pb_floats(int64_t N, int32_t M) {
    const int64_t Mp = (M + 7) / 8 * 8, npanel = (N + tsc::TF_ROWS - 1) / tsc::TF_ROWS;
    return npanel * (tsc::TF_ROWS / tsc::TF_J) * Mp * tsc::TF_N;
}

extern "C" int tsc_pack_tf32(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M, float* PA,
                             float* PB, float* PR, double* G, double* sG, float* CT, void* stream) {
    using namespace tsc;
    if (N <= 0 || M <= 0) return 0;
    const int Mp = (M + 7) / 8 * 8;
    const int64_t rows_pad = (N + TF_ROWS - 1) / TF_ROWS * TF_ROWS;
    pack_tf32_kernel<<<(unsigned)((rows_pad + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, N, A, heavy_idx, M, Mp,
                                                                                     rows_pad, PA, PB, PR, G, sG, CT);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int64_t tsc_f16_operand_bytes(int64_t N, int32_t M) {       // size of each of PA, PB, PR
    const int64_t Mp = (M + 15) / 16 * 16, rows = (N + tsc::TF_ROWS - 1) / tsc::TF_ROWS * tsc::TF_ROWS;
    return rows * 3 * Mp * 2;
}

extern "C" int tsc_pack_f16(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M, void* PA,
                            void* PB, void* PR, double* G, double* sG, float* CT, void* stream) {
    using namespace tsc;
    if (N <= 0 || M <= 0) return 0;
    const int Mp = (M + 15) / 16 * 16;
    const int64_t rows_pad = (N + TF_ROWS - 1) / TF_ROWS * TF_ROWS;
    pack_f16_kernel<<<(unsigned)((rows_pad + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        S, N, A, heavy_idx, M, Mp, rows_pad, reinterpret_cast<__half*>(PA), reinterpret_cast<__half*>(PB),
        reinterpret_cast<__half*>(PR), G, sG, CT, 0);
    TSC_CHECK_LAUNCH();
    return 0;
}

// The same for conformers [row_begin, row_end) only (row_begin a multiple of 8; the last chunk should end at the
// padded row count ceil(N/128)*128 so that the padding rows are written too).
extern "C" int tsc_pack_f16_rows(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M, void* PA,
                                 void* PB, void* PR, double* G, double* sG, float* CT, int64_t row_begin,
                                 int64_t row_end, void* stream) {
    using namespace tsc;
    if (N <= 0 || M <= 0 || row_end <= row_begin) return 0;
    const int Mp = (M + 15) / 16 * 16;
    const int64_t rows_pad = (N + TF_ROWS - 1) / TF_ROWS * TF_ROWS;
    if (row_end > rows_pad) row_end = rows_pad;
    pack_f16_kernel<<<(unsigned)((row_end - row_begin + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        S, N, A, heavy_idx, M, Mp, row_end, reinterpret_cast<__half*>(PA), reinterpret_cast<__half*>(PB),
        reinterpret_cast<__half*>(PR), G, sG, CT, row_begin);
    TSC_CHECK_LAUNCH();
    return 0;
}

// items (n_items, 4) int32: {panel, first j tile, number of j tiles (<= 128 recommended), local row
// block (32-row units) of the panel's first row inside sim_bits}.
extern "C" int tsc_rmsd_sim_tf32(const float* PA, const float* PB, const double* G, const double* sG, int64_t N,
                                 int32_t M, const int32_t* items, int32_t n_items, double thr, uint32_t* sim_bits,
                                 int32_t grid_ctas, void* stream) {
    using namespace tsc;
    if (n_items <= 0 || N <= 0) return 0;
    TfParams p;
    p.PA = PA; p.PB = PB; p.G = G; p.sG = sG;
    p.items = reinterpret_cast<const int4*>(items);
    p.n_items = n_items;
    p.N = N;
    p.Mp = (M + 7) / 8 * 8;
    p.e_thr = (double)M * thr * thr * (1.0 + 1e-6);
    p.sim_bits16 = reinterpret_cast<uint16_t*>(sim_bits);
    p.W = num_blocks_padded(N);
    const size_t a_bytes = (size_t)1536 * p.Mp, b_bytes = (size_t)192 * p.Mp;
    const size_t budget = 227 * 1024 - 512;
    if (a_bytes + 2 * b_bytes > budget) return (int)cudaErrorInvalidValue;       // caller falls back to variant 0
    int nb = (int)((budget - a_bytes) / b_bytes);
    if (nb > TF_MAX_BSTAGES) nb = TF_MAX_BSTAGES;
    p.nb_stages = nb;
    const size_t smem = a_bytes + nb * b_bytes + 512;
    // grid_ctas < 0 selects an alternative epilogue configuration (tuning aid): -1 = 2 groups x 8
    // columns, -2 = 3 groups x 4 columns, -3 = 2 groups x 4 columns; default (>= 0) = TF_DEFAULT_CFG
    const int cfg = grid_ctas < 0 ? -grid_ctas : TF_DEFAULT_CFG;
    if (grid_ctas < 0) grid_ctas = 0;
    auto kern = cfg == 1 ? rmsd_tf32_kernel<2, 8> : cfg == 2 ? rmsd_tf32_kernel<3, 4> : rmsd_tf32_kernel<2, 4>;
    const int threads = cfg == 2 ? 448 : 320;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = grid_ctas > 0 ? grid_ctas : sms;
    if (grid > n_items) grid = n_items;
    kern<<<grid, threads, smem, (cudaStream_t)stream>>>(p);
    TSC_CHECK_LAUNCH();
    return 0;
}
