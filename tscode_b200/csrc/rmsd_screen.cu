// rmsd_screen.cu — the default all-pairs pre-screen of prune_conformers_rmsd: tcgen05 / TMEM, FP16 operands, FP32
// accumulation, COMPONENT-SEQUENTIAL accumulator buffers (tsc_pack_screen + tsc_rmsd_screen).
//
// Mathematics and output contract: a pair (i, j > i) can only be similar (rmsd_pruning.py:75) if lambda_max of the key
// matrix of its cross-covariance S exceeds lam_t = (G_i + G_j - M thr^2) / 2.  S is computed from FP16-rounded
// coordinates with FP32 accumulation; the operand error is bounded by ||S~ - S||_F <= eps sqrt(G_i)' sqrt(G_j)'
// (eps = 1.05e-3, tf32_common.cuh / pack below) and lam_t is lowered by sqrt(3) times that.  A pair is EXCLUDED only
// when an FP32 test with rigorous forward error bounds proves lambda_max below the lowered threshold (stage 1:
// Samuelson's bound sqrt(3) ||S~||_F; stage 2: sign test of the key-matrix quartic, tsc_math.cuh).  Whatever is not
// excluded is a candidate: bit set in sim_bits, (local row, j) appended to the candidate list, decided exactly in
// FP64 by rmsd_verify.cu.  Final bits and masks are therefore those of the FP64 variants (rmsd_sim.cu).
//
// Tiling — what bounds a tcgen05 kernel here is not arithmetic but (measured, tools/tmem_probe.py, profiles/):
//   * a 128 x N x 16 MMA costs max(N / 2, ~57) cycles, and MMAs that accumulate into the SAME TMEM region form a
//     dependent chain with ~144 cycles of latency, so the tensor pipe only runs at its rate with N >= 96 and >= 3
//     independent accumulators in flight;
//   * the nine covariance entries of a pair cost nine FP32 TMEM cells wherever they are put: 512 columns hold
//     128 x 56 pairs.  The first-generation kernel (rows = conformers i, columns = (a, b, j) for 16 conformers j,
//     three MMAs of N = 48 per K block, three 144-column buffers) was therefore stuck at N = 48: 38 % tensor-pipe
//     activity, 0.48 of the dense 16-bit peak on BASELINE configs[2].
// Here one accumulator buffer holds ONE ROW a of the covariances of a 128 x 32 tile:  D[i, (b, j)] = sum_m
// x_a(i, m) x_b(j, m), one MMA of N = 96 per K block, 96 columns; four buffers (384 columns) next to up to 5 K
// blocks of the stationary panel (120 columns).  The three rows of a tile are three independent chains; the MMA
// thread issues them skewed by a third of a tile, so that at any time three chains are in flight while the epilogue
// drains the fourth buffer.  The epilogue never sees the nine entries together: per pair it keeps T = S~^T S~ (six
// numbers, the sum of the outer products of the rows) in registers across the three buffers of a tile, and both
// exclusion tests work from T alone (f = tr T; quartic coefficients c2 = -2 f, c0 = 2 ||T||_F^2 - f^2, and
// |det S~| <= sqrt(det T + margin) in place of the signed determinant: tsc_math.cuh, quartic32_T_*).
//
// Roles (one persistent CTA per SM, 18 warps): warp 0 producer (bulk-TMA ring of B tiles, tail of the panel for
// K blocks beyond the 5 held in TMEM), warp 1 TMEM allocation + MMA issue (one elected thread), warps 2..17
// epilogue: TMEM lane quarter = warp % 4, and the four warps of a quarter split the 32 columns of a tile.
#include <cuda_fp16.h>
#include "tf32_common.cuh"

namespace tsc {

constexpr int SC_ROWS = 128;                  // conformers per panel (UMMA M, TMEM lanes)
constexpr int SC_J = 32;                      // conformers per B tile
constexpr int SC_N = 3 * SC_J;                // UMMA N = 96: (component b, conformer j)
constexpr int SC_NBUF = 4;                    // accumulator buffers of SC_N columns
constexpr int SC_KT = 5;                      // K blocks of the panel held in TMEM (3 * 8 * 5 = 120 columns)
constexpr int SC_ACC0 = 128;                  // first accumulator column
constexpr int SC_TMEM = 512;
constexpr int SC_MAX_BSTAGES = 12;
constexpr int SC_EPI_WARPS = 16;
constexpr int SC_COLS = SC_J / (SC_EPI_WARPS / 4);     // 8 columns of a tile per epilogue warp
constexpr int SC_Q = 128;                     // candidate queue entries per epilogue warp
constexpr int SC_THREADS = (2 + SC_EPI_WARPS) * 32;
static_assert(3 * 8 * SC_KT <= SC_ACC0 && SC_ACC0 + SC_NBUF * SC_N <= SC_TMEM, "TMEM budget");

struct ScParams {
    const unsigned char* PA;  // [panel][a][kc][128][16 B]     (only the chunks of K blocks >= SC_KT are read)
    const unsigned char* PB;  // [jtile][kc][96 = (b, j)][16 B]
    const unsigned char* PR;  // [row][a][kc][16 B]            row-major image for the TMEM-resident part of the panel
    const double* G;
    const double* sG;
    const float* CT;          // [jtile][64]: 32 x float_rd(hs G_j), 32 x float_ru(sqrt(G_j)')
    const int4* items;        // (panel, first j tile, j tile count, local 32-row block of the panel's first row)
    int n_items;
    int64_t N;
    int nkc;                  // 16-byte K chunks per conformer and component (two per K block)
    int nb_stages;
    double e_thr;
    uint8_t* sim_bits8;
    int64_t W;                // 32-bit words per sim row
    int2* cand;
    int64_t cand_stride;
};

// ---- pack: FP64 AoS -> FP16 operand images + exact G, widened sqrt(G), FP32 column terms -----------------------------
// FP16 keeps a 10-bit mantissa (relative rounding error 2^-11) but has a 5-bit exponent: |x| >= 65520 becomes inf
// (the pair then fails every exclusion test and is decided by the verify kernel), and values below the smallest
// normal 2^-14 are set to ZERO here, explicitly, so that no subnormal ever reaches the tensor core; their absolute
// error (<= 2^-14 each) enters the bound through sqrt(G)' = sqrt(G) + alpha sqrt(T), T = number of zeroed
// coordinates of the conformer, alpha = 2^-14 (1 + 2^-10) / eps:
//   ||S~ - S||_F <= 2 * 2^-11 (1 + 2^-11) sqrt(G_i G_j) + 2^-14 (1 + 2^-11) (sqrt(T_i G_j) + sqrt(G_i T_j))
//                <= eps sqrt(G_i)' sqrt(G_j)'      (eps = 1.05e-3 also leaves 7 % for the FP32 accumulation)
__global__ void __launch_bounds__(256) pack_screen_kernel(const double* __restrict__ S, int64_t N, int A,
                                                          const int32_t* __restrict__ heavy_idx, int M, int Mp,
                                                          int64_t n_rows_end, __half* __restrict__ PA,
                                                          __half* __restrict__ PB, __half* __restrict__ PR,
                                                          double* __restrict__ G, double* __restrict__ sG,
                                                          float* __restrict__ CT, int64_t row_begin) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = row_begin + (int64_t)blockIdx.x * 8 + warp;      // one warp per conformer (incl. padding rows)
    if (i >= n_rows_end) return;
    const bool live = i < N;
    const double* src = S + (live ? i : 0) * (int64_t)A * 3;
    const int nkc = Mp / 8;
    const int64_t panel = i / SC_ROWS, r = i % SC_ROWS;
    const int64_t jt = i / SC_J, jj = i % SC_J;
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
        __half* pa = PA + (((panel * 3) * nkc + kc) * SC_ROWS + r) * 8 + e;
        pa[0] = hx;
        pa[(int64_t)nkc * SC_ROWS * 8] = hy;
        pa[(int64_t)2 * nkc * SC_ROWS * 8] = hz;
        __half* pb = PB + ((jt * nkc + kc) * SC_N + jj) * 8 + e;
        pb[0] = hx;
        pb[SC_J * 8] = hy;
        pb[2 * SC_J * 8] = hz;
        __half* pr = PR + (size_t)i * 3 * Mp + m;
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
        CT[jt * (2 * SC_J) + jj] = __double2float_rd(0.5 * (1.0 - 1e-10) * g);
        CT[jt * (2 * SC_J) + SC_J + jj] = __double2float_ru(sg);
    }
}

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__global__ void __launch_bounds__(SC_THREADS, 1) rmsd_screen_kernel(const ScParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nkc = p.nkc;
    const int nkb = nkc / 2;                                   // K blocks (one MMA per component each)
    const int KT = nkb < SC_KT ? nkb : SC_KT;                  // K blocks of the panel held in TMEM
    const int tail_kc = nkc - 2 * KT;                          // chunks of the panel kept in shared memory
    const uint32_t tail_bytes = (uint32_t)tail_kc * SC_ROWS * 16u;      // per component
    const uint32_t b_bytes = (uint32_t)nkc * SC_N * 16u;
    unsigned char* smA = smem_raw;                             // [a][tail_kc][128][16 B]
    unsigned char* smB = smem_raw + 3u * tail_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)p.nb_stages * b_bytes);
    uint64_t* at_full = bars;                  // tail of the panel landed (TMA)
    uint64_t* am_full = bars + 1;              // panel rows stored to TMEM (16 epilogue warps)
    uint64_t* a_empty = bars + 2;              // all MMAs of the item retired
    uint64_t* b_full = bars + 3;
    uint64_t* b_empty = b_full + SC_MAX_BSTAGES;               // three arrivals: one per row unit of the tile
    uint64_t* t_full = b_empty + SC_MAX_BSTAGES;
    uint64_t* t_empty = t_full + SC_NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + SC_NBUF);
    int2* cand_q = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(bars) + 512);     // [epilogue warp][SC_Q]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(at_full, 1);
        mbar_init(am_full, SC_EPI_WARPS);
        mbar_init(a_empty, 1);
        for (int s = 0; s < p.nb_stages; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 3); }
        for (int t = 0; t < SC_NBUF; t++) { mbar_init(&t_full[t], 1); mbar_init(&t_empty[t], SC_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, SC_TMEM);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer: warp-uniform control flow, one elected lane issues the copies =====================
        int bs = 0; uint32_t bph = 0, aph = 0;
        for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
            const int4 w = p.items[it];
            if (w.z == 0) break;              // empty item: this CTA's list is exhausted (_host.build_screen_items)
            if (tail_kc > 0) {
                mbar_wait(a_empty, aph ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(at_full, 3u * tail_bytes);
                    const unsigned char* src = p.PA + (size_t)w.x * 3u * nkc * SC_ROWS * 16u;
                    for (int a = 0; a < 3; a++)
                        for (uint32_t off = 0; off < tail_bytes; off += 32768u)          // bulk copies of <= 32 KB
                            bulk_g2s(smA + (size_t)a * tail_bytes + off,
                                     src + ((size_t)a * nkc + 2 * KT) * SC_ROWS * 16u + off,
                                     tail_bytes - off < 32768u ? tail_bytes - off : 32768u, at_full);
                }
                __syncwarp();
                aph ^= 1u;
            }
            for (int t = 0; t < w.z; t++) {
                mbar_wait(&b_empty[bs], bph ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&b_full[bs], b_bytes);
                    for (uint32_t off = 0; off < b_bytes; off += 32768u)
                        bulk_g2s(smB + (size_t)bs * b_bytes + off, p.PB + (size_t)(w.y + t) * b_bytes + off,
                                 b_bytes - off < 32768u ? b_bytes - off : 32768u, &b_full[bs]);
                }
                __syncwarp();
                if (++bs == p.nb_stages) { bs = 0; bph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issue: one elected thread runs the whole schedule =====================
        // Unit = (tile, row a of the covariances) = nkb MMAs into accumulator buffer (unit number mod 4).  Chain a
        // ("lane" l = a) works through its units of consecutive tiles; per round one MMA of every chain is issued,
        // chain l lagging chain 0 by l * delta K blocks, so that units complete — and buffers are needed — evenly
        // spaced in time, in the order the epilogue consumes them.
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(SC_ROWS, SC_N);
            const uint32_t a_lbo = SC_ROWS * 16u, b_lbo = SC_N * 16u;
            const uint64_t bd0 = umma_desc_kmajor(smem_u32(smB), b_lbo, 128u);       // stage 0, K block 0
            const uint64_t ad0 = umma_desc_kmajor(smem_u32(smA), a_lbo, 128u);       // panel tail, component x
            const uint64_t bd_step = (2u * b_lbo) >> 4, ad_step = (2u * a_lbo) >> 4; // start-address field, 16-byte units
            const uint64_t b_stage_step = b_bytes >> 4, a_comp_step = tail_bytes >> 4;
            const uint32_t a_stride = 8u * KT;                                       // TMEM columns per component
            const int delta = nkb >= 2 ? (nkb + 2) / 3 : 0;                          // 2 delta <= nkb: no wait can deadlock
            int st[3] = {0, 0, 0};                                                   // B stage of the chain's current tile
            uint32_t sph[3] = {0, 0, 0};
            uint32_t us[3] = {0, 1, 2};                                              // unit number of the chain's current unit
            uint32_t aph = 0;
            for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
                const int4 w = p.items[it];
                if (w.z == 0) break;
                mbar_wait(am_full, aph);
                if (tail_kc > 0) mbar_wait(at_full, aph);
                aph ^= 1u;
                tcgen05_fence_after();
                const int total = w.z * nkb;
                int k[3] = {0, 0, 0};
                for (int r = 0; r < total + 2 * delta; r++) {
#pragma unroll
                    for (int l = 0; l < 3; l++) {
                        const int kap = r - l * delta;
                        if (kap < 0 || kap >= total) continue;
                        const uint32_t buf = us[l] & 3u;
                        if (k[l] == 0) {
                            mbar_wait(&b_full[st[l]], sph[l]);
                            mbar_wait(&t_empty[buf], ((us[l] >> 2) & 1u) ^ 1u);
                            tcgen05_fence_after();
                        }
                        const uint32_t d = tmem_base + SC_ACC0 + buf * SC_N;
                        const uint64_t bd = bd0 + (uint64_t)st[l] * b_stage_step + (uint64_t)k[l] * bd_step;
                        if (k[l] < KT) {
                            const uint32_t at = tmem_base + (uint32_t)l * a_stride + 8u * (uint32_t)k[l];
                            if (k[l] == 0) umma_tf32_ts_c<false, true>(d, at, bd, idesc);
                            else umma_tf32_ts_c<true, true>(d, at, bd, idesc);
                        } else {                                   // K blocks whose panel block is in shared memory
                            const uint64_t ad = ad0 + (uint64_t)l * a_comp_step + (uint64_t)(k[l] - KT) * ad_step;
                            umma_tf32_ss_c<true, true>(d, ad, bd, idesc);
                        }
                        if (++k[l] == nkb) {
                            umma_commit(&t_full[buf]);             // this row of the tile is ready for the epilogue
                            umma_commit(&b_empty[st[l]]);          // (the stage is free after the third row's commit)
                            k[l] = 0;
                            us[l] += 3u;
                            if (++st[l] == p.nb_stages) { st[l] = 0; sph[l] ^= 1u; }
                        }
                    }
                }
                umma_commit(a_empty);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: 16 warps =====================
        const int ew = warp - 2;
        const int quad = warp & 3;                               // TMEM lane quarter this warp may access
        const int part = ew >> 2;                                // which 8 columns of every tile
        const int row_in_panel = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int c0 = part * SC_COLS;
        uint32_t eph = 0, us = 0;                                // unit number (all units of this CTA, in order)
        // Candidates (rare) go to a per-warp queue in shared memory and reach the global list in bursts with one
        // reservation per burst.
        int2* my_q = cand_q + (size_t)ew * SC_Q;
        int qn = 0;
        auto flush_q = [&]() {
            __syncwarp();
            int64_t base = 0;
            if (lane == 0) base = list_reserve(&p.cand[0].x, qn, p.cand_stride - 1);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (int k = lane; k < qn; k += 32)
                if (base + k < p.cand_stride - 1) p.cand[1 + base + k] = my_q[k];
            __syncwarp();
            qn = 0;
        };
        for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
            const int4 w = p.items[it];
            if (w.z == 0) break;
            const int64_t i = (int64_t)w.x * SC_ROWS + row_in_panel;
            // ---- this thread's row of the panel -> TMEM; the row's 3 * 2 * KT 16-byte chunks are dealt to the four parts
            mbar_wait(a_empty, eph ^ 1u);
            eph ^= 1u;
            tcgen05_fence_after();
            {
                const float4* src = reinterpret_cast<const float4*>(p.PR + (size_t)i * 3 * nkc * 16);
                for (int q = part; q < 3 * 2 * KT; q += SC_EPI_WARPS / 4) {
                    const int a = q / (2 * KT), wi = q - a * 2 * KT;
                    const float4 v = src[a * nkc + wi];
                    tmem_st_x4(tmem_base + lane_addr + (uint32_t)(a * 8 * KT + 4 * wi), __float_as_uint(v.x),
                               __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
                }
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(am_full);
            }
            const TfRow row = tf32_row_consts(p.G[i], p.sG[i], p.e_thr);
            const unsigned long long Af2 = OpsF2::bc(row.Af), nCf2 = OpsF2::bc(-row.Cf);
            uint8_t* out_row = p.sim_bits8 + ((int64_t)w.w * CB + row_in_panel) * (4 * p.W);
            for (int t = 0; t < w.z; t++) {
                const int64_t j0 = (int64_t)(w.y + t) * SC_J + c0;            // first of this warp's 8 columns
                unsigned long long T[6][4];                                  // T = S~^T S~, two columns per register pair
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const uint32_t buf = us & 3u;
                    mbar_wait(&t_full[buf], (us >> 2) & 1u);
                    tcgen05_fence_after();
                    const uint32_t d0 = tmem_base + lane_addr + SC_ACC0 + buf * SC_N + (uint32_t)c0;
                    uint32_t r[24];                                          // row a of S~ for 8 columns: [b][column]
                    tmem_ld_x8_raw(d0, &r[0]);
                    tmem_ld_x8_raw(d0 + SC_J, &r[8]);
                    tmem_ld_x8_raw(d0 + 2 * SC_J, &r[16]);
                    tmem_wait_bind24(r);
                    tcgen05_fence_before();                                  // the buffer goes back to the MMA thread
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[buf]);
                    us++;
#pragma unroll
                    for (int cp = 0; cp < 4; cp++) {
                        const unsigned long long v0 = OpsF2::pack(r[2 * cp], r[2 * cp + 1]);
                        const unsigned long long v1 = OpsF2::pack(r[8 + 2 * cp], r[8 + 2 * cp + 1]);
                        const unsigned long long v2 = OpsF2::pack(r[16 + 2 * cp], r[16 + 2 * cp + 1]);
                        if (a == 0) {
                            T[0][cp] = OpsF2::mul(v0, v0); T[1][cp] = OpsF2::mul(v1, v1); T[2][cp] = OpsF2::mul(v2, v2);
                            T[3][cp] = OpsF2::mul(v0, v1); T[4][cp] = OpsF2::mul(v0, v2); T[5][cp] = OpsF2::mul(v1, v2);
                        } else {
                            T[0][cp] = OpsF2::fma(v0, v0, T[0][cp]); T[1][cp] = OpsF2::fma(v1, v1, T[1][cp]);
                            T[2][cp] = OpsF2::fma(v2, v2, T[2][cp]); T[3][cp] = OpsF2::fma(v0, v1, T[3][cp]);
                            T[4][cp] = OpsF2::fma(v0, v2, T[4][cp]); T[5][cp] = OpsF2::fma(v1, v2, T[5][cp]);
                        }
                    }
                }
                // ---- the tile's verdicts for this warp's 8 columns: column terms B_j (rounded down), D_j (rounded up)
                const float2* ct = reinterpret_cast<const float2*>(p.CT + (int64_t)(w.y + t) * (2 * SC_J) + c0);
                float2 Bj[4], Dj[4];
#pragma unroll
                for (int cp = 0; cp < 4; cp++) { Bj[cp] = __ldg(ct + cp); Dj[cp] = __ldg(ct + SC_J / 2 + cp); }
                uint32_t near = 0;                                           // bit c: pair of column c0 + c not excluded
#pragma unroll
                for (int cp = 0; cp < 4; cp++) {
                    const unsigned long long tt[6] = {T[0][cp], T[1][cp], T[2][cp], T[3][cp], T[4][cp], T[5][cp]};
                    // stage 1, Samuelson: lambda_max <= sqrt(3) ||S~||_F.  With lf = A_i + B_j - C_i D_j (a lower bound
                    // of the lowered threshold eigenvalue up to a factor 1 - 1e-6, folded into the constant together with
                    // the rounding of f and of the products: 3 (1 + 9 u) / (1 - 1e-6)^2 < 3.00004) the pair is excluded
                    // iff lf > 0 and 3.00004 f - lf^2 < 0, read off the sign bits (NaN / inf: not excluded).
                    const unsigned long long f2 = OpsF2::add(OpsF2::add(tt[0], tt[1]), tt[2]);
                    const unsigned long long ab2 = OpsF2::add(Af2, OpsF2::pack(__float_as_uint(Bj[cp].x), __float_as_uint(Bj[cp].y)));
                    const unsigned long long lf2 = OpsF2::fma(nCf2, OpsF2::pack(__float_as_uint(Dj[cp].x), __float_as_uint(Dj[cp].y)), ab2);
                    const unsigned long long t2 = OpsF2::fma(OpsF2::bc(3.00004f), f2, OpsF2::mul(OpsF2::mul(lf2, lf2), OpsF2::bc(-1.0f)));
                    float lf[2], ab[2], tv[2];
                    OpsF2::unpack(lf2, lf[0], lf[1]);
                    OpsF2::unpack(ab2, ab[0], ab[1]);
                    OpsF2::unpack(t2, tv[0], tv[1]);
                    uint32_t und = 0;
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        und |= (((__float_as_uint(tv[h]) & ~__float_as_uint(lf[h])) >> 31) ^ 1u) << h;
                    // stage 2, FP32 sign test of the key-matrix quartic from T, for a column pair with an undecided lane
                    if (__any_sync(0xffffffffu, und)) {
                        float lam[2];
#pragma unroll
                        for (int h = 0; h < 2; h++) lam[h] = fmaf(-2e-7f, fabsf(ab[h]) + fabsf(lf[h]), lf[h]);
                        const unsigned long long lam2 = OpsF2::pack(__float_as_uint(lam[0]), __float_as_uint(lam[1]));
                        unsigned long long fq, cc0, da, p0, p1, p2, m0, m1, m2;
                        quartic32_T_coeffs<OpsF2>(tt, fq, cc0, da);
                        float da0, da1;
                        OpsF2::unpack(da, da0, da1);
                        const unsigned long long d2 = OpsF2::pack(__float_as_uint(sqrt_approx(fmaxf(da0, 0.f))),
                                                                  __float_as_uint(sqrt_approx(fmaxf(da1, 0.f))));
                        quartic32_T_values<OpsF2>(fq, cc0, d2, lam2, p0, p1, p2);
                        quartic32_margins<OpsF2>(p0, p1, p2, fq, lam2, m0, m1, m2);
                        float a1[2], b0[2], b1[2], b2[2];
                        OpsF2::unpack(p1, a1[0], a1[1]);
                        OpsF2::unpack(m0, b0[0], b0[1]);
                        OpsF2::unpack(m1, b1[0], b1[1]);
                        OpsF2::unpack(m2, b2[0], b2[1]);
#pragma unroll
                        for (int h = 0; h < 2; h++)
                            if (quartic32_decide(lam[h], a1[h], b0[h], b1[h], b2[h])) und &= ~(1u << h);
                    }
                    near |= und << (2 * cp);
                }
                // validity: j > i, j < N   (rows i >= N are not stored)
                uint32_t valid = 0xffu;
                if (j0 + SC_COLS - 1 >= p.N) valid = (j0 >= p.N) ? 0u : (0xffu >> (j0 + SC_COLS - p.N));
                if (j0 <= i) valid &= (i - j0 >= SC_COLS - 1) ? 0u : (0xffu << (i - j0 + 1));
                const uint32_t bits = (i < p.N) ? (near & valid & 0xffu) : 0u;
                if (p.cand && __any_sync(0xffffffffu, bits != 0u)) {          // append (local row, j) of every bit set
                    uint32_t bb = bits;
                    const int cnt = __popc(bb);
                    int pre = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, pre, o);
                        if (lane >= o) pre += v;
                    }
                    const int total = __shfl_sync(0xffffffffu, pre, 31);
                    int at = pre - cnt;
                    const int32_t lrow = w.w * CB + row_in_panel;
                    if (qn + total > SC_Q) flush_q();
                    if (total > SC_Q) {                                      // a dense tile: straight to the global list
                        int64_t base = 0;
                        if (lane == 0) base = list_reserve(&p.cand[0].x, total, p.cand_stride - 1);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        while (bb) {
                            const int b = __ffs(bb) - 1;
                            bb &= bb - 1;
                            if (base + at < p.cand_stride - 1) p.cand[1 + base + at] = make_int2(lrow, (int32_t)(j0 + b));
                            at++;
                        }
                    } else {
                        while (bb) {
                            const int b = __ffs(bb) - 1;
                            bb &= bb - 1;
                            my_q[qn + at++] = make_int2(lrow, (int32_t)(j0 + b));
                        }
                        qn += total;
                    }
                }
                if (i < p.N && (j0 >> 5) < p.W) out_row[j0 >> 3] = (uint8_t)bits;
            }
        }
        if (p.cand && qn) flush_q();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, SC_TMEM);
}

}  // namespace tsc

// size in bytes of each of the three operand images (PA, PB, PR)
extern "C" int64_t tsc_screen_operand_bytes(int64_t N, int32_t M) {
    const int64_t Mp = (M + 15) / 16 * 16, rows = (N + tsc::SC_ROWS - 1) / tsc::SC_ROWS * tsc::SC_ROWS;
    return rows * 3 * Mp * 2;
}
extern "C" int64_t tsc_screen_ct_floats(int64_t N) {
    return (N + tsc::SC_ROWS - 1) / tsc::SC_ROWS * tsc::SC_ROWS * 2;
}
// largest number of heavy atoms the screen takes (panel tail + three B stages must fit in shared memory); above it
// the FP64 tensor-core variant runs (tsc_rmsd_sim_tiles)
extern "C" int32_t tsc_screen_max_atoms(void) {
    int best = 0;
    for (int M = 16; M <= 1024; M += 16) {
        const int nkc = M / 8, nkb = nkc / 2, KT = nkb < tsc::SC_KT ? nkb : tsc::SC_KT;
        const size_t a_bytes = (size_t)3 * (nkc - 2 * KT) * tsc::SC_ROWS * 16, b_bytes = (size_t)nkc * tsc::SC_N * 16;
        const size_t budget = 227 * 1024 - 512 - (size_t)tsc::SC_EPI_WARPS * tsc::SC_Q * sizeof(int2);
        if (a_bytes + 3 * b_bytes <= budget) best = M;
    }
    return best;
}

// rows [row_begin, row_end) only (row_begin a multiple of 8; the last chunk should end at the padded row count
// ceil(N/128)*128 so that the padding rows are written too); row_end <= 0: all rows
extern "C" int tsc_pack_screen(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M, void* PA,
                               void* PB, void* PR, double* G, double* sG, float* CT, int64_t row_begin, int64_t row_end,
                               void* stream) {
    using namespace tsc;
    if (N <= 0 || M <= 0) return 0;
    const int Mp = (M + 15) / 16 * 16;
    const int64_t rows_pad = (N + SC_ROWS - 1) / SC_ROWS * SC_ROWS;
    if (row_end <= 0 || row_end > rows_pad) row_end = rows_pad;
    if (row_begin < 0) row_begin = 0;
    if (row_end <= row_begin) return 0;
    pack_screen_kernel<<<(unsigned)((row_end - row_begin + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        S, N, A, heavy_idx, M, Mp, row_end, reinterpret_cast<__half*>(PA), reinterpret_cast<__half*>(PB),
        reinterpret_cast<__half*>(PR), G, sG, CT, row_begin);
    TSC_CHECK_LAUNCH();
    return 0;
}

// items (n_items, 4) int32: {panel, first j tile (32 conformers), number of j tiles, local row block (32-row units)
// of the panel's first row inside sim_bits}; dealt round-robin to the CTAs of the persistent grid, an item with
// count 0 ends a CTA's list (_host.build_screen_items).  cand_list: header + (local row, j) entries, or NULL.
extern "C" int tsc_rmsd_screen(const void* PA, const void* PB, const void* PR, const double* G, const double* sG,
                               const float* CT, int64_t N, int32_t M, const int32_t* items, int32_t n_items, double thr,
                               uint32_t* sim_bits, int32_t* cand_list, int64_t cand_stride, int32_t grid_ctas, void* stream) {
    using namespace tsc;
    if (n_items <= 0 || N <= 0) return 0;
    ScParams p;
    p.PA = reinterpret_cast<const unsigned char*>(PA);
    p.PB = reinterpret_cast<const unsigned char*>(PB);
    p.PR = reinterpret_cast<const unsigned char*>(PR);
    p.G = G; p.sG = sG; p.CT = CT;
    p.items = reinterpret_cast<const int4*>(items);
    p.n_items = n_items;
    p.N = N;
    p.nkc = (M + 15) / 16 * 2;                                  // atoms padded to whole K blocks
    p.e_thr = (double)M * thr * thr * (1.0 + 1e-6);
    p.sim_bits8 = reinterpret_cast<uint8_t*>(sim_bits);
    p.W = num_blocks_padded(N);
    p.cand = reinterpret_cast<int2*>(cand_list);
    p.cand_stride = cand_stride;
    const int nkb = p.nkc / 2, KT = nkb < SC_KT ? nkb : SC_KT;
    const size_t a_bytes = (size_t)3 * (p.nkc - 2 * KT) * SC_ROWS * 16, b_bytes = (size_t)p.nkc * SC_N * 16;
    const size_t q_bytes = (size_t)SC_EPI_WARPS * SC_Q * sizeof(int2);
    const size_t budget = 227 * 1024 - 512 - q_bytes;
    if (a_bytes + 3 * b_bytes > budget) return (int)cudaErrorInvalidValue;      // more atoms than tsc_screen_max_atoms()
    int nb = (int)((budget - a_bytes) / b_bytes);
    if (nb > SC_MAX_BSTAGES) nb = SC_MAX_BSTAGES;
    p.nb_stages = nb;
    const size_t smem = a_bytes + nb * b_bytes + 512 + q_bytes;
    cudaError_t e = cudaFuncSetAttribute(rmsd_screen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = grid_ctas > 0 ? grid_ctas : sms;
    if (grid > n_items) grid = n_items;
    rmsd_screen_kernel<<<grid, SC_THREADS, smem, (cudaStream_t)stream>>>(p);
    TSC_CHECK_LAUNCH();
    return 0;
}
