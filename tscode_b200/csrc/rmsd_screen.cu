// rmsd_screen.cu — the default all-pairs pre-screen of prune_conformers_rmsd: tcgen05 / TMEM, FP16 operands, FP32
// accumulation, COMPONENT-SEQUENTIAL accumulator buffers (tsc_pack_screen + tsc_rmsd_screen).
//
// Mathematics and output contract: a pair (i, j > i) can only be similar (rmsd_pruning.py:75) if lambda_max of the key
// matrix of its cross-covariance S exceeds lam_t = (G_i + G_j - M thr^2) / 2.  S is computed from FP16-rounded
// coordinates with FP32 accumulation; the operand error is bounded by ||S~ - S||_F <= eps sqrt(G_i)' sqrt(G_j)'
// (eps = 1.05e-3, screen_common.cuh / pack below) and lam_t is lowered by sqrt(3) times that.  A pair is EXCLUDED only
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
// Here one accumulator buffer holds ONE ROW a of the covariances of a 128 x J tile:  D[i, (b, j)] = sum_m
// x_a(i, m) x_b(j, m), one MMA of N = 3 J per K block.  The three rows of a tile are three independent chains, each
// issued by its own thread, so that several chains are in flight while the epilogue drains a buffer.  The epilogue never
// sees the nine entries together: per pair it keeps f = ||S~||_F^2 (mode 0) or T = S~^T S~ (six numbers, the sum of the
// outer products of the rows; modes 1 / 2) in registers across the three buffers of a tile, and both exclusion tests
// work from those alone (f = tr T; quartic coefficients c2 = -2 f, c0 = 2 ||T||_F^2 - f^2, and
// |det S~| <= sqrt(det T + margin) in place of the signed determinant: tsc_math.cuh, quartic32_T_*).
// Three tilings (template parameter J, table below): J = 48 with THREE buffers, one per row a, so that every row's MMAs
// may start a whole tile ahead of the epilogue (mode 0: with two buffers of J = 64 the per-tile final stage outlasted
// the look-ahead and the tensor pipe idled, profiles/r02_ncu_kernels.md); J = 32 with four rotating buffers for the
// T-accumulating modes (six numbers per pair: 8 columns per thread); J = 64 kept for comparison.
//
// Roles (one persistent CTA per SM): warp 0 producer (bulk-TMA ring of B tiles, tail of the panel for the K blocks
// beyond those held in TMEM), warps 1..3 MMA issue (one elected thread each, one per row of the covariances; warp 1
// also owns the TMEM allocation), the remaining 12 (J = 48) or 16 warps epilogue: TMEM lane quarter = warp % 4, and the
// warps of a quarter split the columns of a tile.
#include <cuda_fp16.h>
#include "screen_common.cuh"

namespace tsc {

constexpr int SC_ROWS = 128;                  // conformers per panel (UMMA M, TMEM lanes)
// tile width J (conformers per B tile) is a template parameter:
//   32: UMMA N =  96, four accumulator buffers, 16 epilogue warps x  8 columns  (modes 1, 2: six numbers per pair)
//   48: UMMA N = 144, three buffers = one per row a,  12 epilogue warps x 16 columns  (mode 0)
//   64: UMMA N = 192, two buffers,  16 epilogue warps x 16 columns  (mode 0, kept for comparison: tsc_rmsd_screen mode 3)
constexpr int SC_TMEM = 512;
constexpr int SC_MAX_BSTAGES = 12;
constexpr int SC_MAX_EPI_WARPS = 16;
constexpr int SC_Q = 128;                     // candidate queue entries per epilogue warp
constexpr int SC_CT_BYTES = 128;              // per epilogue warp: staged column terms of a tile (16 x B_j, 16 x D_j)
constexpr int SC_MMA_WARPS = 3;                // one issuing thread per row a of the covariances ("chain")
// K blocks of the panel held in TMEM (8 columns per block and component) and first accumulator column
__host__ __device__ constexpr int sc_kt(int J) { return J == 48 ? 3 : 5; }
__host__ __device__ constexpr int sc_acc0(int J) { return J == 48 ? 80 : 128; }
__host__ __device__ constexpr int sc_nbuf(int J) { return J == 32 ? 4 : J == 48 ? 3 : 2; }
__host__ __device__ constexpr int sc_epi_warps(int J) { return J == 48 ? 12 : 16; }
__host__ __device__ constexpr int sc_threads(int J) { return (1 + SC_MMA_WARPS + sc_epi_warps(J)) * 32; }
static_assert(3 * 8 * sc_kt(32) <= sc_acc0(32) && sc_acc0(32) + 4 * 96 <= SC_TMEM && sc_acc0(64) + 2 * 192 <= SC_TMEM &&
              3 * 8 * sc_kt(48) <= sc_acc0(48) && sc_acc0(48) + 3 * 144 <= SC_TMEM, "TMEM budget");

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
    int pace;                 // cycles between two MMAs of one chain (0 = unpaced), see the MMA warps
    double e_thr;
    uint8_t* sim_bits8;
    int64_t W;                // 32-bit words per sim row
    int2* cand;
    int64_t cand_stride;
    int scaled;               // the B-side image holds scaled coordinates (ScFrame): the quartic stage un-scales T first
    float inv_tt[6];          // 1 / (t_b t_c) for T's entries (00, 11, 22, 01, 02, 12)
#ifdef TSC_SCREEN_TRACE
    long long* trace;         // measurement build only (tools/probes): clock64 stamps of CTA 0, see tools/screen_trace.py
#endif
};

#ifdef TSC_SCREEN_TRACE
#define SC_STAMP(cond, idx) do { if ((cond) && p.trace && blockIdx.x == 0) p.trace[(idx)] = clock64(); } while (0)
#else
#define SC_STAMP(cond, idx) do { } while (0)
#endif

// ---- pack: FP64 AoS -> FP16 operand images + exact G, widened sqrt(G), FP32 column terms -----------------------------
// FP16 keeps a 10-bit mantissa (relative rounding error 2^-11) but has a 5-bit exponent: |x| >= 65520 becomes inf
// (the pair then fails every exclusion test and is decided by the verify kernel), and values below the smallest
// normal 2^-14 are set to ZERO here, explicitly, so that no subnormal ever reaches the tensor core; their absolute
// error (<= 2^-14 each) enters the bound through sqrt(G)' = sqrt(G) + alpha sqrt(T), T = number of zeroed
// coordinates of the conformer, alpha = 2^-14 (1 + 2^-10) / eps:
//   ||S~ - S||_F <= 2 * 2^-11 (1 + 2^-11) sqrt(G_i G_j) + 2^-14 (1 + 2^-11) (sqrt(T_i G_j) + sqrt(G_i T_j))
//                <= eps sqrt(G_i)' sqrt(G_j)'      (eps = 1.05e-3 also leaves 7 % for the FP32 accumulation)
// Images: PA [panel][a][kc][128][16 B] and PR [row][a][kc][16 B] (stationary panel: shared-memory tail / TMEM part),
// PB [j tile of J conformers][kc][(b, j) = 3 J rows][16 B] (canonical no-swizzle K-major core-matrix order: a plain
// bulk copy of a tile is the shared-memory operand of the MMAs), CT [j tile][2 J] column terms.
//
// Frame and column weights (ScFrame).  Samuelson's bound sqrt(3) ||S||_F >= s1 + s2 + s3 is sharp only for isotropic
// covariances; for an elongated or planar molecule it excluded nothing.  Two exact facts fix that at no cost to the
// kernel: (1) rotating BOTH conformers by the same orthogonal Q leaves the singular values of S, hence lambda_max,
// unchanged; (2) the nuclear norm is at most the sum of the column norms, and by Cauchy-Schwarz, for any weights
// w_b > 0 with sum_b 1 / w_b <= 1,
//     s1 + s2 + s3 <= sum_b ||col_b(S)|| <= sqrt(sum_b w_b ||col_b(S)||^2) = sqrt(3) || S diag(t) ||_F,  t_b = sqrt(w_b / 3).
// S diag(t) is the covariance of conformer i with conformer j's coordinates scaled by t_b along axis b — what the MMAs
// compute if the B-side image holds the scaled coordinates.  In the principal-axes frame of the molecule (Q from the
// second-moment tensor of the first structure: there S is nearly diagonal, the column norms are nearly the singular
// values) with w_b = (l1 + l2 + l3) / l_b the bound is within 1-3 % of lambda_max for elongated and planar ensembles
// (tools/weighted_bound_probe.py), as Samuelson's is for isotropic ones, where Q = I, t = 1 reproduce it exactly.  Soundness
// needs only Q orthogonal and sum 1 / w_b <= 1 (checked by tsc_pack_screen), never that the frame is well chosen.
// The A-side images hold the rotated, unscaled coordinates.  Error bound: per component, |E_ab| <= 2^-10 (1 + ..)
// ||p_a|| ||q^_b|| (relative rounding is scale invariant), so ||S^~ - S^||_F <= eps sqrt(G_i)' sqrt(G^_j)' with G^_j the
// squared norm of the SCALED column conformer; the modes that recover T = S^T S from T^ = diag(t) T diag(t) for the
// quartic test (entries divided by t_b t_c) get ||S~ - S||_F <= eps sqrt(G_i)' sqrt(G_j)' the same way.  The column
// term D_j is the larger of the two widened roots, valid for both.
struct ScFrame {
    double q[9];              // rows = the frame's axes: x' = q[0] x + q[1] y + q[2] z, ...
    double t[3];              // B-side scale per axis
    double inv_tmin;          // 1 / min(t): worst growth of a flushed (absolute-error) coordinate when T is un-scaled
};
template <int J>
__global__ void __launch_bounds__(256) pack_screen_kernel(const double* __restrict__ S, int64_t N, int A,
                                                          const int32_t* __restrict__ heavy_idx, int M, int Mp,
                                                          int64_t n_rows_end, __half* __restrict__ PA,
                                                          __half* __restrict__ PB, __half* __restrict__ PR,
                                                          double* __restrict__ G, double* __restrict__ sG,
                                                          float* __restrict__ CT, int64_t row_begin, const ScFrame fr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = row_begin + (int64_t)blockIdx.x * 8 + warp;      // one warp per conformer (incl. padding rows)
    if (i >= n_rows_end) return;
    const bool live = i < N;
    const double* src = S + (live ? i : 0) * (int64_t)A * 3;
    const int nkc = Mp / 8;
    const int64_t panel = i / SC_ROWS, r = i % SC_ROWS;
    const int64_t jt = i / J, jj = i % J;
    double g = 0.0, gb = 0.0;                                   // exact squared norm (as given); of the scaled B side
    int tiny = 0, tiny_b = 0;
    const double fmin_normal = 6.103515625e-05;                 // 2^-14
    auto to_half = [&](double v, int& n_tiny) {
        n_tiny += (v != 0.0 && fabs(v) < fmin_normal);
        return fabs(v) < fmin_normal ? __float2half_rn(0.f) : __double2half(v);
    };
    for (int m = lane; m < Mp; m += 32) {
        double x = 0.0, y = 0.0, z = 0.0;
        if (live && m < M) {
            const double* a = src + (int64_t)heavy_idx[m] * 3;
            const double x0 = a[0], y0 = a[1], z0 = a[2];
            g = fma(x0, x0, fma(y0, y0, fma(z0, z0, g)));
            x = fma(fr.q[0], x0, fma(fr.q[1], y0, fr.q[2] * z0));
            y = fma(fr.q[3], x0, fma(fr.q[4], y0, fr.q[5] * z0));
            z = fma(fr.q[6], x0, fma(fr.q[7], y0, fr.q[8] * z0));
        }
        const __half hx = to_half(x, tiny), hy = to_half(y, tiny), hz = to_half(z, tiny);
        const double xb = fr.t[0] * x, yb = fr.t[1] * y, zb = fr.t[2] * z;
        gb = fma(xb, xb, fma(yb, yb, fma(zb, zb, gb)));
        const __half bx = to_half(xb, tiny_b), by = to_half(yb, tiny_b), bz = to_half(zb, tiny_b);
        const int kc = m >> 3, e = m & 7;
        __half* pa = PA + (((panel * 3) * nkc + kc) * SC_ROWS + r) * 8 + e;
        pa[0] = hx;
        pa[(int64_t)nkc * SC_ROWS * 8] = hy;
        pa[(int64_t)2 * nkc * SC_ROWS * 8] = hz;
        __half* pb = PB + ((jt * nkc + kc) * (3 * J) + jj) * 8 + e;
        pb[0] = bx;
        pb[J * 8] = by;
        pb[2 * J * 8] = bz;
        __half* pr = PR + (size_t)i * 3 * Mp + m;
        pr[0] = hx;
        pr[Mp] = hy;
        pr[2 * Mp] = hz;
    }
    g = warp_sum(g);
    gb = warp_sum(gb);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tiny += __shfl_xor_sync(0xffffffffu, tiny, o);
        tiny_b += __shfl_xor_sync(0xffffffffu, tiny_b, o);
    }
    if (lane == 0) {
        const double alpha = 6.103515625e-05 * (1.0 + 9.765625e-4) / TF_EPS;
        const double rot = 1.0 + 1e-12;                          // |Q x|^2 <= (1 + 1e-12) |x|^2 (Q checked to 1e-13)
        const double sg = sqrt(g) * rot + alpha * sqrt((double)tiny);                  // row side: rotated, unscaled
        const double sgb = sqrt(gb) * rot + alpha * sqrt((double)tiny_b);              // column side as the MMAs see it
        const double sgu = sqrt(g) * rot + alpha * fr.inv_tmin * sqrt((double)tiny_b); // ... and un-scaled again
        G[i] = g; sG[i] = sg;
        CT[jt * (2 * J) + jj] = __double2float_rd(0.5 * (1.0 - 1e-10) * g);
        CT[jt * (2 * J) + J + jj] = __double2float_ru(fmax(sg, fmax(sgb, sgu)));
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// MODE 0: "isotropic" form — only f = ||S~||_F^2 is accumulated across the three rows and Samuelson's bound decides;
//         what it cannot exclude is a candidate (half the arithmetic, an eighth of the registers: 16 columns per
//         thread, tiles of J = 64).
// MODE 1: T = S~^T S~ accumulated; Samuelson, then the quartic sign test for column pairs with an undecided lane.
// MODE 2: T accumulated; the quartic sign test for every pair (anisotropic ensembles, where Samuelson never excludes).
template <int MODE, int J>
__global__ void __launch_bounds__(sc_threads(J), 1) rmsd_screen_kernel(const ScParams p) {
    constexpr int NN = 3 * J;                      // UMMA N: (component b, conformer j)
    constexpr int NBUF = sc_nbuf(J);               // accumulator buffers of NN columns
    constexpr int LBUF = J == 32 ? 2 : 1;          // log2(NBUF) (J = 48: three buffers, buffer = row a, no shifts)
    constexpr int LREL = J == 32 ? 12 : J == 48 ? 3 : 6;   // lcm(NBUF, 3): ring of buffer-release barriers, see t_rel
    constexpr int SC_EPI_WARPS = sc_epi_warps(J);
    constexpr int SC_KT = sc_kt(J);
    constexpr int SC_ACC0 = sc_acc0(J);
    constexpr int COLS = J / (SC_EPI_WARPS / 4);   // columns of a tile per epilogue warp
    constexpr int NCP = COLS / 2;                  // column pairs (two FP32 lanes per packed instruction)
    constexpr int NR = 3 * COLS;                   // accumulator words a thread reads per unit
    static_assert(J == 32 || J == 48 || J == 64, "tile width");
    static_assert(MODE == 0 || J == 32, "the T-accumulating modes keep 6 numbers per pair: 8 columns per thread");
    static_assert(COLS == 8 || COLS == 16, "whole bytes of the bit rows per warp");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nkc = p.nkc;
    const int nkb = nkc / 2;                                   // K blocks (one MMA per component each)
    const int KT = nkb < SC_KT ? nkb : SC_KT;                  // K blocks of the panel held in TMEM
    const int tail_kc = nkc - 2 * KT;                          // chunks of the panel kept in shared memory
    const uint32_t tail_bytes = (uint32_t)tail_kc * SC_ROWS * 16u;      // per component
    const uint32_t b_bytes = (uint32_t)nkc * NN * 16u;
    unsigned char* smA = smem_raw;                             // [a][tail_kc][128][16 B]
    unsigned char* smB = smem_raw + 3u * tail_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)p.nb_stages * b_bytes);
    uint64_t* at_full = bars;                  // tail of the panel landed (TMA)
    uint64_t* am_full = bars + 1;              // panel rows stored to TMEM (16 epilogue warps)
    uint64_t* a_empty = bars + 2;              // all MMAs of the item retired
    uint64_t* b_full = bars + 3;
    uint64_t* b_empty = b_full + SC_MAX_BSTAGES;               // three arrivals: one per row unit of the tile
    uint64_t* t_full = b_empty + SC_MAX_BSTAGES;
    // Buffer release.  A waiter on an mbarrier must see every phase of it (parity waits alias two phases apart), but
    // with three issuing threads and NBUF buffers the uses of one buffer rotate through the chains.  So the releases
    // go to a ring of LREL = lcm(NBUF, 3) barriers indexed by the unit that may START: the epilogue, having drained
    // unit u, arrives on t_rel[(u + NBUF) % LREL]; unit v waits on t_rel[v % LREL] — always the same chain for a
    // given slot, once every LREL units.  (With one barrier per buffer the 64-wide form dead-locked; the 32-wide one
    // only worked because its chains happened to arrive late.)
    uint64_t* t_rel = t_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_rel + 12);
    int2* cand_q = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(bars) + 512);     // [epilogue warp][SC_Q]
    unsigned char* ct_stage = reinterpret_cast<unsigned char*>(cand_q + SC_MAX_EPI_WARPS * SC_Q);   // [epilogue warp][SC_CT_BYTES]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(at_full, 1);
        mbar_init(am_full, SC_EPI_WARPS);
        mbar_init(a_empty, SC_MMA_WARPS);
        for (int s = 0; s < p.nb_stages; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], SC_MMA_WARPS); }
        for (int t = 0; t < NBUF; t++) mbar_init(&t_full[t], 1);
        for (int t = 0; t < LREL; t++) mbar_init(&t_rel[t], SC_EPI_WARPS);
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
    } else if (warp <= SC_MMA_WARPS) {
        // ===================== MMA issue: three warps, one elected thread each =====================
        // Chain l = warp - 1 owns row l of the covariances: for every tile of the CTA it issues the nkb MMAs of unit
        // (tile, l) into accumulator buffer (unit number mod NBUF) and commits them to that buffer's barrier.  MMAs
        // that accumulate into the same buffer are a dependent chain (~144 cycles each when issued back to back); with
        // three issuing threads the chains of different units overlap in the tensor pipe whenever buffers are free.
        // (A single thread issuing three skewed chains needs ~70 instructions of bookkeeping per MMA: measured 245
        // cycles per MMA.)  `pace` > 0 additionally spaces the MMAs of a chain by that many cycles (measurement aid:
        // no gain, the epilogue is what bounds this kernel).
        if (elect_one()) {
            const int l = warp - 1;
            const uint32_t idesc = umma_idesc_f16(SC_ROWS, NN);
            const uint32_t a_lbo = SC_ROWS * 16u, b_lbo = NN * 16u;
            const uint64_t bd0 = umma_desc_kmajor(smem_u32(smB), b_lbo, 128u);       // stage 0, K block 0
            const uint64_t ad0 = umma_desc_kmajor(smem_u32(smA), a_lbo, 128u) + (uint64_t)l * (tail_bytes >> 4);
            const uint64_t bd_step = (2u * b_lbo) >> 4, ad_step = (2u * a_lbo) >> 4; // start-address field, 16-byte units
            const uint64_t b_stage_step = b_bytes >> 4;
            const uint32_t at0 = tmem_base + (uint32_t)l * 8u * (uint32_t)KT;        // this row's part of the panel in TMEM
            int st = 0;
            uint32_t sph = 0, aph = 0, rph = 0, us = (uint32_t)l;                    // unit number of the current unit
            long long t_next = clock64();
            const long long pace = p.pace;
            auto paced = [&]() {
                if (pace > 0) {
                    long long now = clock64();
                    while (now < t_next) now = clock64();
                    t_next = now + pace;
                }
            };
            for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
                const int4 w = p.items[it];
                if (w.z == 0) break;
                mbar_wait(am_full, aph);
                if (tail_kc > 0) mbar_wait(at_full, aph);
                aph ^= 1u;
                for (int t = 0; t < w.z; t++) {
                    const uint32_t buf = NBUF == 3 ? (uint32_t)l : us & (NBUF - 1);   // three buffers: chain l owns buffer l
                    SC_STAMP(us < 192, (us * 8) + 0);
                    mbar_wait(&b_full[st], sph);
                    SC_STAMP(us < 192, (us * 8) + 1);
                    if (us >= NBUF) {                           // the unit that used this buffer before has been drained
                        if (NBUF == 3) {
                            mbar_wait(&t_rel[l], rph);
                            rph ^= 1u;
                        } else {
                            const uint32_t slot = us % LREL, use = us / LREL - (slot < NBUF ? 1u : 0u);
                            mbar_wait(&t_rel[slot], use & 1u);
                        }
                    }
                    tcgen05_fence_after();
                    SC_STAMP(us < 192, (us * 8) + 2);
                    const uint32_t d = tmem_base + SC_ACC0 + buf * NN;
                    uint64_t bd = bd0 + (uint64_t)st * b_stage_step;
                    uint64_t ad = ad0;
                    uint32_t at = at0;
                    paced();
                    umma_f16_ts<false>(d, at, bd, idesc);                   // K block 0 overwrites the buffer
                    for (int kb = 1; kb < KT; kb++) {
                        bd += bd_step;
                        at += 8;
                        paced();
                        umma_f16_ts<true>(d, at, bd, idesc);
                    }
                    for (int kb = KT; kb < nkb; kb++) {                              // panel block in shared memory
                        bd += bd_step;
                        paced();
                        umma_f16_ss<true>(d, ad, bd, idesc);
                        ad += ad_step;
                    }
                    umma_commit(&t_full[buf]);                 // this row of the tile is ready for the epilogue
                    umma_commit(&b_empty[st]);                 // (the stage is free after the third chain's commit)
                    SC_STAMP(us < 192, (us * 8) + 3);
                    us += 3u;
                    if (++st == p.nb_stages) { st = 0; sph ^= 1u; }
                }
                umma_commit(a_empty);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: 16 warps =====================
        const int ew = warp - (1 + SC_MMA_WARPS);
        const int quad = warp & 3;                               // TMEM lane quarter this warp may access
        const int part = ew >> 2;                                // which COLS columns of every tile
        const int row_in_panel = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int c0 = part * COLS;
        uint32_t eph = 0, us = 0;                                // unit number (all units of this CTA, in order)
        uint32_t rel_slot = NBUF % LREL;                         // (us + NBUF) % LREL: where this unit's release goes
        uint32_t tph = 0;                                        // three buffers: parity of t_full[a] = tile count & 1
        unsigned char* my_ct = ct_stage + (size_t)ew * SC_CT_BYTES;
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
            const ScRow row = screen_row_consts(p.G[i], p.sG[i], p.e_thr);
            const unsigned long long Af2 = OpsF2::bc(row.Af), nCf2 = OpsF2::bc(-row.Cf);
            uint8_t* out_row = p.sim_bits8 + ((int64_t)w.w * CB + row_in_panel) * (4 * p.W);
            for (int t = 0; t < w.z; t++) {
                const int64_t j0 = (int64_t)(w.y + t) * J + c0;              // first of this warp's COLS columns
                const float4* ct = reinterpret_cast<const float4*>(p.CT + (int64_t)(w.y + t) * (2 * J) + c0);
                float4 ctB[COLS / 4], ctD[COLS / 4];                         // B_j, D_j of this warp's columns
                unsigned long long T[MODE == 0 ? 1 : 6][NCP];                // T = S~^T S~ (MODE 0: f), two columns per register pair
                if (MODE == 0) {
                    // column terms of this tile -> this warp's staging slot, asynchronously: they are needed after the
                    // third unit, and an L2 round trip there (16 warps waiting on it) cost ~650 cycles per tile
                    __syncwarp();
                    if (lane < 8) cp_async16(my_ct + lane * 16, reinterpret_cast<const float*>(ct) + (lane < 4 ? lane * 4 : J + (lane - 4) * 4));
                    cp_async_commit();
                }
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    uint32_t r[NR];                                          // row a of S~ for COLS columns: [b][column]
                    const uint32_t buf = NBUF == 3 ? (uint32_t)a : us & (NBUF - 1);
                    SC_STAMP(ew == 0 && lane == 0 && us < 192, (us * 8) + 4);
                    mbar_wait(&t_full[buf], NBUF == 3 ? tph : (us >> LBUF) & 1u);
                    tcgen05_fence_after();
                    SC_STAMP(ew == 0 && lane == 0 && us < 192, (us * 8) + 5);
                    const uint32_t d0 = tmem_base + lane_addr + SC_ACC0 + buf * NN + (uint32_t)c0;
#pragma unroll
                    for (int b = 0; b < 3; b++)
#pragma unroll
                        for (int h = 0; h < COLS / 8; h++) tmem_ld_x8_raw(d0 + (uint32_t)(b * J + 8 * h), &r[b * COLS + 8 * h]);
#pragma unroll
                    for (int h = 0; h < NR / 24; h++) tmem_wait_bind24(&r[24 * h]);
                    SC_STAMP(ew == 0 && lane == 0 && us < 192, (us * 8) + 6);
                    tcgen05_fence_before();                                  // the buffer goes back to its MMA thread
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_rel[rel_slot]);
                    rel_slot = rel_slot + 1 == LREL ? 0u : rel_slot + 1u;
                    SC_STAMP(ew == SC_EPI_WARPS - 1 && lane == 0 && us < 192, (us * 8) + 7);
                    us++;
                    if (a == 2 && MODE != 0) {                               // column terms: in flight while T is finished
#pragma unroll
                        for (int h = 0; h < COLS / 4; h++) { ctB[h] = __ldg(ct + h); ctD[h] = __ldg(ct + J / 4 + h); }
                    }
#pragma unroll
                    for (int cp = 0; cp < NCP; cp++) {
                        const unsigned long long v0 = OpsF2::pack(r[2 * cp], r[2 * cp + 1]);
                        const unsigned long long v1 = OpsF2::pack(r[COLS + 2 * cp], r[COLS + 2 * cp + 1]);
                        const unsigned long long v2 = OpsF2::pack(r[2 * COLS + 2 * cp], r[2 * COLS + 2 * cp + 1]);
                        if (MODE == 0) {
                            T[0][cp] = (a == 0) ? OpsF2::mul(v0, v0) : OpsF2::fma(v0, v0, T[0][cp]);
                            T[0][cp] = OpsF2::fma(v1, v1, T[0][cp]);
                            T[0][cp] = OpsF2::fma(v2, v2, T[0][cp]);
                        } else if (a == 0) {
                            T[0][cp] = OpsF2::mul(v0, v0); T[1][cp] = OpsF2::mul(v1, v1); T[2][cp] = OpsF2::mul(v2, v2);
                            T[3][cp] = OpsF2::mul(v0, v1); T[4][cp] = OpsF2::mul(v0, v2); T[5][cp] = OpsF2::mul(v1, v2);
                        } else {
                            T[0][cp] = OpsF2::fma(v0, v0, T[0][cp]); T[1][cp] = OpsF2::fma(v1, v1, T[1][cp]);
                            T[2][cp] = OpsF2::fma(v2, v2, T[2][cp]); T[3][cp] = OpsF2::fma(v0, v1, T[3][cp]);
                            T[4][cp] = OpsF2::fma(v0, v2, T[4][cp]); T[5][cp] = OpsF2::fma(v1, v2, T[5][cp]);
                        }
                    }
                }
                tph ^= 1u;
                if (MODE == 0) {
                    cp_async_wait<0>();
                    __syncwarp();
                    const float4* cs = reinterpret_cast<const float4*>(my_ct);
#pragma unroll
                    for (int h = 0; h < COLS / 4; h++) { ctB[h] = cs[h]; ctD[h] = cs[4 + h]; }
                }
                // ---- the tile's verdicts for this warp's columns: column terms B_j (rounded down), D_j (rounded up)
                uint32_t near = 0;                                           // bit c: pair of column c0 + c not excluded
                if (MODE == 0) {
                    // Samuelson only (the bound and its constant: see the general form below), leaner bit gathering:
                    // with lfp = max(lf, 0) the single value t = 3.00004 f - lfp^2 is negative exactly for the excluded
                    // pairs (lf <= 0 or NaN -> lfp = 0 -> t >= 0; f NaN -> t = NaN, sign clear), and its sign bits are
                    // shifted into the word one funnel shift per pair, last column first
                    uint32_t excl = 0;
#pragma unroll
                    for (int cp = NCP - 1; cp >= 0; cp--) {
                        const float4 cb = ctB[cp / 2], cd = ctD[cp / 2];
                        const unsigned long long Bj2 = (cp & 1) ? OpsF2::pack(__float_as_uint(cb.z), __float_as_uint(cb.w))
                                                               : OpsF2::pack(__float_as_uint(cb.x), __float_as_uint(cb.y));
                        const unsigned long long Dj2 = (cp & 1) ? OpsF2::pack(__float_as_uint(cd.z), __float_as_uint(cd.w))
                                                               : OpsF2::pack(__float_as_uint(cd.x), __float_as_uint(cd.y));
                        const unsigned long long lf2 = OpsF2::fma(nCf2, Dj2, OpsF2::add(Af2, Bj2));
                        float l0, l1, t0, t1;
                        OpsF2::unpack(lf2, l0, l1);
                        const unsigned long long lp2 = OpsF2::pack(__float_as_uint(fmaxf(l0, 0.0f)), __float_as_uint(fmaxf(l1, 0.0f)));
                        const unsigned long long t2 =
                            OpsF2::fma(OpsF2::bc(3.00004f), T[0][cp], OpsF2::mul(OpsF2::mul(lp2, lp2), OpsF2::bc(-1.0f)));
                        OpsF2::unpack(t2, t0, t1);
                        excl = __funnelshift_l(__float_as_uint(t1), excl, 1);
                        excl = __funnelshift_l(__float_as_uint(t0), excl, 1);
                    }
                    near = ~excl;
                }
#pragma unroll
                for (int cp = 0; cp < (MODE == 0 ? 0 : NCP); cp++) {
                    unsigned long long tt[6];
#pragma unroll
                    for (int q = 0; q < 6; q++) tt[q] = T[MODE == 0 ? 0 : q][cp];
                    const float4 cb = ctB[cp / 2], cd = ctD[cp / 2];
                    const unsigned long long Bj2 = (cp & 1) ? OpsF2::pack(__float_as_uint(cb.z), __float_as_uint(cb.w))
                                                           : OpsF2::pack(__float_as_uint(cb.x), __float_as_uint(cb.y));
                    const unsigned long long Dj2 = (cp & 1) ? OpsF2::pack(__float_as_uint(cd.z), __float_as_uint(cd.w))
                                                           : OpsF2::pack(__float_as_uint(cd.x), __float_as_uint(cd.y));
                    // stage 1, Samuelson: lambda_max <= sqrt(3) ||S~||_F.  With lf = A_i + B_j - C_i D_j (a lower bound
                    // of the lowered threshold eigenvalue up to a factor 1 - 1e-6, folded into the constant together with
                    // the rounding of f and of the products: 3 (1 + 9 u) / (1 - 1e-6)^2 < 3.00004) the pair is excluded
                    // iff lf > 0 and 3.00004 f - lf^2 < 0, read off the sign bits (NaN / inf: not excluded).
                    const unsigned long long f2 = MODE == 0 ? tt[0] : OpsF2::add(OpsF2::add(tt[0], tt[1]), tt[2]);
                    const unsigned long long ab2 = OpsF2::add(Af2, Bj2);
                    const unsigned long long lf2 = OpsF2::fma(nCf2, Dj2, ab2);
                    const unsigned long long t2 = OpsF2::fma(OpsF2::bc(3.00004f), f2, OpsF2::mul(OpsF2::mul(lf2, lf2), OpsF2::bc(-1.0f)));
                    float lf[2], ab[2], tv[2];
                    OpsF2::unpack(lf2, lf[0], lf[1]);
                    OpsF2::unpack(ab2, ab[0], ab[1]);
                    OpsF2::unpack(t2, tv[0], tv[1]);
                    uint32_t und = 3u;
                    if (MODE != 2) {
                        und = 0;
#pragma unroll
                        for (int h = 0; h < 2; h++)
                            und |= (((__float_as_uint(tv[h]) & ~__float_as_uint(lf[h])) >> 31) ^ 1u) << h;
                    }
                    // stage 2, FP32 sign test of the key-matrix quartic from T, for a column pair with an undecided lane
                    if (MODE == 2 || (MODE == 1 && __any_sync(0xffffffffu, und))) {
                        float lam[2];
#pragma unroll
                        for (int h = 0; h < 2; h++) lam[h] = fmaf(-2e-7f, fabsf(ab[h]) + fabsf(lf[h]), lf[h]);
                        if (p.scaled) {                                      // T^ = diag(t) T diag(t)  ->  T
#pragma unroll
                            for (int q = 0; q < 6; q++) tt[q] = OpsF2::mul(tt[q], OpsF2::bc(p.inv_tt[q]));
                        }
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
                constexpr uint32_t ALL = (1u << COLS) - 1u;
                uint32_t valid = ALL;
                if (j0 + COLS - 1 >= p.N) valid = (j0 >= p.N) ? 0u : (ALL >> (j0 + COLS - p.N));
                if (j0 <= i) valid &= (i - j0 >= COLS - 1) ? 0u : (ALL << (i - j0 + 1));
                const uint32_t bits = (i < p.N) ? (near & valid & ALL) : 0u;
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
                if (i < p.N && (j0 >> 5) < p.W) {
                    if (COLS == 8) out_row[j0 >> 3] = (uint8_t)bits;
                    else *reinterpret_cast<uint16_t*>(out_row + (j0 >> 3)) = (uint16_t)bits;
                }
            }
        }
        if (p.cand && qn) flush_q();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, SC_TMEM);
}

}  // namespace tsc

// rows the images are padded to: whole panels of 128 AND whole tiles of 32 / 48 / 64 conformers (lcm = 384); the padding
// rows are written as zeros by tsc_pack_screen
extern "C" int64_t tsc_screen_rows_padded(int64_t N) { return (N + 383) / 384 * 384; }
// size in bytes of each of the three operand images (PA, PB, PR)
extern "C" int64_t tsc_screen_operand_bytes(int64_t N, int32_t M) {
    const int64_t Mp = (M + 15) / 16 * 16;
    return tsc_screen_rows_padded(N) * 3 * Mp * 2;
}
extern "C" int64_t tsc_screen_ct_floats(int64_t N) { return tsc_screen_rows_padded(N) * 2; }
static size_t screen_side_bytes() {      // barriers, candidate queues and column-term slots of the epilogue warps
    return 512 + (size_t)tsc::SC_MAX_EPI_WARPS * (tsc::SC_Q * sizeof(int2) + tsc::SC_CT_BYTES);
}
static bool screen_fits(int M, int J, size_t* a_out, size_t* b_out) {
    const int nkc = (M + 15) / 16 * 2, nkb = nkc / 2, KT = nkb < tsc::sc_kt(J) ? nkb : tsc::sc_kt(J);
    const size_t a_bytes = (size_t)3 * (nkc - 2 * KT) * tsc::SC_ROWS * 16, b_bytes = (size_t)nkc * 3 * J * 16;
    const size_t budget = 227 * 1024 - screen_side_bytes();
    if (a_out) *a_out = a_bytes;
    if (b_out) *b_out = b_bytes;
    return a_bytes + 3 * b_bytes <= budget;
}
// largest number of heavy atoms the screen takes with tiles of `tile_j` (32, 48 or 64) conformers (panel tail + three
// B stages must fit in shared memory); above it the FP64 tensor-core variant runs (tsc_rmsd_sim_tiles)
extern "C" int32_t tsc_screen_max_atoms(int32_t tile_j) {
    int best = 0;
    if (tile_j != 32 && tile_j != 48 && tile_j != 64) return 0;
    for (int M = 16; M <= 1024; M += 16)
        if (screen_fits(M, tile_j, nullptr, nullptr)) best = M;
    return best;
}

// rows [row_begin, row_end) only (row_begin a multiple of 8; the last chunk should end at tsc_screen_rows_padded(N), or
// pass row_end <= 0 for "to the end", so that the padding rows are written too).  tile_j: 32, 48 or 64, the tile
// width of the screen mode that will read PB / CT (tsc_rmsd_screen: mode 0 -> 48, modes 1 and 2 -> 32, mode 3 -> 64).
// frame (host memory, 12 doubles, or NULL = identity): orthogonal Q (row-major, rows = axes) and the B-side scales t
// (ScFrame above; _host.screen_frame).  Returns cudaErrorInvalidValue unless |Q Q^T - I| <= 1e-13 and
// sum_b 1 / (3 t_b^2) <= 1 — the two facts the exclusion tests rest on.
static bool screen_frame_from(const double* frame, tsc::ScFrame& fr) {
    for (int k = 0; k < 9; k++) fr.q[k] = frame ? frame[k] : (k % 4 == 0 ? 1.0 : 0.0);
    for (int k = 0; k < 3; k++) fr.t[k] = frame ? frame[9 + k] : 1.0;
    double inv_w = 0.0, tmin = fr.t[0];
    for (int b = 0; b < 3; b++) {
        if (!(fr.t[b] > 1e-3 && fr.t[b] < 1e3)) return false;
        inv_w += 1.0 / (3.0 * fr.t[b] * fr.t[b]);
        tmin = fr.t[b] < tmin ? fr.t[b] : tmin;
        for (int c = 0; c < 3; c++) {
            double d = (b == c) ? -1.0 : 0.0;
            for (int k = 0; k < 3; k++) d += fr.q[3 * b + k] * fr.q[3 * c + k];
            if (!(d <= 1e-13 && d >= -1e-13)) return false;
        }
    }
    fr.inv_tmin = 1.0 / tmin;
    return inv_w <= 1.0;
}

extern "C" int tsc_pack_screen(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M, void* PA,
                               void* PB, void* PR, double* G, double* sG, float* CT, int64_t row_begin, int64_t row_end,
                               int32_t tile_j, const double* frame, void* stream) {
    using namespace tsc;
    if (N <= 0 || M <= 0) return 0;
    ScFrame fr;
    if (!screen_frame_from(frame, fr)) return (int)cudaErrorInvalidValue;
    if (tile_j != 32 && tile_j != 48 && tile_j != 64) return (int)cudaErrorInvalidValue;
    const int Mp = (M + 15) / 16 * 16;
    const int64_t rows_pad = tsc_screen_rows_padded(N);
    if (row_end <= 0 || row_end > rows_pad) row_end = rows_pad;
    if (row_begin < 0) row_begin = 0;
    if (row_end <= row_begin) return 0;
    const unsigned grid = (unsigned)((row_end - row_begin + 7) / 8);
    if (tile_j == 64)
        pack_screen_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(
            S, N, A, heavy_idx, M, Mp, row_end, reinterpret_cast<__half*>(PA), reinterpret_cast<__half*>(PB),
            reinterpret_cast<__half*>(PR), G, sG, CT, row_begin, fr);
    else if (tile_j == 48)
        pack_screen_kernel<48><<<grid, 256, 0, (cudaStream_t)stream>>>(
            S, N, A, heavy_idx, M, Mp, row_end, reinterpret_cast<__half*>(PA), reinterpret_cast<__half*>(PB),
            reinterpret_cast<__half*>(PR), G, sG, CT, row_begin, fr);
    else
        pack_screen_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(
            S, N, A, heavy_idx, M, Mp, row_end, reinterpret_cast<__half*>(PA), reinterpret_cast<__half*>(PB),
            reinterpret_cast<__half*>(PR), G, sG, CT, row_begin, fr);
    TSC_CHECK_LAUNCH();
    return 0;
}

#ifdef TSC_SCREEN_TRACE
static long long* g_screen_trace = nullptr;
extern "C" void tsc_screen_set_trace(void* dev_ptr) { g_screen_trace = reinterpret_cast<long long*>(dev_ptr); }
#endif

// items (n_items, 4) int32: {panel, first j tile, number of j tiles, local row block (32-row units) of the panel's
// first row inside sim_bits}; j tiles are 48 conformers wide in mode 0, 32 in modes 1 / 2 and 64 in mode 3 (the images
// must have been packed with the matching tile_j; the first tile of a panel is the one holding its first column);
// dealt round-robin to the CTAs of the persistent grid, an item with count 0 ends a CTA's list
// (_host.build_screen_items).  cand_list: header + (local row, j) entries, or NULL.
// mode: 0 = isotropic form (Samuelson only), 1 = Samuelson then quartic, 2 = quartic for every pair, 3 = mode 0 on
// 64-wide tiles with two accumulator buffers (see the kernel).
// pace: 0 = none; > 0 = cycles between two MMAs of one chain (measurement aid).  frame: the one given to tsc_pack_screen.
extern "C" int tsc_rmsd_screen(const void* PA, const void* PB, const void* PR, const double* G, const double* sG,
                               const float* CT, int64_t N, int32_t M, const int32_t* items, int32_t n_items, double thr,
                               uint32_t* sim_bits, int32_t* cand_list, int64_t cand_stride, int32_t grid_ctas, int32_t mode,
                               int32_t pace, const double* frame, void* stream) {
    using namespace tsc;
    if (n_items <= 0 || N <= 0) return 0;
    if (mode < 0 || mode > 3) return (int)cudaErrorInvalidValue;
    ScParams p;
    ScFrame fr;
    if (!screen_frame_from(frame, fr)) return (int)cudaErrorInvalidValue;
    p.scaled = !(fr.t[0] == 1.0 && fr.t[1] == 1.0 && fr.t[2] == 1.0);
    {
        const int bb[6] = {0, 1, 2, 0, 0, 1}, cc[6] = {0, 1, 2, 1, 2, 2};
        for (int q = 0; q < 6; q++) p.inv_tt[q] = (float)(1.0 / (fr.t[bb[q]] * fr.t[cc[q]]));
    }
    p.pace = pace <= 0 ? 0 : pace;
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
#ifdef TSC_SCREEN_TRACE
    p.trace = g_screen_trace;
#endif
    const int J = mode == 0 ? 48 : mode == 3 ? 64 : 32;
    size_t a_bytes, b_bytes;
    if (!screen_fits(M, J, &a_bytes, &b_bytes)) return (int)cudaErrorInvalidValue;   // more atoms than tsc_screen_max_atoms(J)
    const size_t budget = 227 * 1024 - screen_side_bytes();
    int nb = (int)((budget - a_bytes) / b_bytes);
    if (nb > SC_MAX_BSTAGES) nb = SC_MAX_BSTAGES;
    p.nb_stages = nb;
    const size_t smem = a_bytes + nb * b_bytes + screen_side_bytes();
    auto kern = mode == 0 ? rmsd_screen_kernel<0, 48> : mode == 3 ? rmsd_screen_kernel<0, 64>
              : mode == 2 ? rmsd_screen_kernel<2, 32> : rmsd_screen_kernel<1, 32>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = grid_ctas > 0 ? grid_ctas : sms;
    if (grid > n_items) grid = n_items;
    kern<<<grid, sc_threads(J), smem, (cudaStream_t)stream>>>(p);
    TSC_CHECK_LAUNCH();
    return 0;
}
