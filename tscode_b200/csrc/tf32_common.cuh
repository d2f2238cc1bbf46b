// tf32_common.cuh — pieces shared by the two tcgen05 pre-screen kernels (rmsd_tf32.cu: both MMA
// operands from shared memory; rmsd_tf32ts.cu: the stationary operand held in TMEM).
#pragma once
#include "tsc_common.cuh"
#include "tsc_math.cuh"

namespace tsc {

constexpr int TF_ROWS = 128;                   // conformers per A panel  (UMMA M)
constexpr int TF_J = 16;                       // conformers per B tile
constexpr int TF_N = 3 * TF_J;                 // UMMA N = 48
constexpr int TF_ACC_COLS = 3 * TF_N;          // 144 TMEM columns per accumulator buffer
constexpr int TF_NACC = 3;                     // accumulator buffers of rmsd_tf32.cu (432 of 512 columns)
constexpr int TF_TMEM_COLS = 512;
constexpr int TF_MAX_BSTAGES = 6;
constexpr int TF_MAX_GROUPS = 3;               // epilogue groups of 4 warps
constexpr double TF_EPS = 1.05e-3;             // see header
constexpr int TF_DEFAULT_CFG = 3;              // see tsc_rmsd_sim_tf32

struct TfParams {
    const float* PA;          // [panel][a][kc][128][4]
    const float* PB;          // [jtile][kc][48][4]
    const double* G;          // (>= njt*16) squared norms, exact FP64
    const double* sG;         // sqrt(G)
    const int4* items;        // (panel, jt_begin, jt_count, local_row_block_of_panel)
    int n_items;
    int64_t N;
    int Mp;                   // atoms padded to a multiple of 8
    int nb_stages;
    double e_thr;             // M thr^2 (1 + 1e-6)
    uint16_t* sim_bits16;
    int64_t W;                // words per sim row
};

__device__ __forceinline__ void tmem_ld_x8_raw(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x4_raw(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_bind12(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11])::"memory");
}
// tcgen05.wait::ld carrying 24 registers as in/out operands so that no use can be hoisted above it
__device__ __forceinline__ void tmem_wait_bind24(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                   "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                   "+r"(r[22]), "+r"(r[23])::"memory");
}

__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T : the stationary operand read from tensor memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Lean issue forms for the MMA loop.  A single thread issues every MMA of the CTA, and a 128x48x8
// tf32 MMA only occupies the tensor pipe for ~24 cycles, so the instruction stream between two
// MMAs must stay below that: the accumulate flag is a compile-time constant (ptxas folds the
// predicate), descriptors are advanced with one integer add (tools/umma_probe.py: a loop that
// rebuilds descriptors and predicates per MMA issues one MMA every ~130 cycles whatever N is).
template <bool ACC, bool F16 = false>
__device__ __forceinline__ void umma_tf32_ts_c(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    if (F16)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0)
            : "memory");
}
template <bool ACC, bool F16 = false>
__device__ __forceinline__ void umma_tf32_ss_c(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    if (F16)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0)
            : "memory");
}

// One 128 x 16 tile of the epilogue for the thread owning row i: read the 16 x 9 accumulators of
// this row from TMEM (base address d0, lane already selected), release the buffer (t_empty) once
// they are in registers, screen the 16 pairs and return the 16 result bits (validity-masked).
//
// Per pair the threshold eigenvalue, lowered by the TF32 error bound, is
//     lam = 0.5 (1-1e-10) (G_i + G_j) - 0.5 e_thr - sqrt(3) eps sqrt(G_i) sqrt(G_j).
//   fast path (FP32 only): a guaranteed LOWER bound lf of lam from directed-rounded per-conformer
//     terms (A_i + B_j - C_i D_j, then scaled by 1 - 1e-6 to cover the two FP32 roundings), and
//     Samuelson's bound lambda_max <= sqrt(3) ||S~||_F: the pair is excluded if 3.00003 f <= lf^2
//     with f = sum S~^2 (relative rounding error of f <= 9 * 2^-24) — one warp vote per STEP columns;
//   full path (FP64, branch-free, taken when some lane of the warp could not exclude a pair):
//     Budan-Fourier sign test on the key-matrix quartic at lam.
//   gvf: lanes 0..15 hold B_j = float_rd(hs G[j0+lane]), lanes 16..31 D_j = float_ru(sqrt(G)[j0+lane-16]).
struct TfRow {                 // per-thread (row i) constants
    float Af, Cf;              // A_i = float_rd(hs G_i - 0.5 e_thr),  C_i = float_ru(sqrt(3) eps sqrt(G_i))
    double hi, ci;             // the same in FP64: hi = hs G_i - 0.5 e_thr,  ci = -sqrt(3) eps sqrt(G_i)
};
__device__ __forceinline__ TfRow tf32_row_consts(double Gi, double sGi, double e_thr) {
    const double hs = 0.5 * (1.0 - 1e-10), cc = 1.7320508075688772 * TF_EPS;
    TfRow r;
    r.hi = fma(hs, Gi, -0.5 * e_thr);
    r.ci = -cc * sGi;
    r.Af = __double2float_rd(r.hi);
    r.Cf = __double2float_ru(cc * sGi);
    return r;
}
__device__ __forceinline__ float tf32_col_term(const double* __restrict__ G, const double* __restrict__ sG,
                                               int64_t j0, int lane) {
    const double hs = 0.5 * (1.0 - 1e-10);
    return (lane < 16) ? __double2float_rd(hs * G[j0 + lane]) : __double2float_ru(sG[j0 + lane - 16]);
}

// NCOL columns starting at column C0 of the tile are handled by this thread (NCOL = 16: whole tile).
// The TMEM reads are software-pipelined: the loads of column group s+1 are in flight while group s
// is screened (ncu on the first version: the epilogue warps issued 570 instructions per tile but
// needed ~3800 cycles for them, almost all of it exposed tcgen05.ld latency — four load/wait/compute
// rounds per tile — while the tensor pipe idled 56 % of the time waiting for a free accumulator).
template <int STEP>
__device__ __forceinline__ void tf32_ld_cols(uint32_t d0, int st, uint32_t* r) {
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const uint32_t ta = d0 + (uint32_t)(a * TF_N + b * TF_J + st * STEP);
            if (STEP == 8) tmem_ld_x8_raw(ta, &r[(3 * a + b) * STEP]);
            else tmem_ld_x4_raw(ta, &r[(3 * a + b) * STEP]);
        }
}
template <int STEP>
__device__ __forceinline__ void tf32_ld_wait(uint32_t* r) {
    if (STEP == 8) { tmem_wait_bind24(&r[0]); tmem_wait_bind24(&r[24]); tmem_wait_bind24(&r[48]); }
    else { tmem_wait_bind12(&r[0]); tmem_wait_bind12(&r[12]); tmem_wait_bind12(&r[24]); }
}

template <int STEP, int NCOL = TF_J>
__device__ __forceinline__ uint32_t tf32_epilogue_tile(uint32_t d0, float gvf, const TfRow& row,
                                                       const double* __restrict__ G, const double* __restrict__ sG,
                                                       int64_t i, int64_t j0, int64_t N, int lane,
                                                       uint64_t* t_empty_bar, int C0 = 0) {
    constexpr int NST = NCOL / STEP;
    uint32_t bits = 0;
    uint32_t rr[2][9 * STEP];
    tf32_ld_cols<STEP>(d0, C0 / STEP, rr[0]);
    tf32_ld_wait<STEP>(rr[0]);
#pragma unroll
    for (int st0 = 0; st0 < NST; st0++) {
        const int st = st0 + C0 / STEP;
        uint32_t* r = rr[st0 & 1];
        if (st0 + 1 < NST) {
            tf32_ld_cols<STEP>(d0, st + 1, rr[(st0 + 1) & 1]);       // in flight while this group is screened
        } else {                          // every value this thread needs from the buffer is now in registers
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty_bar);
        }
        uint32_t near = 0;                // pairs the FP32 bound cannot exclude
#pragma unroll
        for (int c = 0; c < STEP; c++) {
            const float Bj = __shfl_sync(0xffffffffu, gvf, st * STEP + c);
            const float Dj = __shfl_sync(0xffffffffu, gvf, 16 + st * STEP + c);
            const float lf = fmaf(-row.Cf, Dj, row.Af + Bj) * 0.999999f;
            float f = 0.f;
#pragma unroll
            for (int q = 0; q < 9; q++) { const float v = __uint_as_float(r[q * STEP + c]); f = fmaf(v, v, f); }
            const bool far = (lf > 0.f) && (3.00003f * f <= lf * lf);
            near |= (far ? 0u : 1u) << c;
        }
        if (__any_sync(0xffffffffu, near != 0u)) {
            const double hs = 0.5 * (1.0 - 1e-10);
#pragma unroll
            for (int c = 0; c < STEP; c++) {
                const int64_t j = j0 + st * STEP + c;
                double S[9];
#pragma unroll
                for (int q = 0; q < 9; q++) S[q] = (double)__uint_as_float(r[q * STEP + c]);
                double c2, c1, c0;
                key_charpoly(S, c2, c1, c0);
                const double l1 = fma(row.ci, sG[j], fma(hs, G[j], row.hi));
                const bool excluded = quartic_excluded(l1, c2, c1, c0);
                if (!excluded && ((near >> c) & 1u)) bits |= 1u << (st * STEP + c);
            }
        }
        if (st0 + 1 < NST) tf32_ld_wait<STEP>(rr[(st0 + 1) & 1]);
    }
    // validity mask: j > i, j < N   (rows i >= N are not stored by the caller)
    uint32_t valid = 0xffffu;
    if (j0 + 15 >= N) valid = (j0 >= N) ? 0u : (0xffffu >> (j0 + 16 - N));
    if (j0 <= i) valid &= (i - j0 >= 15) ? 0u : (0xffffu << (i - j0 + 1));
    return bits & valid;
}


// Release-early form of the tile epilogue (16 columns, 4 per TMEM load round) used by the default
// TMEM-operand kernel.  A clock64 trace of the previous form (tools/trace_probe.py) showed the MMA warp
// needs ~1100 cycles to issue a tile and most tiles hand over in ~1400, but every tile in which some
// warp took the FP64 path (one in three on C3: a single undecided pair drags 4 columns x 32 rows
// through ~80 FP64 instructions each, behind two dependent global loads) held its accumulator buffer
// for 3000-5000 cycles.  Here
//   pass 1  reads the buffer once (loads of round s+1 in flight while round s is reduced) and keeps only
//           f = ||S~||_F^2 per pair (16 registers), then takes the FP32 Samuelson decision for all 16;
//   stash   the COLUMNS in which some lane is undecided (one REDUX over the warp) are re-read from TMEM,
//           one 32x32b.x1 load per covariance entry: the first two into registers, any further ones are
//           screened on the spot (rare);
//   release the buffer goes back to the MMA warp;
//   pass 2  the FP64 quartic sign test runs on the stashed columns, after the release.
// gvf holds the column terms of the tile (lanes 0..15 B_j, 16..31 D_j), prefetched by the caller.
__device__ __forceinline__ void tmem_ld_col9(uint32_t d0, int c, uint32_t* r) {
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];"
                         : "=r"(r[3 * a + b])
                         : "r"(d0 + (uint32_t)(a * TF_N + b * TF_J + c)));
}
__device__ __forceinline__ void tmem_wait_bind9(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8])::"memory");
}
// FP64 Budan-Fourier test of one pair on the TF32 covariance: true = provably not similar
__device__ __forceinline__ bool tf32_quartic_excluded(const uint32_t* r, const TfRow& row, double Gj, double sGj) {
    const double hs = 0.5 * (1.0 - 1e-10);
    double S[9];
#pragma unroll
    for (int q = 0; q < 9; q++) S[q] = (double)__uint_as_float(r[q]);
    double c2, c1, c0;
    key_charpoly(S, c2, c1, c0);
    const double l1 = fma(row.ci, sGj, fma(hs, Gj, row.hi));
    return quartic_excluded(l1, c2, c1, c0);
}

// NCOL (16 or 8) columns starting at column C0 of the tile are handled by the calling thread.
template <int NCOL>
__device__ __forceinline__ uint32_t tf32_epilogue_tile_v4(uint32_t d0, float gvf, const TfRow& row,
                                                          const double* __restrict__ G, const double* __restrict__ sG,
                                                          int64_t i, int64_t j0, int64_t N, int lane,
                                                          uint64_t* t_empty_bar, int C0) {
    constexpr int NST = NCOL / 4;
    float f[NCOL];
    {
        uint32_t rr[2][36];
        tf32_ld_cols<4>(d0, C0 / 4, rr[0]);
        tf32_ld_wait<4>(rr[0]);
#pragma unroll
        for (int st = 0; st < NST; st++) {
            const uint32_t* r = rr[st & 1];
            if (st + 1 < NST) tf32_ld_cols<4>(d0, C0 / 4 + st + 1, rr[(st + 1) & 1]);
            // ||S~||_F^2 of two columns per instruction (fma.rn.f32x2 -> FFMA2; per-lane IEEE fma, same values)
#pragma unroll
            for (int cp = 0; cp < 2; cp++) {
                unsigned long long acc = 0ull;
#pragma unroll
                for (int q = 0; q < 9; q++) {
                    unsigned long long v;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(r[q * 4 + 2 * cp]), "r"(r[q * 4 + 2 * cp + 1]));
                    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(v));
                }
                uint32_t lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(acc));
                f[st * 4 + 2 * cp] = __uint_as_float(lo);
                f[st * 4 + 2 * cp + 1] = __uint_as_float(hi);
            }
            if (st + 1 < NST) tf32_ld_wait<4>(rr[(st + 1) & 1]);
        }
    }
    // Samuelson decision per pair, branch- and select-free: with lf = A_i + B_j - C_i D_j (lower bound of the
    // threshold eigenvalue up to a factor 1 - 1e-6, folded into the constant: 3 (1 + 9 2^-24) / (1 - 1e-6)^2 <
    // 3.00004 also covers the roundings of the two products) the pair is excluded iff lf > 0 and
    // t = 3.00004 f - lf^2 < 0, i.e. iff the sign bit of t is set and that of lf is clear (an FMA keeps the sign
    // of the exact result; NaN / inf from an overflowed FP16 operand give t = +inf or the canonical NaN
    // 0x7fffffff: not excluded).  The sign bits are funnel-shifted into a mask, one SHF per pair.
    uint32_t farm = 0;                    // bit (NCOL-1-c) = pair of column C0 + c excluded
#pragma unroll
    for (int c = 0; c < NCOL; c++) {
        const float Bj = __shfl_sync(0xffffffffu, gvf, C0 + c);
        const float Dj = __shfl_sync(0xffffffffu, gvf, 16 + C0 + c);
        const float lf = fmaf(-row.Cf, Dj, row.Af + Bj);
        const float t = fmaf(3.00004f, f[c], -(lf * lf));
        farm = __funnelshift_l(__float_as_uint(t) & ~__float_as_uint(lf), farm, 1);
    }
    // pairs the FP32 bound cannot exclude (bit = column of the tile)
    const uint32_t near = ((~__brev(farm)) >> (32 - NCOL)) << C0;
    uint32_t bits = 0;
    uint32_t cols = __reduce_or_sync(0xffffffffu, near);      // columns with an undecided pair in some lane
    uint32_t sa[9], sb[9];
    int ca = -1, cb = -1;
    if (cols) {
        ca = __ffs(cols) - 1; cols &= cols - 1;
        if (cols) { cb = __ffs(cols) - 1; cols &= cols - 1; }
        while (cols) {                    // third and further undecided columns of a tile: screened before the release
            const int c = __ffs(cols) - 1;
            cols &= cols - 1;
            uint32_t r[9];
            tmem_ld_col9(d0, c, r);
            tmem_wait_bind9(r);
            const int64_t j = j0 + c;
            if (((near >> c) & 1u) && !tf32_quartic_excluded(r, row, G[j], sG[j])) bits |= 1u << c;
        }
        tmem_ld_col9(d0, ca, sa);
        if (cb >= 0) tmem_ld_col9(d0, cb, sb);
        tmem_wait_bind9(sa);
        if (cb >= 0) tmem_wait_bind9(sb);
    }
    tcgen05_fence_before();               // every value needed from the buffer has been read
    __syncwarp();
    if (lane == 0) mbar_arrive(t_empty_bar);
    if (ca >= 0) {
        const int64_t j = j0 + ca;
        if (((near >> ca) & 1u) && !tf32_quartic_excluded(sa, row, G[j], sG[j])) bits |= 1u << ca;
        if (cb >= 0) {
            const int64_t jb = j0 + cb;
            if (((near >> cb) & 1u) && !tf32_quartic_excluded(sb, row, G[jb], sG[jb])) bits |= 1u << cb;
        }
    }
    uint32_t valid = 0xffffu;
    if (j0 + 15 >= N) valid = (j0 >= N) ? 0u : (0xffffu >> (j0 + 16 - N));
    if (j0 <= i) valid &= (i - j0 >= 15) ? 0u : (0xffffu << (i - j0 + 1));
    return bits & valid;
}


// ------------------------------------------------------------------------------------------------------------------
// Two-stage FP32 epilogue (default of the TMEM-operand kernel since round 1 / r04): no FP64, no second TMEM pass.
//   stage 1  Samuelson's bound, as in tf32_epilogue_tile_v4, taken per group of 4 columns while the group's 36
//            accumulators are in registers;
//   stage 2  only for a column pair in which some lane is still undecided (one warp vote): the FP32 sign test of
//            the key-matrix quartic (tsc_math.cuh: quartic32_values / quartic32_decide, rigorous forward error
//            bounds) on the same registers, two columns per instruction (fma/mul/add.f32x2);
//   the rest is a candidate: the verify kernel decides it exactly.
// Why: (a) the FP64 test of v4 ran behind a second TMEM read and two global loads, whole-warp, for one or two
// undecided lanes — a clock64 trace showed tiles with an undecided pair holding their epilogue group for 3000-5000
// cycles (normal: ~900) and the MMA warp stalling on the third buffer behind them; on C3 practically every pair
// Samuelson cannot exclude is a true candidate anyway, so the FP64 test only confirmed what verify re-derives;
// (b) Samuelson's bound needs a near-isotropic covariance: for elongated or planar molecules it excludes nothing and
// v4 ran the FP64 test for every pair (3x slower screen, tools/aniso_probe.py).  The FP32 quartic excludes those
// pairs at ~36 instructions per pair.
// The accumulator buffer is released as soon as the last group is in registers.
// ------------------------------------------------------------------------------------------------------------------
struct OpsF2 {                 // two FP32 lanes per 64-bit register (per-lane IEEE, same values as OpsF32)
    typedef unsigned long long T;
    static __device__ __forceinline__ T fma(T a, T b, T c) { T d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
    static __device__ __forceinline__ T mul(T a, T b) { T d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
    static __device__ __forceinline__ T add(T a, T b) { T d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
    static __device__ __forceinline__ T sub(T a, T b) { T d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
    static __device__ __forceinline__ T pack(uint32_t lo, uint32_t hi) { T d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }
    static __device__ __forceinline__ T bc(float x) { return pack(__float_as_uint(x), __float_as_uint(x)); }
    static __device__ __forceinline__ void unpack(T v, float& lo, float& hi) {
        uint32_t a, b;
        asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
        lo = __uint_as_float(a); hi = __uint_as_float(b);
    }
};

template <int NCOL>
__device__ __forceinline__ uint32_t tf32_epilogue_tile_v5(uint32_t d0, float gvf, const TfRow& row, int64_t i, int64_t j0,
                                                          int64_t N, int lane, uint64_t* t_empty_bar, int C0) {
    constexpr int NST = NCOL / 4;
    uint32_t near = 0;                     // bit = column of the tile: pair not excluded
    uint32_t rr[2][36];
    tf32_ld_cols<4>(d0, C0 / 4, rr[0]);
    tf32_ld_wait<4>(rr[0]);
#pragma unroll
    for (int st = 0; st < NST; st++) {
        const uint32_t* r = rr[st & 1];
        if (st + 1 < NST) {
            tf32_ld_cols<4>(d0, C0 / 4 + st + 1, rr[(st + 1) & 1]);   // in flight while this group is screened
        } else {                           // the last group is in registers: the buffer goes back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty_bar);
        }
        // ||S~||_F^2 of the four columns, two per instruction
        unsigned long long fp[2];
#pragma unroll
        for (int cp = 0; cp < 2; cp++) {
            unsigned long long acc = 0ull;
#pragma unroll
            for (int q = 0; q < 9; q++) {
                const unsigned long long v = OpsF2::pack(r[q * 4 + 2 * cp], r[q * 4 + 2 * cp + 1]);
                acc = OpsF2::fma(v, v, acc);
            }
            fp[cp] = acc;
        }
        float f4[4];
        OpsF2::unpack(fp[0], f4[0], f4[1]);
        OpsF2::unpack(fp[1], f4[2], f4[3]);
        // stage 1: Samuelson (see tf32_epilogue_tile_v4 for the constant and the sign-bit form of the decision)
        float lf4[4], ab4[4];
        uint32_t und = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int col = C0 + 4 * st + c;
            const float Bj = __shfl_sync(0xffffffffu, gvf, col);
            const float Dj = __shfl_sync(0xffffffffu, gvf, 16 + col);
            ab4[c] = row.Af + Bj;
            lf4[c] = fmaf(-row.Cf, Dj, ab4[c]);
            const float t = fmaf(3.00004f, f4[c], -(lf4[c] * lf4[c]));
            const uint32_t excl = (__float_as_uint(t) & ~__float_as_uint(lf4[c])) >> 31;
            und |= (excl ^ 1u) << c;
        }
        // stage 2: FP32 quartic sign test, per column pair with an undecided lane
#pragma unroll
        for (int cp = 0; cp < 2; cp++) {
            if (__any_sync(0xffffffffu, (und >> (2 * cp)) & 3u)) {
                unsigned long long S2[9];
#pragma unroll
                for (int q = 0; q < 9; q++) S2[q] = OpsF2::pack(r[q * 4 + 2 * cp], r[q * 4 + 2 * cp + 1]);
                // test point: a rigorous lower bound of the (already lowered) threshold eigenvalue; lf carries the
                // roundings of A_i + B_j and of the FMA: |lf - exact| <= u (|A_i + B_j| + |lf|) (1 + u)
                float lam[2];
#pragma unroll
                for (int h = 0; h < 2; h++)
                    lam[h] = fmaf(-2e-7f, fabsf(ab4[2 * cp + h]) + fabsf(lf4[2 * cp + h]), lf4[2 * cp + h]);
                unsigned long long p0, p1, p2, m0, m1, m2;
                const unsigned long long lam2 = OpsF2::pack(__float_as_uint(lam[0]), __float_as_uint(lam[1]));
                quartic32_values<OpsF2>(S2, fp[cp], lam2, p0, p1, p2);
                quartic32_margins<OpsF2>(p0, p1, p2, fp[cp], lam2, m0, m1, m2);
                float a1[2], b0[2], b1[2], b2[2];
                OpsF2::unpack(p1, a1[0], a1[1]);
                OpsF2::unpack(m0, b0[0], b0[1]);
                OpsF2::unpack(m1, b1[0], b1[1]);
                OpsF2::unpack(m2, b2[0], b2[1]);
#pragma unroll
                for (int h = 0; h < 2; h++)
                    if (quartic32_decide(lam[h], a1[h], b0[h], b1[h], b2[h])) und &= ~(1u << (2 * cp + h));
            }
        }
        near |= und << (C0 + 4 * st);
        if (st + 1 < NST) tf32_ld_wait<4>(rr[(st + 1) & 1]);
    }
    uint32_t valid = 0xffffu;
    if (j0 + 15 >= N) valid = (j0 >= N) ? 0u : (0xffffu >> (j0 + 16 - N));
    if (j0 <= i) valid &= (i - j0 >= 15) ? 0u : (0xffffu << (i - j0 + 1));
    return near & valid;
}

}  // namespace tsc
