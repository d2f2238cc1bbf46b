// screen_common.cuh — PTX wrappers and small helpers of the tcgen05 pre-screen (rmsd_screen.cu).
#pragma once
#include "tsc_common.cuh"
#include "tsc_math.cuh"

namespace tsc {

// Operand error bound of the FP16 screen: ||S~ - S||_F <= TF_EPS sqrt(G_i)' sqrt(G_j)' — two roundings of 2^-11
// (Cauchy-Schwarz over the atoms) plus a 7 % allowance for the tensor core's FP32 accumulation; see pack_screen_kernel.
constexpr double TF_EPS = 1.05e-3;

__device__ __forceinline__ void tmem_ld_x8_raw(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
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
// Lean issue forms for the MMA loop (kind::f16, FP16 operands, FP32 accumulate): the accumulate flag is a compile-time
// constant (ptxas folds the predicate) and descriptors are advanced with one integer add by the caller.
//   ts: D[tmem] (+)= A[tmem] * B[smem]^T (stationary operand read from tensor memory);  ss: both from shared memory
template <bool ACC>
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
template <bool ACC>
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}

// Per pair the threshold eigenvalue, lowered by the operand error bound, is
//     lam = 0.5 (1-1e-10) (G_i + G_j) - 0.5 e_thr - sqrt(3) eps sqrt(G_i)' sqrt(G_j)'.
// The epilogue works with a guaranteed LOWER bound lf = A_i + B_j - C_i D_j from directed-rounded FP32 terms:
// per row A_i = float_rd(hs G_i - 0.5 e_thr), C_i = float_ru(sqrt(3) eps sqrt(G_i)'); per column (pack)
// B_j = float_rd(hs G_j), D_j = float_ru(sqrt(G_j)').
struct ScRow {                 // per-thread (row i) constants
    float Af, Cf;              // A_i = float_rd(hs G_i - 0.5 e_thr),  C_i = float_ru(sqrt(3) eps sqrt(G_i))
    double hi, ci;             // the same in FP64: hi = hs G_i - 0.5 e_thr,  ci = -sqrt(3) eps sqrt(G_i)
};
__device__ __forceinline__ ScRow screen_row_consts(double Gi, double sGi, double e_thr) {
    const double hs = 0.5 * (1.0 - 1e-10), cc = 1.7320508075688772 * TF_EPS;
    ScRow r;
    r.hi = fma(hs, Gi, -0.5 * e_thr);
    r.ci = -cc * sGi;
    r.Af = __double2float_rd(r.hi);
    r.Cf = __double2float_ru(cc * sGi);
    return r;
}
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

}  // namespace tsc
