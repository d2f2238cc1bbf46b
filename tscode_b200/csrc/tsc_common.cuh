// tsc_common.cuh — layout constants, PTX wrappers (mbarrier, bulk-TMA, DMMA) and error plumbing
// shared by the sm_100a kernels of tscode_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsc {

// ------------------------------------------------------------------------------------------
// Packed ensemble layout in HBM ("tiled SoA", produced by pack.cu, consumed by every RMSD kernel)
//
//   packed[slab s][block b][comp a][conformer c][atom k]      doubles
//     CB = 32 conformers per block, KS = 20 atoms per slab, comp in {x,y,z}
//     element (conformer i, heavy atom m, comp a) lives at
//       (((m / KS) * nb_pad + i / CB) * 3 + a) * (CB * KS) + (i % CB) * KS + (m % KS)
//   nb_pad = number of conformer blocks rounded up to even, nslab = ceil(M / KS); padding
//   conformers / atoms are zero (they contribute nothing to covariances or norms).
//
// One (slab, block) chunk is 3*32*20 doubles = 15 360 B contiguous: the unit of a bulk-TMA
// copy, and already the shared-memory image the MMA fragments are read from.  The row stride
// of 20 doubles (== 4 mod 16) makes the DMMA fragment loads (lane -> conformer lane/4, atom
// lane%4) hit 16 distinct 8-byte banks per half-warp.
// ------------------------------------------------------------------------------------------
constexpr int CB = 32;
constexpr int KS = 20;
constexpr int CHUNK_D = 3 * CB * KS;          // doubles per (slab, block) chunk
constexpr int CHUNK_BYTES = CHUNK_D * 8;      // 15360

__host__ __device__ inline int64_t num_blocks_padded(int64_t N) {
    int64_t nb = (N + CB - 1) / CB;
    nb += nb & 1;
    return nb < 2 ? 2 : nb;
}
__host__ __device__ inline int num_slabs(int M) { return (M + KS - 1) / KS; }
__host__ __device__ inline int64_t packed_index(int64_t i, int m, int a, int64_t nb_pad) {
    return (((int64_t)(m / KS) * nb_pad + i / CB) * 3 + a) * (CB * KS) + (i % CB) * KS + (m % KS);
}

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// non-blocking form: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// 1-D bulk TMA: global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col)   (SASS: DMMA.8x8x4)
//   A: lane holds A[lane>>2][lane&3];  B: lane holds B[lane&3][lane>>2];
//   C/D: lane holds C[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// One elected lane of a fully converged warp.  tcgen05.mma / tcgen05.commit / cp.async.bulk are
// issued through the uniform datapath: inside a plain `if (lane == 0)` region ptxas cannot prove
// uniformity and wraps EVERY such instruction in an ELECT / BRA.U.ANY retry loop (~90 cycles per
// MMA measured); a region guarded by elect.sync compiles to straight-line code.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}
// ---- tcgen05 (5th-gen tensor cores, accumulators in TMEM) ------------------------------------
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): 8-row x 16-byte core
// matrices stored contiguously (128 B); SBO = byte distance between core matrices along M/N,
// LBO = byte distance between the two 16-byte K chunks one MMA consumes (layout checked against
// cute/arch/mma_sm100_desc.hpp: start[0,14) LBO[16,30) SBO[32,46) version[46,48)=1 layout[61,64)=0).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::tf32, FP32 accumulate, A and B K-major:
// c_format[4,6)=1 (F32), a_format[7,10)=2 (TF32), b_format[10,13)=2, n_dim[17,23)=N>>3, m_dim[24,29)=M>>4.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 with FP16 operands (a_format = b_format = 0), FP32 accumulate: K = 16 per instruction.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 lanes x 8 consecutive 32-bit columns: thread t of warp w gets TMEM lane 32*(w%4)+t (SASS: LDTM)
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float (&v)[8]) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Reserve `n` slots of a fixed-capacity device list whose header word is the running entry count (candidate list,
// confirmed-pair list).  Returns the first slot, or `cap` — so that the caller's `slot < cap` test skips every store —
// once the header has left [0, cap]: it then stays out of range (that is how consumers see the overflow) and is not
// advanced any further, so it can neither turn negative nor wrap around to a value that looks like a complete list,
// however many entries a very redundant ensemble produces (a header only ever overshoots by the few reservations
// that were already in flight; hosts keep cap below 2^31 - 2^24).
__device__ __forceinline__ int64_t list_reserve(int* header, int n, int64_t cap) {
    const int cur = *reinterpret_cast<volatile int*>(header);
    if (cur < 0 || (int64_t)cur > cap) return cap;
    const int base = atomicAdd(header, n);
    return base < 0 ? cap : (int64_t)base;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif

}  // namespace tsc

#define TSC_CHECK_LAUNCH()                                  \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)
