// peaks.cu — self-measured FP64 peaks (register-resident DFMA / DMMA loops on every SM).
//
// MEASURED_PEAKS.json carries HBM and bf16 numbers only; the RMSD kernel is bounded by the FP64
// pipes, so bench.py measures their ceiling on the same GPU in the same run and quotes
// fractions "of self-measured FP64 peak" (SURVEY 8(d)).  Built into tools/probes/libtsc_probe.so: not part of the product library.
#include "../../tscode_b200/csrc/screen_common.cuh"

namespace tsc {

// kind::tf32 issue forms for the probe (the product only issues kind::f16)
template <bool ACC>
__device__ __forceinline__ void umma_tf32_ts_c(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}
template <bool ACC>
__device__ __forceinline__ void umma_tf32_ss_c(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}

constexpr int PK_CHAINS = 12;

__global__ void __launch_bounds__(512) peak_dfma_kernel(double* out, int iters, double seed) {
    double a[PK_CHAINS];
    const double x = 1.0 + seed * 1e-9, y = seed * 1e-7 * (threadIdx.x & 7);
#pragma unroll
    for (int c = 0; c < PK_CHAINS; c++) a[c] = c + threadIdx.x * 1e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < PK_CHAINS; c++) a[c] = fma(a[c], x, y);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < PK_CHAINS; c++) s += a[c];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(512) peak_dmma_kernel(double* out, int iters, double seed) {
    double c0[PK_CHAINS], c1[PK_CHAINS];
    const double a = 1.0 + seed * 1e-9 * threadIdx.x, b = 1e-3 * seed * (threadIdx.x & 3);
#pragma unroll
    for (int c = 0; c < PK_CHAINS; c++) { c0[c] = c; c1[c] = -c; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < PK_CHAINS; c++) dmma884(c0[c], c1[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < PK_CHAINS; c++) s += c0[c] + c1[c];
    if (s == 12345.678) out[0] = s;
}

// even warps DMMA, odd warps DFMA: do the two share a pipe?
__global__ void __launch_bounds__(512) peak_mixed_kernel(double* out, int iters, double seed) {
    const int warp = threadIdx.x >> 5;
    double c0[PK_CHAINS], c1[PK_CHAINS];
    const double a = 1.0 + seed * 1e-9 * threadIdx.x, b = 1e-3 * seed * (threadIdx.x & 3);
#pragma unroll
    for (int c = 0; c < PK_CHAINS; c++) { c0[c] = c; c1[c] = -c; }
    if (warp & 1) {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < PK_CHAINS; c++) c0[c] = fma(c0[c], a, b);
        }
    } else {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < PK_CHAINS; c++) dmma884(c0[c], c1[c], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < PK_CHAINS; c++) s += c0[c] + c1[c];
    if (s == 12345.678) out[0] = s;
}

}  // namespace tsc

// kind 0 = DFMA, 1 = DMMA, 2 = mixed (even warps DMMA, odd warps DFMA).
// Synchronous (times itself with CUDA events on `stream`).  flops_out = FP64 flop executed,
// for kind 2: flops_out[0] = DMMA part, flops_out[1] = DFMA part.  ms_out = elapsed.
extern "C" int tsc_bench_fp64(int32_t kind, int32_t iters, int32_t ctas_per_sm, int32_t threads, double* scratch,
                              double* flops_out, float* ms_out, void* stream) {
    using namespace tsc;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (threads <= 0 || threads > 512) threads = 512;
    if (ctas_per_sm <= 0) ctas_per_sm = 2;
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {        // rep 0 = warm-up
        cudaEventRecord(e0, st);
        if (kind == 0) peak_dfma_kernel<<<grid, threads, 0, st>>>(scratch, iters, 1.0);
        else if (kind == 1) peak_dmma_kernel<<<grid, threads, 0, st>>>(scratch, iters, 1.0);
        else peak_mixed_kernel<<<grid, threads, 0, st>>>(scratch, iters, 1.0);
        cudaEventRecord(e1, st);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) return (int)e;
    }
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double nthreads = (double)grid * threads, nwarps = nthreads / 32.0;
    const double per = (double)iters * 4 * PK_CHAINS;
    if (kind == 0) { flops_out[0] = 0; flops_out[1] = nthreads * per * 2.0; }
    else if (kind == 1) { flops_out[0] = nwarps * per * 512.0; flops_out[1] = 0; }
    else { flops_out[0] = (nwarps / 2) * per * 512.0; flops_out[1] = (nthreads / 2) * per * 2.0; }
    TSC_CHECK_LAUNCH();
    return 0;
}

// ---- tcgen05.mma issue-rate probe (measurement aid) -------------------------------------------------
// One CTA per SM; one thread issues `reps` rounds of (nsets x 3) kind::tf32 MMAs of shape 128 x N x 8,
// each set accumulating into its own TMEM region, operands from (zeroed) shared memory or TMEM.
// Answers: how many cycles does one such MMA cost as a function of N and of the distance between
// MMAs that accumulate into the same region?
namespace tsc {
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(int N, int nsets, int reps, int a_in_tmem,
                                                            long long* cycles_out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 * 32 + 256 * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tb = slot;
    if (warp == 1) {
        const uint32_t idesc = umma_idesc_tf32(128, N);
        const uint64_t ad = umma_desc_kmajor(smem_u32(sm), 128 * 16, 128);
        const uint64_t bd = umma_desc_kmajor(smem_u32(sm + 128 * 32), (uint32_t)N * 16, 128);
        // nsets < 0: |nsets| accumulators in total (instead of nsets * 3)
        const int nacc = nsets < 0 ? -nsets : nsets * 3;
        const long long t0 = clock64();
        if (elect_one()) {
            // straight-line issue, like the screen kernel: groups of `nacc` MMAs into distinct accumulators
            for (int r = 0; r < reps; r++) {
#pragma unroll 1
                for (int a0 = 0; a0 < nacc; a0 += 3) {
                    const uint32_t d = tb + 64 + (uint32_t)(a0 * N);
                    if (a_in_tmem) {
                        umma_tf32_ts_c<true>(d, tb, bd, idesc);
                        if (a0 + 1 < nacc) umma_tf32_ts_c<true>(d + N, tb + 8, bd, idesc);
                        if (a0 + 2 < nacc) umma_tf32_ts_c<true>(d + 2 * N, tb + 16, bd, idesc);
                    } else {
                        umma_tf32_ss_c<true>(d, ad, bd, idesc);
                        if (a0 + 1 < nacc) umma_tf32_ss_c<true>(d + N, ad, bd, idesc);
                        if (a0 + 2 < nacc) umma_tf32_ss_c<true>(d + 2 * N, ad, bd, idesc);
                    }
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) cycles_out[0] = t1 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}
}  // namespace tsc

// cycles for reps * nsets * 3 MMAs (SM 0), N in {16..256}, nsets*3*N + 64 <= 512
extern "C" int tsc_bench_umma(int32_t N, int32_t nsets, int32_t reps, int32_t a_in_tmem, long long* cycles_dev,
                              void* stream) {
    using namespace tsc;
    if (N % 16 || N < 16 || N > 256 || nsets == 0 || 64 + (nsets < 0 ? -nsets : nsets * 3) * N > 512)
        return (int)cudaErrorInvalidValue;
    const size_t smem = 128 * 32 + 256 * 32 + 1024;
    cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    umma_probe_kernel<<<sms, 128, smem, (cudaStream_t)stream>>>(N, nsets, reps, a_in_tmem, cycles_dev);
    TSC_CHECK_LAUNCH();
    return 0;
}
