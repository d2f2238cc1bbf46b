// tmem_probe.cu — measurement aids for the tcgen05 screen (NOT part of libtscode_b200.so).
//   tsc_probe_ld_layout : which (TMEM lane, column) lands in which (thread, register) for every tcgen05.ld shape
//   tsc_probe_ld_rate   : cycles to read a 128-lane x ncols accumulator tile with a given shape / warp count
#include "../../tscode_b200/csrc/screen_common.cuh"

namespace tsc {

// out[shape][half][thread 0..127][reg 0..3] = value read; value written = (lane << 16) | column
__global__ void __launch_bounds__(128, 1) ld_layout_kernel(uint32_t* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 64);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tb = slot;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < 64; c += 4) {
        const uint32_t v = ((uint32_t)(warp * 32 + lane) << 16);
        tmem_st_x4(tb + lane_addr + c, v | (c + 0), v | (c + 1), v | (c + 2), v | (c + 3));
    }
    tmem_st_wait();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    for (int half = 0; half < 2; half++) {
        const uint32_t ta = tb + lane_addr + ((uint32_t)(16 * half) << 16);
        uint32_t r[4];
        // shape 0: 16x256b.x1 (4 regs)
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ta));
        tmem_ld_wait();
        for (int q = 0; q < 4; q++) out[((0 * 2 + half) * 128 + threadIdx.x) * 4 + q] = r[q];
        // shape 1: 16x128b.x1 (2 regs)
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(ta));
        tmem_ld_wait();
        r[2] = r[3] = 0xffffffffu;
        for (int q = 0; q < 4; q++) out[((1 * 2 + half) * 128 + threadIdx.x) * 4 + q] = r[q];
        // shape 2: 16x64b.x1 (1 reg)
        asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(ta));
        tmem_ld_wait();
        r[1] = r[2] = r[3] = 0xffffffffu;
        for (int q = 0; q < 4; q++) out[((2 * 2 + half) * 128 + threadIdx.x) * 4 + q] = r[q];
        // shape 3: 16x256b.x2 (8 regs): only the second repetition is stored (regs 4..7)
        uint32_t s[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]) : "r"(ta));
        tmem_ld_wait();
        for (int q = 0; q < 4; q++) out[((3 * 2 + half) * 128 + threadIdx.x) * 4 + q] = s[4 + q];
        // shape 4: 16x32bx2.x1 with half-split offset 8 (1 reg)
        asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x1.b32 {%0}, [%1], 8;" : "=r"(r[0]) : "r"(ta));
        tmem_ld_wait();
        r[1] = r[2] = r[3] = 0xffffffffu;
        for (int q = 0; q < 4; q++) out[((4 * 2 + half) * 128 + threadIdx.x) * 4 + q] = r[q];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 64);
}

// every warp reads `ncols` columns of its lane quarter `reps` times; mode 0 = 32x32b.x4 loads (ncols/4 of them),
// mode 1 = 16x256b.x4 (two halves, 16 regs each: ncols/8 per half... see below), mode 2 = 16x256b.x8, mode 3 = 32x32b.x16
__global__ void __launch_bounds__(1024, 1) ld_rate_kernel(int ncols, int reps, int mode, long long* cycles_out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        if (mode == 0) {
            for (int c = 0; c < ncols; c += 16) {
                uint32_t v[16];
#pragma unroll
                for (int k = 0; k < 4; k++) asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[4 * k]), "=r"(v[4 * k + 1]), "=r"(v[4 * k + 2]), "=r"(v[4 * k + 3]) : "r"(tb + c + 4 * k));
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; k++) acc ^= v[k];
            }
        } else if (mode == 1) {
            // 16x256b.x4 = 16 lanes x 32 columns -> 16 regs; two halves cover 32 lanes x 32 columns
            for (int c = 0; c < ncols; c += 32) {
                uint32_t v[32];
#pragma unroll
                for (int h = 0; h < 2; h++)
                    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(v[16 * h + 0]), "=r"(v[16 * h + 1]), "=r"(v[16 * h + 2]), "=r"(v[16 * h + 3]),
                                   "=r"(v[16 * h + 4]), "=r"(v[16 * h + 5]), "=r"(v[16 * h + 6]), "=r"(v[16 * h + 7]),
                                   "=r"(v[16 * h + 8]), "=r"(v[16 * h + 9]), "=r"(v[16 * h + 10]), "=r"(v[16 * h + 11]),
                                   "=r"(v[16 * h + 12]), "=r"(v[16 * h + 13]), "=r"(v[16 * h + 14]), "=r"(v[16 * h + 15])
                                 : "r"(tb + c + ((uint32_t)(16 * h) << 16)));
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 32; k++) acc ^= v[k];
            }
        } else if (mode == 2) {
            // 16x256b.x2 = 16 lanes x 16 columns -> 8 regs
            for (int c = 0; c < ncols; c += 16) {
                uint32_t v[16];
#pragma unroll
                for (int h = 0; h < 2; h++)
                    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(v[8 * h + 0]), "=r"(v[8 * h + 1]), "=r"(v[8 * h + 2]), "=r"(v[8 * h + 3]),
                                   "=r"(v[8 * h + 4]), "=r"(v[8 * h + 5]), "=r"(v[8 * h + 6]), "=r"(v[8 * h + 7])
                                 : "r"(tb + c + ((uint32_t)(16 * h) << 16)));
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; k++) acc ^= v[k];
            }
        } else {
            for (int c = 0; c < ncols; c += 16) {
                uint32_t v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(tb + c));
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; k++) acc ^= v[k];
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles_out[0] = t1 - t0;
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

}  // namespace tsc

extern "C" int tsc_probe_ld_layout(uint32_t* out_dev, void* stream) {
    tsc::ld_layout_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(out_dev);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_probe_ld_rate(int32_t ncols, int32_t reps, int32_t mode, int32_t warps, long long* cycles_dev,
                                 uint32_t* sink_dev, void* stream) {
    if (warps < 4 || warps > 32 || ncols % 32 || ncols > 512) return (int)cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    tsc::ld_rate_kernel<<<sms, warps * 32, 0, (cudaStream_t)stream>>>(ncols, reps, mode, cycles_dev, sink_dev);
    TSC_CHECK_LAUNCH();
    return 0;
}
