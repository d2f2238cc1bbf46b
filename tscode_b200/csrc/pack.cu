// pack.cu — heavy-atom gather + AoS -> tiled-SoA repack + per-conformer squared norms.
//
// Replaces the host-side list comprehension `structures[:, atomnos != 1]`
// (tscode/rmsd_pruning.py:178-179) and lays the ensemble out once in the form every later
// kernel streams with bulk-TMA (see tsc_common.cuh for the layout).
//
// HBM-bound: reads 24*A bytes and writes 24*M_pad bytes per conformer, once.
#include "tsc_common.cuh"

namespace tsc {

__global__ void __launch_bounds__(256) pack_kernel(const double* __restrict__ S, int64_t N, int A,
                                                   const int32_t* __restrict__ heavy_idx, int M, int nslab,
                                                   int64_t nb_pad, double* __restrict__ packed,
                                                   double* __restrict__ G, int block_begin) {
    const int b = block_begin + blockIdx.x;       // conformer block
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Mp = nslab * KS;
    for (int c = warp; c < CB; c += 8) {
        const int64_t i = (int64_t)b * CB + c;
        const bool live = i < N;
        const double* src = S + (live ? i : 0) * (int64_t)A * 3;
        double g = 0.0;
        for (int m = lane; m < Mp; m += 32) {
            double x = 0.0, y = 0.0, z = 0.0;
            if (live && m < M) {
                const double* a = src + (int64_t)heavy_idx[m] * 3;
                x = a[0]; y = a[1]; z = a[2];
                g = fma(x, x, fma(y, y, fma(z, z, g)));
            }
            const int64_t o = (((int64_t)(m / KS) * nb_pad + b) * 3) * (CB * KS) + c * KS + (m % KS);
            packed[o] = x;
            packed[o + CB * KS] = y;
            packed[o + 2 * CB * KS] = z;
        }
        g = warp_sum(g);
        if (lane == 0) G[i] = g;                  // G has nb_pad*CB entries
    }
}

}  // namespace tsc

extern "C" int tsc_pack(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M,
                        double* packed, double* G, void* stream) {
    if (N <= 0 || M <= 0) return 0;
    const int64_t nb_pad = tsc::num_blocks_padded(N);
    tsc::pack_kernel<<<(unsigned)nb_pad, 256, 0, (cudaStream_t)stream>>>(S, N, A, heavy_idx, M, tsc::num_slabs(M),
                                                                         nb_pad, packed, G, 0);
    TSC_CHECK_LAUNCH();
    return 0;
}

// The same for the conformer blocks (32 rows each) [block_begin, block_end) only: lets the caller repack a chunk of
// the ensemble as soon as its host-to-device copy has landed while later chunks are still in flight.
extern "C" int tsc_pack_blocks(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M,
                               double* packed, double* G, int64_t block_begin, int64_t block_end, void* stream) {
    if (N <= 0 || M <= 0 || block_end <= block_begin) return 0;
    const int64_t nb_pad = tsc::num_blocks_padded(N);
    if (block_end > nb_pad) block_end = nb_pad;
    tsc::pack_kernel<<<(unsigned)(block_end - block_begin), 256, 0, (cudaStream_t)stream>>>(
        S, N, A, heavy_idx, M, tsc::num_slabs(M), nb_pad, packed, G, (int)block_begin);
    TSC_CHECK_LAUNCH();
    return 0;
}
