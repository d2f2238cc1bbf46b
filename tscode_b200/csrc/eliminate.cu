// eliminate.cu — the k-ladder greedy elimination of prune_conformers_rmsd on bit rows.
//
// One ladder round of the reference (tscode/rmsd_pruning.py:123-162 calling :81-121 and
// :43-79) is, for every active row i of every chunk [first, last):
//     walk j = i+1 .. last-1 over rows active at ROUND START;
//       (first, first + j - i) in cache  -> keep i, stop            (:65-67)
//       sim(i, j)                        -> drop i, emit that key   (:75-77)
//     nothing found                      -> keep i
// The cache only grows between rounds (:204) and its key depends on (chunk start, offset)
// rather than on the pair — a quirk that changes the result and is reproduced exactly
// (SURVEY.md fact 4, Appendix A.2).  Rows are independent within a round, so a round is a
// "first hit" scan per row over   active & (sim_row | cache_shifted)   — pure bit work.
//
//   tsc_elim_cachebits : key list -> N-bit map for this round's chunking (bit s set iff
//                        (first, s) is cached and `first` is a chunk start of this round)
//   tsc_elim_round     : one warp per row, 32 words per step, early exit on first hit
//   tsc_elim_commit    : per-row state -> active words, byte mask, active count, key-list append
// Latency-bound (a few hundred kB per round); reported as time, not against a roofline.
//
// Device-side gating.  Whether a round of the ladder runs depends on the number of structures
// still active (`k == 1 or 20*k < count_nonzero(mask)`, rmsd_pruning.py:192).  Every kernel takes
// a pointer to that count as it stood BEFORE the round and returns immediately when the gate is
// closed, so the host can enqueue every candidate round without reading anything back (no
// synchronisation until the final mask) and all ranks of a multi-GPU run take the same decision.
#include "tsc_common.cuh"

namespace tsc {

__device__ __forceinline__ bool gate_open(const int32_t* gate, int64_t k) {
    return gate == nullptr || k == 1 || 20 * k < (int64_t)*gate;
}

__device__ __forceinline__ void chunk_of(int64_t i, int64_t N, int64_t cs, int64_t k, int64_t& first,
                                         int64_t& last) {
    if (cs <= 0) { first = 0; last = N; return; }     // int(N // k) == 0: only the last chunk is non-empty
    int64_t c = i / cs;
    if (c > k - 1) c = k - 1;
    first = c * cs;
    last = (c == k - 1) ? N : first + cs;
}

__global__ void elim_cachebits_kernel(const int32_t* __restrict__ key_first, const int32_t* __restrict__ key_second,
                                      const int32_t* __restrict__ n_keys, int64_t N, int64_t cs, int64_t k,
                                      uint32_t* cachebits, const int32_t* __restrict__ gate) {
    if (!gate_open(gate, k)) return;
    const int n = *n_keys;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int64_t f = key_first[t], s = key_second[t];
        int64_t last;
        if (cs <= 0) {
            if (f != 0) continue;
            last = N;
        } else {
            if (f % cs != 0) continue;
            const int64_t c = f / cs;
            if (c >= k) continue;
            last = (c == k - 1) ? N : f + cs;
        }
        if (s < last) atomicOr(&cachebits[s >> 5], 1u << (s & 31));
    }
}

__global__ void __launch_bounds__(256) elim_round_kernel(const uint32_t* __restrict__ sim_bits, int64_t W,
                                                         const int32_t* __restrict__ row_blocks, int n_rb,
                                                         const uint32_t* __restrict__ active,
                                                         const uint32_t* __restrict__ cachebits, int64_t N,
                                                         int64_t cs, int64_t k, int32_t* __restrict__ row_state,
                                                         const int32_t* __restrict__ gate) {
    if (!gate_open(gate, k)) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t Wc = (N + 31) >> 5;
    for (int64_t row = warp_g; row < (int64_t)n_rb * CB; row += nwarps) {
        const int64_t i = (int64_t)row_blocks[row / CB] * CB + (row % CB);
        if (i >= N) continue;
        if (!((active[i >> 5] >> (i & 31)) & 1u)) {
            if (lane == 0) row_state[i] = -2;                    // inactive
            continue;
        }
        int64_t first, last;
        chunk_of(i, N, cs, k, first, last);
        const int64_t delta = i - first;
        const uint32_t* srow = sim_bits + row * W;
        const int64_t wbeg = (i + 1) >> 5, wend = (last - 1) >> 5;    // inclusive; empty if i+1 >= last
        int keep = 1;
        int32_t key = -1;
        if (i + 1 < last) {
            for (int64_t w0 = wbeg; w0 <= wend; w0 += 32) {
                const int64_t w = w0 + lane;
                uint32_t hit = 0, cbits = 0;
                if (w <= wend) {
                    const uint32_t a = active[w];
                    const uint32_t s = srow[w];
                    const int64_t pos = w * 32 - delta;               // cache bit index of this word's bit 0
                    const int64_t lo = pos >> 5;
                    const uint32_t sh = (uint32_t)(pos & 31);
                    const uint32_t c_lo = (lo >= 0 && lo < Wc) ? cachebits[lo] : 0u;
                    const uint32_t c_hi = (lo + 1 >= 0 && lo + 1 < Wc) ? cachebits[lo + 1] : 0u;
                    cbits = __funnelshift_r(c_lo, c_hi, sh);
                    uint32_t valid = 0xffffffffu;
                    if (w * 32 <= i) valid &= (i - w * 32 >= 31) ? 0u : (0xffffffffu << (i - w * 32 + 1));
                    if (w * 32 + 31 >= last) valid &= 0xffffffffu >> (w * 32 + 31 - last + 1);
                    hit = a & (s | cbits) & valid;
                }
                const uint32_t any = __ballot_sync(0xffffffffu, hit != 0u);
                if (any) {
                    const int src = __ffs(any) - 1;
                    const uint32_t h = __shfl_sync(0xffffffffu, hit, src);
                    const uint32_t c = __shfl_sync(0xffffffffu, cbits, src);
                    const int b = __ffs(h) - 1;
                    const int64_t j = (w0 + src) * 32 + b;
                    if (!((c >> b) & 1u)) {          // cache is consulted first (:65-67), then sim (:70-77)
                        keep = 0;
                        key = (int32_t)(first + j - i);
                    }
                    break;
                }
            }
        }
        if (lane == 0) row_state[i] = keep ? -1 : key;          // -1 kept, >= 0 dropped with this cache key
    }
}

__global__ void __launch_bounds__(256) elim_commit_kernel(const int32_t* __restrict__ row_state, int64_t N, int64_t cs,
                                                          int64_t k, uint32_t* __restrict__ active_out,
                                                          uint8_t* __restrict__ mask_out, int32_t* key_first,
                                                          int32_t* key_second, int32_t* n_keys,
                                                          const int32_t* __restrict__ gate, int32_t* n_active_out) {
    if (!gate_open(gate, k)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *n_active_out = *gate;     // round skipped: count carries over
        return;
    }
    const int lane = threadIdx.x & 31;
    const int64_t Npad = (N + 31) & ~int64_t(31);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Npad; i += (int64_t)gridDim.x * blockDim.x) {
        const bool live = i < N;
        const int32_t st = live ? row_state[i] : -2;
        const bool m = st == -1;
        const uint32_t word = __ballot_sync(0xffffffffu, m);
        if (lane == 0) {
            active_out[i >> 5] = word;
            if (word) atomicAdd(n_active_out, __popc(word));
        }
        if (live) {
            mask_out[i] = (uint8_t)m;
            if (st >= 0) {
                int64_t first, last;
                chunk_of(i, N, cs, k, first, last);
                const int slot = atomicAdd(n_keys, 1);
                key_first[slot] = (int32_t)first;
                key_second[slot] = st;
            }
        }
    }
}

}  // namespace tsc

extern "C" int tsc_elim_cachebits(const int32_t* key_first, const int32_t* key_second, const int32_t* n_keys,
                                  int64_t N, int64_t cs, int64_t k, uint32_t* cachebits, const int32_t* gate,
                                  void* stream) {
    if (N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(cachebits, 0, (size_t)((N + 31) / 32) * 4, st);
    if (e != cudaSuccess) return (int)e;
    tsc::elim_cachebits_kernel<<<148, 256, 0, st>>>(key_first, key_second, n_keys, N, cs, k, cachebits, gate);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_elim_round(const uint32_t* sim_bits, const int32_t* row_blocks, int32_t n_rb,
                              const uint32_t* active_words, const uint32_t* cachebits, int64_t N, int64_t cs,
                              int64_t k, int32_t* row_state, const int32_t* gate, void* stream) {
    if (N <= 0 || n_rb <= 0) return 0;
    const int64_t W = tsc::num_blocks_padded(N);
    int64_t rows = (int64_t)n_rb * tsc::CB;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 64) blocks = 148 * 64;
    tsc::elim_round_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        sim_bits, W, row_blocks, n_rb, active_words, cachebits, N, cs, k, row_state, gate);
    TSC_CHECK_LAUNCH();
    return 0;
}

extern "C" int tsc_elim_commit(const int32_t* row_state, int64_t N, int64_t cs, int64_t k,
                               uint32_t* active_words_out, uint8_t* mask_out, int32_t* key_first,
                               int32_t* key_second, int32_t* n_keys, const int32_t* gate, int32_t* n_active_out,
                               void* stream) {
    if (N <= 0) return 0;
    int64_t blocks = (N + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tsc::elim_commit_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        row_state, N, cs, k, active_words_out, mask_out, key_first, key_second, n_keys, gate, n_active_out);
    TSC_CHECK_LAUNCH();
    return 0;
}
