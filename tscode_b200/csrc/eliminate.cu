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


// ==========================================================================================
// Fused ladder on a PAIR LIST: the whole k-ladder in ONE persistent cooperative launch.
//
// The bit-row kernels above need three launches per round (33 for C3) and, on several GPUs, an
// all-gather per round; at the speed of the tcgen05 screen those latency-bound rounds had become
// 6 % of a single-GPU prune and > 50 % of an 8-GPU one.  The similar pairs are sparse (C3: 2.5e5 of
// 1.25e9), so tsc_rmsd_verify also emits them as an (i, j) list; lists of all ranks are all-gathered
// ONCE and every rank runs this kernel redundantly on the complete list.  Per round:
//   phase A  key list -> cache bitmap of this round's chunking; pair list -> first_sim[i] = smallest
//            active j > i of i's chunk similar to i (atomicMin); active_out := active_in
//   phase B  one warp per active row: first active j in (i, min(first_sim, last-1)] whose cache bit
//            (first, first + j - i) is set -> keep (cache is consulted before sim, :65-77); else
//            first_sim found -> drop + emit key; else keep.  Bitmaps are staged in shared memory.
// separated by a grid-wide barrier (monotone counter; bounded spin so that a lost CTA can never hang
// the device).  The reference's gate (`k == 1 or 20 k < active`) is evaluated by every CTA from the
// same global counter.  Results are identical to the bit-row path (tests run both).
// ==========================================================================================
constexpr int EF_THREADS = 1024;
constexpr int EF_HDR = 64;            // int32 words: [0] barrier, [1] abort, [2] n_keys, [3] status, [4] rounds ran,
                                      // [5] dropped this round, [6]/[7] cache-bitmap-non-empty flags, [8] n_active
constexpr int EF_MAX_LADDER = 24;
constexpr long long EF_SPIN_LIMIT = 4000000000LL;     // ~2 s of SM clocks

struct ElimFusedParams {
    const int2* lists;        // n_lists blocks of `stride` int2: block[0].x = pair count, pairs from block[1]
    int n_lists;
    int64_t stride;
    int64_t N;
    int32_t* ws;              // workspace, tsc_elim_fused_ws_words(N) int32 (header zeroed by the entry point)
    uint8_t* out;             // N mask bytes, padded to a multiple of 4, then 64 int32 of info
    int n_ladder;
    int gate;
    int64_t ladder[EF_MAX_LADDER];
    // several ranks, peer-written lists (tsc_pairs_push): block l of `lists` is complete once flags[l] == epoch
    const int32_t* flags;     // NULL: the lists are complete at launch
    int32_t epoch;
};

__device__ __forceinline__ long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// returns false if the barrier was abandoned (timeout / another CTA aborted)
__device__ __forceinline__ bool ef_grid_barrier(uint32_t* hdr, uint32_t& epoch, int* s_flag) {
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&hdr[0], 1u);
        const uint32_t target = epoch * gridDim.x;
        const long long t0 = clock64();
        int ok = 1;
        uint32_t spins = 0;
        while (ld_acquire_u32(&hdr[0]) < target) {
            if ((++spins & 63u) == 0u && (ld_acquire_u32(&hdr[1]) != 0u || clock64() - t0 > EF_SPIN_LIMIT)) {
                atomicExch(&hdr[1], 1u);
                ok = 0;
                break;
            }
        }
        __threadfence();
        *s_flag = ok;
    }
    __syncthreads();
    return *s_flag != 0;
}

constexpr int EF_SMEM_CHUNKS = 4096;       // per-chunk cache windows staged in shared memory up to this many chunks

// chunk of row i in 32-bit arithmetic (N < 2^31; cs >= 1 whenever a round's gate is open)
__device__ __forceinline__ void chunk_of32(uint32_t i, uint32_t N, uint32_t cs, uint32_t k, uint32_t& c, uint32_t& first,
                                           uint32_t& last) {
    c = i / cs;
    if (c > k - 1) c = k - 1;
    first = c * cs;
    last = (c == k - 1) ? N : first + cs;
}

__global__ void __launch_bounds__(EF_THREADS, 1) elim_fused_kernel(const ElimFusedParams p) {
    extern __shared__ uint32_t ef_smem[];
    __shared__ int s_flag;
    const uint32_t N = (uint32_t)p.N;
    const uint32_t Wc = (N + 31) >> 5;
    const uint32_t NC = N / 16 + 2;                      // chunks of a round: k < N / gate
    uint32_t* hdr = reinterpret_cast<uint32_t*>(p.ws);
    int32_t* hist = p.ws + EF_HDR;                       // active count after round r (at r + 1)
    int32_t* rounds_k = hist + 32;
    uint32_t* act[2] = {reinterpret_cast<uint32_t*>(rounds_k + 32), nullptr};
    act[1] = act[0] + (Wc + 2);
    uint32_t* cb[2] = {act[1] + (Wc + 2), nullptr};
    cb[1] = cb[0] + (Wc + 2);
    int32_t* first_sim = reinterpret_cast<int32_t*>(cb[1] + (Wc + 2));
    int32_t* key_first = first_sim + N;
    int32_t* key_second = key_first + N;
    int32_t* cmin[2] = {key_second + N, nullptr};        // per chunk: smallest / largest cached offset
    cmin[1] = cmin[0] + NC;
    int32_t* cmax[2] = {cmin[1] + NC, nullptr};
    cmax[1] = cmax[0] + NC;
    uint32_t* s_act = ef_smem;
    uint32_t* s_cb = ef_smem + (Wc + 2);
    int32_t* s_cmin = reinterpret_cast<int32_t*>(s_cb + (Wc + 2));
    int32_t* s_cmax = s_cmin + EF_SMEM_CHUNKS;
    int32_t* info = reinterpret_cast<int32_t*>(p.out + (((int64_t)N + 3) & ~int64_t(3)));

    const int lane = threadIdx.x & 31;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gthreads = gridDim.x * blockDim.x;
    const uint32_t gwarp = gtid >> 5, gwarps = gthreads >> 5;
    uint32_t epoch = 0;
    constexpr int32_t INF = 0x7fffffff;
    const long long t_start = globaltimer_ns();          // info[32..]: ns since start at phase boundaries (thread 0)
    int n_stamp = 0;
#define EF_STAMP() do { if (gtid == 0 && n_stamp < 30) info[32 + n_stamp] = (int32_t)(globaltimer_ns() - t_start); n_stamp++; } while (0)

    // ---- peer-written lists: wait until every rank has published its block (system-scope acquire; bounded) ----
    if (p.flags) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            int ok = 1;
            for (int l = 0; l < p.n_lists && ok; l++) {
                uint32_t spins = 0;
                while (true) {
                    int32_t v;
                    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p.flags + l) : "memory");
                    if (v == p.epoch) break;
                    if ((++spins & 63u) == 0u && clock64() - t0 > EF_SPIN_LIMIT) { ok = 0; break; }
                }
            }
            s_flag = ok;
        }
        __syncthreads();
        if (!s_flag) {
            if (gtid == 0) { info[1] = 0; __threadfence(); info[0] = 3; }      // 3: a peer never published its list
            return;
        }
    }
    // ---- overflow check (uniform: every thread reads the same headers) ----
    bool overflow = false;
    for (int l = 0; l < p.n_lists; l++) {
        const int64_t cnt = p.lists[l * p.stride].x;
        overflow |= cnt < 0 || cnt > p.stride - 1;           // any count outside [0, capacity] is an overflow
    }
    if (overflow) {
        if (gtid == 0) { info[1] = 0; __threadfence(); info[0] = 1; }
        return;
    }
    // ---- init ----
    for (uint32_t w = gtid; w < Wc + 2; w += gthreads) {
        uint32_t v = 0;
        if (w < Wc) v = (w * 32 + 31 < N) ? 0xffffffffu : (0xffffffffu >> (w * 32 + 32 - N));
        act[0][w] = v; act[1][w] = v;
        cb[0][w] = 0; cb[1][w] = 0;
    }
    for (uint32_t i = gtid; i < N; i += gthreads) first_sim[i] = INF;
    for (uint32_t c = gtid; c < NC; c += gthreads) { cmin[0][c] = INF; cmin[1][c] = INF; cmax[0][c] = 0; cmax[1][c] = 0; }
    if (!ef_grid_barrier(hdr, epoch, &s_flag)) return;
    EF_STAMP();

    uint32_t n_active = N;
    int cur = 0, r = 0;
    for (int li = 0; li < p.n_ladder; li++) {
        if (!(p.ladder[li] == 1 || (int64_t)p.gate * p.ladder[li] < (int64_t)n_active)) continue;
        const uint32_t k = (uint32_t)p.ladder[li];       // gate open => k <= N
        const uint32_t cs = N / k;
        const uint32_t* a_in = act[cur];
        uint32_t* a_out = act[cur ^ 1];
        uint32_t* cbits = cb[cur];
        // ---------------- phase A ----------------
        // round-start mask -> shared memory (valid for both phases: drops go to a_out) and -> a_out
        for (uint32_t w = threadIdx.x; w < Wc + 2; w += blockDim.x) s_act[w] = __ldcg(&a_in[w]);
        for (uint32_t w = gtid; w < Wc; w += gthreads) a_out[w] = __ldcg(&a_in[w]);
        {
            const uint32_t n_keys = __ldcg(&hdr[2]);
            for (uint32_t t = gtid; t < n_keys; t += gthreads) {
                const uint32_t f = (uint32_t)__ldcg(&key_first[t]), s = (uint32_t)__ldcg(&key_second[t]);
                const uint32_t c = f / cs;
                if (c * cs != f || c >= k) continue;                 // `first` is not a chunk start of this round
                const uint32_t last = (c == k - 1) ? N : f + cs;
                if (s < last) {
                    atomicOr(&cbits[s >> 5], 1u << (s & 31));
                    atomicMin(&cmin[cur][c], (int32_t)(s - f));
                    atomicMax(&cmax[cur][c], (int32_t)(s - f));
                    hdr[6 + cur] = 1u;
                }
            }
        }
        __syncthreads();
        for (int l = 0; l < p.n_lists; l++) {
            const int2* blk = p.lists + (int64_t)l * p.stride;
            const uint32_t n = (uint32_t)blk[0].x;
            for (uint32_t t = gtid; t < n; t += gthreads) {
                const int2 ij = blk[1 + t];
                const uint32_t i = (uint32_t)ij.x, j = (uint32_t)ij.y;
                if (!((s_act[i >> 5] >> (i & 31)) & 1u) || !((s_act[j >> 5] >> (j & 31)) & 1u)) continue;
                uint32_t c, first, last;
                chunk_of32(i, N, cs, k, c, first, last);
                if (j < last) atomicMin(&first_sim[i], (int32_t)j);
            }
        }
        EF_STAMP();
        if (!ef_grid_barrier(hdr, epoch, &s_flag)) return;
        EF_STAMP();
        // ---------------- phase B ----------------
        const bool any_cb = __ldcg(&hdr[6 + cur]) != 0u;
        const bool win_smem = k <= (uint32_t)EF_SMEM_CHUNKS;
        if (any_cb) {
            for (uint32_t w = threadIdx.x; w < Wc + 2; w += blockDim.x) s_cb[w] = __ldcg(&cbits[w]);
            if (win_smem)
                for (uint32_t c = threadIdx.x; c < k; c += blockDim.x) {
                    s_cmin[c] = __ldcg(&cmin[cur][c]);
                    s_cmax[c] = __ldcg(&cmax[cur][c]);
                }
        }
        __syncthreads();
        {   // the other cache bitmap / offset windows (used by the previous round) are cleared for the next one
            uint32_t* nb = cb[cur ^ 1];
            for (uint32_t w = gtid; w < Wc + 2; w += gthreads) nb[w] = 0;
            for (uint32_t c = gtid; c < NC; c += gthreads) { cmin[cur ^ 1][c] = INF; cmax[cur ^ 1][c] = 0; }
            if (gtid == 0) hdr[6 + (cur ^ 1)] = 0u;
        }
        uint32_t dropped = 0;
        // every lane owns one row of a batch of 32 (rows dealt round-robin to the warps of the grid): all global
        // reads of the batch are in flight together; only rows whose chunk has cached offsets inside the row's
        // window are scanned, warp-cooperatively, over that window
        for (uint64_t b0 = gwarp; b0 < N; b0 += (uint64_t)gwarps * 32) {
            const uint64_t i64 = b0 + (uint64_t)lane * gwarps;
            const uint32_t i = (uint32_t)(i64 < N ? i64 : 0);
            const bool act_i = i64 < N && ((s_act[i >> 5] >> (i & 31)) & 1u);
            int32_t js = INF;
            uint32_t c = 0, first = 0, last = 0;
            int64_t lo_j = 1, hi_j = 0;
            if (act_i) {
                js = __ldcg(&first_sim[i]);
                chunk_of32(i, N, cs, k, c, first, last);
                if (any_cb) {
                    const int64_t omin = win_smem ? s_cmin[c] : __ldcg(&cmin[cur][c]);
                    const int64_t omax = win_smem ? s_cmax[c] : __ldcg(&cmax[cur][c]);
                    const int64_t limit = (js != INF) ? (int64_t)js : (int64_t)last - 1;        // inclusive
                    lo_j = (int64_t)i + (omin < 1 ? 1 : omin);
                    hi_j = ((int64_t)i + omax < limit) ? (int64_t)i + omax : limit;
                }
                if (js != INF) first_sim[i] = INF;
            }
            bool cache_hit = false;
            uint32_t todo = __ballot_sync(0xffffffffu, act_i && lo_j <= hi_j);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t ri = __shfl_sync(0xffffffffu, i, src), rfirst = __shfl_sync(0xffffffffu, first, src);
                const int64_t rlo = __shfl_sync(0xffffffffu, lo_j, src), rhi = __shfl_sync(0xffffffffu, hi_j, src);
                const int64_t delta = ri - rfirst;
                bool hit_any = false;
                for (int64_t w0 = rlo >> 5; w0 <= (rhi >> 5) && !hit_any; w0 += 32) {
                    const int64_t w = w0 + lane;
                    uint32_t hit = 0;
                    if (w <= (rhi >> 5)) {
                        const int64_t pos = w * 32 - delta;          // cache bit index of this word's bit 0 (>= -31)
                        const int64_t lo = pos >> 5;
                        const uint32_t sh = (uint32_t)(pos & 31);
                        const uint32_t c_lo = (lo >= 0) ? s_cb[lo] : 0u;
                        const uint32_t c_hi = s_cb[lo + 1];
                        uint32_t valid = 0xffffffffu;
                        if (w * 32 < rlo) valid &= (rlo - w * 32 >= 32) ? 0u : (0xffffffffu << (rlo - w * 32));
                        if (w * 32 + 31 > rhi) valid &= 0xffffffffu >> (w * 32 + 31 - rhi);
                        hit = s_act[w] & __funnelshift_r(c_lo, c_hi, sh) & valid;
                    }
                    hit_any = __any_sync(0xffffffffu, hit != 0u);
                }
                if (lane == src) cache_hit = hit_any;
            }
            const bool drop = act_i && !cache_hit && js != INF;
            if (drop) {
                atomicAnd(&a_out[i >> 5], ~(1u << (i & 31)));
                const uint32_t slot = atomicAdd(&hdr[2], 1u);
                key_first[slot] = (int32_t)first;
                key_second[slot] = (int32_t)(first + (uint32_t)js - i);
            }
            dropped += __popc(__ballot_sync(0xffffffffu, drop));
        }
        if (lane == 0 && dropped) atomicAdd(&hdr[8 + 1 + r], dropped);
        if (r < 3) EF_STAMP();
        if (!ef_grid_barrier(hdr, epoch, &s_flag)) return;
        if (r < 3) EF_STAMP();
        n_active -= __ldcg(&hdr[8 + 1 + r]);
        if (gtid == 0) { rounds_k[r] = (int32_t)k; hist[r + 1] = (int32_t)n_active; }
        cur ^= 1;
        r++;
    }
    // ---- result ----
    const uint32_t* a_fin = act[cur];
    for (uint32_t i = gtid; i < N; i += gthreads) p.out[i] = (uint8_t)((__ldcg(&a_fin[i >> 5]) >> (i & 31)) & 1u);
    if (gtid == 0) {
        info[1] = r;
        info[2] = (int32_t)n_active;
        info[3] = (int32_t)hdr[2];
        for (int t = 0; t < EF_MAX_LADDER; t++) info[8 + t] = (t < r) ? rounds_k[t] : 0;
        __threadfence();
        info[0] = 0;                 // 0 = complete, 1 = pair-list overflow, -1 (host pre-fill) = aborted
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

extern "C" int64_t tsc_elim_fused_ws_words(int64_t N) {
    const int64_t Wc = (N + 31) / 32;
    return tsc::EF_HDR + 64 + 4 * (Wc + 2) + 3 * (N > 0 ? N : 1) + 4 * (N / 16 + 2) + 16;
}

extern "C" int64_t tsc_elim_fused_out_bytes(int64_t N) { return ((N + 3) / 4) * 4 + 64 * 4; }

static int elim_fused_impl(const int32_t* lists, int32_t n_lists, int64_t stride, int64_t N, int32_t gate, int32_t* ws,
                           uint8_t* out, const int32_t* flags, int32_t epoch, void* stream);

extern "C" int tsc_elim_fused(const int32_t* lists, int32_t n_lists, int64_t stride, int64_t N, int32_t gate,
                              int32_t* ws, uint8_t* out, void* stream) {
    return elim_fused_impl(lists, n_lists, stride, N, gate, ws, out, nullptr, 0, stream);
}

// The same on lists that the ranks write into each other's memory (tsc_pairs_push): block l is read only after
// flags[l] (device memory of THIS rank, written by rank l) holds `epoch`.  Status 3 in the info words: a flag never
// arrived within ~2 s.
extern "C" int tsc_elim_fused_p2p(const int32_t* lists, int32_t n_lists, int64_t stride, int64_t N, int32_t gate,
                                  int32_t* ws, uint8_t* out, const int32_t* flags, int32_t epoch, void* stream) {
    if (!flags) return (int)cudaErrorInvalidValue;
    return elim_fused_impl(lists, n_lists, stride, N, gate, ws, out, flags, epoch, stream);
}

namespace tsc {
// rank `rank` copies its confirmed-pair list (header = count, then the pairs) into block `rank` of every peer's
// list array; peer_lists[q] = base address of rank q's array as seen from this GPU (NVLink peer mapping)
__global__ void __launch_bounds__(256) pairs_push_kernel(const int2* __restrict__ local, int64_t stride,
                                                         int2* const* __restrict__ peer_lists, int rank, int world) {
    const int64_t cnt = local[0].x;
    const int64_t n = (cnt < 0 || cnt > stride - 1) ? 0 : cnt;                 // overflowed: only the header travels
    const int64_t total = (n + 1) * world;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(t / (n + 1));
        const int64_t e = t - (int64_t)q * (n + 1);
        peer_lists[q][(int64_t)rank * stride + e] = local[e];
    }
}
// after the copies (stream order): flags of block `rank` on every peer <- epoch, released at system scope
__global__ void pairs_flag_kernel(int32_t* const* __restrict__ peer_flags, int rank, int world, int32_t epoch) {
    const int q = threadIdx.x;
    if (q < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peer_flags[q] + rank), "r"(epoch) : "memory");
    }
}
}  // namespace tsc

// All-gather of the confirmed-pair lists by peer writes, sized on the device: local_list (header + pairs, this rank's
// verify output) -> block `rank` of every rank's list array, then flags[rank] <- epoch on every rank.  peer_lists /
// peer_flags: device arrays of `world` addresses (symmetric-memory mappings, the own rank included).  Nothing is read
// back; tsc_elim_fused_p2p waits for the flags.
extern "C" int tsc_pairs_push(const int32_t* local_list, int64_t stride, void* const* peer_lists,
                              void* const* peer_flags, int32_t rank, int32_t world, int32_t epoch, void* stream) {
    using namespace tsc;
    if (world <= 0 || world > 64 || rank < 0 || rank >= world || !local_list || !peer_lists || !peer_flags)
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    pairs_push_kernel<<<148, 256, 0, st>>>(reinterpret_cast<const int2*>(local_list), stride,
                                           reinterpret_cast<int2* const*>(peer_lists), rank, world);
    TSC_CHECK_LAUNCH();
    pairs_flag_kernel<<<1, 64, 0, st>>>(reinterpret_cast<int32_t* const*>(peer_flags), rank, world, epoch);
    TSC_CHECK_LAUNCH();
    return 0;
}

static int elim_fused_impl(const int32_t* lists, int32_t n_lists, int64_t stride, int64_t N, int32_t gate, int32_t* ws,
                           uint8_t* out, const int32_t* flags, int32_t epoch, void* stream) {
    using namespace tsc;
    if (N <= 0) return 0;
    static const int64_t LADDER[18] = {500000, 200000, 100000, 50000, 20000, 10000, 5000, 2000, 1000,
                                       500, 200, 100, 50, 20, 10, 5, 2, 1};       // rmsd_pruning.py:186-188
    ElimFusedParams p;
    p.lists = reinterpret_cast<const int2*>(lists);
    p.n_lists = n_lists;
    p.stride = stride;
    p.N = N;
    p.ws = ws;
    p.out = out;
    p.n_ladder = 18;
    p.gate = gate;
    p.flags = flags;
    p.epoch = epoch;
    for (int i = 0; i < 18; i++) p.ladder[i] = LADDER[i];
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)(2 * ((N + 31) / 32 + 2) + 2 * EF_SMEM_CHUNKS) * 4;
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;      // N > ~800k: use the bit-row kernels
    cudaError_t e = cudaMemsetAsync(ws, 0, (EF_HDR + 64) * 4, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(out + ((N + 3) / 4) * 4, 0xFF, 64 * 4, st);
    if (e != cudaSuccess) return (int)e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(elim_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#ifdef TSC_EF_GRID
    sms = TSC_EF_GRID;                                           // measurement builds (tools/probes): grid-size sweep
#endif
    void* args[] = {(void*)&p};
    e = cudaLaunchCooperativeKernel((const void*)elim_fused_kernel, dim3(sms), dim3(EF_THREADS), args, smem, st);
    return (int)e;
}
