// rmsd_tf32ts.cu — tcgen05 pre-screen with the stationary operand in TENSOR MEMORY: rmsd_ts_kernel.
// FP16 operands (kind::f16, tsc_rmsd_sim_f16ts) are the default screen of prune_conformers_rmsd; TF32 operands
// (kind::tf32, tsc_rmsd_sim_tf32ts) are the same kernel with K = 8 per MMA.
//
// Same mathematics and same output contract as rmsd_tf32.cu (read its header first).  What changed,
// and why: ncu on rmsd_tf32_kernel showed the tensor pipe "busy" 68 % of the time with only 26 % of
// it doing math — every 128x48x8 MMA re-read its whole 4 KB A block from shared memory (8.7 MAC per
// operand byte; 16 are needed to stay compute-bound) and the epilogue warps sat waiting for
// accumulators.  The A panel (128 conformers) is the same for every tile of a work item, so it is
// now written ONCE per item into TMEM (tcgen05.st, one row per thread) and the MMAs take it from
// there (tcgen05.mma [d], [a_tmem], b_desc): shared-memory operand traffic drops to the 1.5 KB B
// block per MMA.
//
// TMEM map (512 columns x 128 lanes x 32 bit), a K block = 8 TF32 or 16 FP16 atoms = 32 bytes per row = 8 columns:
//   "full"   (NACC = 2, KTM = 9):  [0, 216)  A: component a, K block kb at columns a*8*KT + 8*kb .. +8, KT = min(K blocks, 9)
//                                  [216, 504) two accumulator buffers of 144 columns (D_x | D_y | D_z, each 3*16 wide)
//   "hybrid" (NACC = 3, KTM = 3):  [0, 72)   A: the first 3 K blocks;   [72, 504) three accumulator buffers
// K blocks of the panel that are not in TMEM stay in shared memory (bulk-TMA of the tail of the panel image) and are
// multiplied with the shared-memory form of the instruction (63 cycles per MMA instead of ~41-56).  launch_ts picks
// hybrid up to 5 K blocks (M <= 80 with FP16 operands) and full above.
//
// Roles (one persistent CTA per SM): warp 0 producer (A tail + ring of B tiles), warp 1 TMEM allocation + MMA
// issue, the rest epilogue in groups of 4 warps: at the start of a work item they load their rows of the A panel
// (global -> registers -> tcgen05.st), then every group takes the tiles of its accumulator buffer.  The epilogue
// (EPI = 5, tf32_common.cuh::tf32_epilogue_tile_v5) is FP32 only: Samuelson's bound, then the FP32 sign test of the
// key-matrix quartic with rigorous error bounds; what neither excludes is a candidate for the exact verify kernel.
// Template parameters select measured alternatives (see launch_ts): the FP64 second stage of round 1's first
// versions (EPI = 4), column-split groups, the panel entirely in shared memory, several MMA warps.
#include "tf32_common.cuh"

namespace tsc {

__device__ __forceinline__ long long ts_globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Two TMEM budgets (template parameter NACC):
//   NACC = 2: up to 9 K blocks of the A panel in TMEM (216 columns) + 2 accumulator buffers (288);
//   NACC = 3: the whole A panel stays in shared memory (SS-form MMAs) and TMEM holds 3 accumulator buffers
//             (432 columns), one per epilogue group: the MMA warp can run two tiles ahead of the epilogue.
//             With FP16 operands an MMA reads 4 KB (A) + 1.5 KB (B) of shared memory = 43 cycles at 128 B/clk,
//             about the ~40 cycles the tensor pipe needs for a 128x48 instruction anyway.
constexpr int TS_KT_MAX = 9;                   // K blocks held in TMEM (NACC = 2)
constexpr int TS_KT_HYB = 3;                   // K blocks held in TMEM next to THREE accumulator buffers (72 + 432 columns)
constexpr int TS_MAX_NACC = 3;
constexpr int TS_MAX_BSTAGES = 12;

// Operand images are addressed in 16-byte K chunks: 4 TF32 values (kind::tf32, K = 8 per MMA) or
// 8 FP16 values (kind::f16, K = 16 per MMA).  FP16 has the same 10-bit mantissa as TF32, so the error
// bound of rmsd_tf32.cu carries over (pack zeroes |x| < 2^-14 itself and widens sqrt(G) for it), while
// every instruction covers twice as many atoms: the MMA warp, which a clock64 trace showed to be the
// limit (35-44 cycles per 128x48 MMA whatever K is, tools/umma_probe.py), issues half as many.
struct TsParams {
    const unsigned char* PA;  // [panel][a][kc][128][16 B]    (only chunks kc >= 2*KT are read)
    const unsigned char* PB;  // [jtile][kc][48][16 B]
    const unsigned char* PR;  // [row][a][nkc][16 B]          row-major image for the TMEM part
    const double* G;
    const double* sG;
    const float* CT;          // [jtile][32]: 16 x float_rd(hs G_j), 16 x float_ru(sqrt(G_j))  (pack_tf32)
    const int4* items;        // (panel, jt_begin, jt_count, local_row_block_of_panel)
    int n_items;
    int64_t N;
    int nkc;                  // 16-byte K chunks per conformer and component
    int nb_stages;
    double e_thr;
    uint16_t* sim_bits16;
    int64_t W;
    int2* cand;               // candidate list (header + capacity entries (local row, j)), or NULL
    int64_t cand_stride;
    long long* trace;         // measurement aid (tsc_set_trace_buffer): clock64 stamps of CTA 0's first item, or NULL
};

constexpr int TS_Q = 128;              // candidate queue entries per epilogue warp (shared memory)
constexpr int TS_TRACE_TILES = 96;    // tiles of the first item that are stamped (8 slots each)

// NACC*CH epilogue groups of 4 warps (CH column parts per tile); STEP columns per TMEM load round
// NMMA MMA-issuing warps (tile t is issued by warp t % NMMA): while one of them goes through its per-tile
// bookkeeping (two mbarrier waits, descriptor set-up, commits: ~460 cycles in the trace) the other keeps the
// tensor pipe fed, and the MMAs of two tiles (six accumulator chains instead of three) interleave in the pipe.
// SPLIT: CH groups in total instead of NACC*CH — group g takes column part g of EVERY tile (both buffers in
// turn), so all epilogue warps work on the tile that has just finished and the buffer is handed back after half
// the per-warp work: the serial chain MMA -> completion -> epilogue-until-release -> MMA that bounds the tile
// rate with two buffers gets shorter, at unchanged total epilogue work.
template <int CH, int STEP, bool F16, int NACC, int NMMA = 1, bool SPLIT = false, int KTM = (NACC == 2 ? TS_KT_MAX : 0), int EPI = 4>
__global__ void __launch_bounds__((1 + NMMA + 4 * (SPLIT ? CH : NACC * CH)) * 32, 1) rmsd_ts_kernel(const TsParams p) {
    constexpr int NG = SPLIT ? CH : NACC * CH;
    constexpr int TS_NACC = NACC;
    constexpr int KT_MAX = KTM;
    static_assert(3 * 8 * KTM + NACC * TF_ACC_COLS <= TF_TMEM_COLS, "TMEM budget");
    constexpr int TS_ACC0 = 3 * 8 * KT_MAX;        // first accumulator column
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nkc = p.nkc;                                     // 16-byte K chunks
    const int nkb = nkc / 2;                                   // K blocks (one MMA each) per tile
    const int KT = nkb < KT_MAX ? nkb : KT_MAX;                // K blocks with A in TMEM
    const int tail_kc = nkc - 2 * KT;                          // chunks of A kept in shared memory
    const uint32_t tail_bytes = (uint32_t)tail_kc * TF_ROWS * 16u;      // per component
    const uint32_t b_bytes = (uint32_t)nkc * TF_N * 16u;
    unsigned char* smA = smem_raw;                             // [a][tail_kc][128][4]
    unsigned char* smB = smem_raw + 3u * tail_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)p.nb_stages * b_bytes);
    uint64_t* at_full = bars;                  // A tail landed (TMA)
    uint64_t* am_full = bars + 1;              // A rows stored to TMEM (8 epilogue warps)
    uint64_t* a_empty = bars + 2;              // all MMAs of the item retired
    uint64_t* b_full = bars + 3;
    uint64_t* b_empty = b_full + TS_MAX_BSTAGES;
    uint64_t* t_full = b_empty + TS_MAX_BSTAGES;
    uint64_t* t_empty = t_full + TS_MAX_NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + TS_MAX_NACC);
    int2* cand_q = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(bars) + 512);     // [epilogue warp][TS_Q]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(at_full, 1);
        mbar_init(am_full, 4 * NG);
        mbar_init(a_empty, NMMA);
        for (int s = 0; s < p.nb_stages; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int t = 0; t < TS_NACC; t++) { mbar_init(&t_full[t], 1); mbar_init(&t_empty[t], 4 * CH); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TF_TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer: warp-uniform control flow, one elected lane issues the copies =====================
        {
            int bs = 0; uint32_t bph = 0, aph = 0;
            for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
                const int4 w = p.items[it];
                if (w.z == 0) break;              // empty item: this CTA's list is exhausted (_host.build_tf32_items_balanced)
                if (tail_kc > 0) {
                    mbar_wait(a_empty, aph ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(at_full, 3u * tail_bytes);
                        const unsigned char* src = p.PA + (size_t)w.x * 3u * nkc * TF_ROWS * 16u;
                        for (int a = 0; a < 3; a++)
                            bulk_g2s(smA + (size_t)a * tail_bytes,
                                     src + ((size_t)a * nkc + 2 * KT) * TF_ROWS * 16u, tail_bytes, at_full);
                    }
                    __syncwarp();
                    aph ^= 1u;
                }
                for (int t = 0; t < w.z; t++) {
                    mbar_wait(&b_empty[bs], bph ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&b_full[bs], b_bytes);
                        bulk_g2s(smB + (size_t)bs * b_bytes,
                                 p.PB + (size_t)(w.y + t) * b_bytes, b_bytes,
                                 &b_full[bs]);
                    }
                    __syncwarp();
                    if (++bs == p.nb_stages) { bs = 0; bph ^= 1u; }
                }
            }
        }
    } else if (warp <= NMMA) {
        // ===================== MMA issuer: the whole warp runs the control flow, one elected lane issues =====================
        {
            const uint32_t idesc = F16 ? umma_idesc_f16(TF_ROWS, TF_N) : umma_idesc_tf32(TF_ROWS, TF_N);
            const uint32_t a_lbo = TF_ROWS * 16u, b_lbo = TF_N * 16u;
            const uint64_t bd0 = umma_desc_kmajor(smem_u32(smB), b_lbo, 128u);       // stage 0, K block 0
            const uint64_t ad0 = umma_desc_kmajor(smem_u32(smA), a_lbo, 128u);       // A tail, component x
            const uint64_t bd_step = (2u * b_lbo) >> 4, ad_step = (2u * a_lbo) >> 4; // start-address field, 16-byte units
            const uint32_t a_stride = 8u * KT;                                       // TMEM columns per component
            const int mw = warp - 1;                                                 // this MMA warp's index
            int bs = mw % p.nb_stages, acc = mw % TS_NACC; uint32_t bph = 0, aph = 0, tph = 0;
            int64_t tseq = 0;                                                        // tiles of this CTA so far
            for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
                const int4 w = p.items[it];
                if (w.z == 0) break;
                mbar_wait(am_full, aph);
                if (tail_kc > 0) mbar_wait(at_full, aph);
                aph ^= 1u;
                for (int t = 0; t < w.z; t++, tseq++) {
                    if (NMMA > 1 && (int)(tseq % NMMA) != mw) continue;
                    const bool tr = p.trace && it == 0 && t < TS_TRACE_TILES && lane == 0;
                    if (tr) p.trace[t * 8 + 0] = clock64();
                    mbar_wait(&b_full[bs], bph);
                    if (tr) p.trace[t * 8 + 1] = clock64();
                    mbar_wait(&t_empty[acc], tph ^ 1u);
                    if (tr) p.trace[t * 8 + 2] = clock64();
                    tcgen05_fence_after();
                    if (elect_one()) {
                    const uint32_t d0 = tmem_base + TS_ACC0 + (uint32_t)acc * TF_ACC_COLS;
                    // descriptors advance by one add per K block (two 16-byte chunks = 2*LBO bytes)
                    uint64_t bd = bd0 + (uint64_t)((uint32_t)bs * (b_bytes >> 4));
                    uint64_t ad = ad0;
                    if (KT_MAX > 0) {
                        uint32_t at = tmem_base;
                        // K block 0 overwrites the accumulators, the others accumulate
                        umma_tf32_ts_c<false, F16>(d0, at, bd, idesc);
                        umma_tf32_ts_c<false, F16>(d0 + TF_N, at + a_stride, bd, idesc);
                        umma_tf32_ts_c<false, F16>(d0 + 2 * TF_N, at + 2 * a_stride, bd, idesc);
#pragma unroll 4
                        for (int kb = 1; kb < KT; kb++) {
                            bd += bd_step;
                            at += 8;
                            umma_tf32_ts_c<true, F16>(d0, at, bd, idesc);
                            umma_tf32_ts_c<true, F16>(d0 + TF_N, at + a_stride, bd, idesc);
                            umma_tf32_ts_c<true, F16>(d0 + 2 * TF_N, at + 2 * a_stride, bd, idesc);
                        }
                    } else {                                      // whole panel in shared memory
                        umma_tf32_ss_c<false, F16>(d0, ad, bd, idesc);
                        umma_tf32_ss_c<false, F16>(d0 + TF_N, ad + (tail_bytes >> 4), bd, idesc);
                        umma_tf32_ss_c<false, F16>(d0 + 2 * TF_N, ad + 2 * (tail_bytes >> 4), bd, idesc);
                        ad += ad_step;
                    }
#pragma unroll 4
                    for (int kb = (KT_MAX > 0 ? KT : 1); kb < nkb; kb++) {       // K blocks whose A block is in shared memory
                        bd += bd_step;
                        umma_tf32_ss_c<true, F16>(d0, ad, bd, idesc);
                        umma_tf32_ss_c<true, F16>(d0 + TF_N, ad + (tail_bytes >> 4), bd, idesc);
                        umma_tf32_ss_c<true, F16>(d0 + 2 * TF_N, ad + 2 * (tail_bytes >> 4), bd, idesc);
                        ad += ad_step;
                    }
                    umma_commit(&b_empty[bs]);
                    umma_commit(&t_full[acc]);
                    }
                    __syncwarp();
                    if (tr) { p.trace[t * 8 + 3] = clock64(); p.trace[t * 8 + 7] = ts_globaltimer_ns(); }
                    bs += NMMA;
                    if (bs >= p.nb_stages) { bs -= p.nb_stages; bph ^= 1u; }
                    acc += NMMA;
                    if (acc >= TS_NACC) { acc -= TS_NACC; tph ^= 1u; }
                }
                if (elect_one()) umma_commit(a_empty);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: 2*CH groups of 4 warps =====================
        // group g works on accumulator buffer g / CH (the tiles with index % 2 == g / CH) and, inside a
        // tile, on the 16/CH columns of part g % CH.  Every group therefore waits on EVERY use of its
        // buffer's barrier, which mbarrier parity waits require (a waiter may never fall two phases behind).
        const int ew = warp - (1 + NMMA);
        const int grp = ew >> 2;
        const int part = grp % CH;
        int buf = SPLIT ? 0 : grp / CH;
        uint32_t tph_b[TS_MAX_NACC] = {0, 0, 0};                 // SPLIT: one phase per buffer
        const int quad = warp & 3;
        const int row_in_panel = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        uint32_t eph = 0, tph = 0;
        uint32_t tile_seq = 0;                          // tiles of this CTA before the current item
        // Candidates (rare: ~1 per 5000 pairs) go to a per-warp queue in shared memory and reach the global
        // list in bursts with ONE atomic per burst: a global atomic per candidate (~1 us round trip inside
        // the tile loop) had cost 0.35 ms on C3.
        int2* my_q = cand_q + (size_t)ew * TS_Q;
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
            const int64_t i = (int64_t)w.x * TF_ROWS + row_in_panel;
            // ---- this thread's row of the A panel -> TMEM; the row's 3*2*KT float4 are dealt round-robin to the groups
            mbar_wait(a_empty, eph ^ 1u);
            eph ^= 1u;
            tcgen05_fence_after();
            {
                const float4* src = reinterpret_cast<const float4*>(p.PR + (size_t)i * 3 * nkc * 16);
                for (int q = grp; q < 3 * 2 * KT; q += NG) {
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
            uint8_t* out_row = reinterpret_cast<uint8_t*>(p.sim_bits16) + ((int64_t)w.w * CB + row_in_panel) * (4 * p.W);
            float gv_next = 0.f;
            bool have_next = false;
            // this group's tiles of the item: every tile (SPLIT) or those whose running index is = buf (mod NACC);
            // stepping the loop by NACC instead of testing every tile removed a 64-bit modulo + branch per tile
            // that ncu showed at 8 % of the kernel's stall samples
            constexpr int AHEAD = SPLIT ? 1 : NACC;                      // distance to this group's next tile
            const int t_first = SPLIT ? 0 : (int)((uint32_t)(buf + NACC - (int)(tile_seq % NACC)) % NACC);
            for (int t = t_first; t < w.z; t += AHEAD) {
                if (SPLIT) buf = (int)((tile_seq + (uint32_t)t) % NACC);
                {
                    const int64_t j0 = (int64_t)(w.y + t) * TF_J;
                    // column terms: prefetched one of this group's tiles ahead (the first of an item is a direct load)
                    const float gvf = have_next ? gv_next : __ldg(&p.CT[(int64_t)(w.y + t) * 32 + lane]);
                    have_next = t + AHEAD < w.z;
                    if (have_next) gv_next = __ldg(&p.CT[(int64_t)(w.y + t + AHEAD) * 32 + lane]);
                    const bool tr = p.trace && it == 0 && t < TS_TRACE_TILES && lane == 0 && quad == 0 && part == 0;
                    if (tr) p.trace[t * 8 + 4] = clock64();
                    if (SPLIT) {
                        // (static indexing keeps the phases in registers)
                        const uint32_t ph = buf == 0 ? tph_b[0] : buf == 1 ? tph_b[1] : tph_b[2];
                        mbar_wait(&t_full[buf], ph);
                        if (buf == 0) tph_b[0] ^= 1u; else if (buf == 1) tph_b[1] ^= 1u; else tph_b[2] ^= 1u;
                    } else {
                        mbar_wait(&t_full[buf], tph);
                        tph ^= 1u;
                    }
                    if (tr) p.trace[t * 8 + 5] = clock64();
                    tcgen05_fence_after();
                    const uint32_t d0 = tmem_base + lane_addr + TS_ACC0 + (uint32_t)buf * TF_ACC_COLS;
                    uint32_t bits;
                    if (EPI == 5)
                        bits = tf32_epilogue_tile_v5<TF_J / CH>(d0, gvf, row, i, j0, p.N, lane, &t_empty[buf],
                                                                part * (TF_J / CH));
                    else if (STEP == 4)
                        bits = tf32_epilogue_tile_v4<TF_J / CH>(d0, gvf, row, p.G, p.sG, i, j0, p.N, lane, &t_empty[buf],
                                                                part * (TF_J / CH));
                    else
                        bits = tf32_epilogue_tile<STEP, TF_J / CH>(d0, gvf, row, p.G, p.sG, i, j0, p.N, lane,
                                                                   &t_empty[buf], part * (TF_J / CH));
                    if (tr) p.trace[t * 8 + 6] = clock64();
                    if (p.cand) {
                        uint32_t bb = (i < p.N) ? (CH == 1 ? bits : (bits & (0xffu << (8 * part)))) : 0u;
                        if (__any_sync(0xffffffffu, bb != 0u)) {      // append (local row, j) of every bit set
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
                            if (qn + total > TS_Q) flush_q();
                            if (total > TS_Q) {                       // a dense tile: straight to the global list
                                int64_t base = 0;
                                if (lane == 0) base = list_reserve(&p.cand[0].x, total, p.cand_stride - 1);
                                base = __shfl_sync(0xffffffffu, base, 0);
                                while (bb) {
                                    const int b = __ffs(bb) - 1;
                                    bb &= bb - 1;
                                    if (base + at < p.cand_stride - 1)
                                        p.cand[1 + base + at] = make_int2(lrow, (int32_t)(j0 + b));
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
                    }
                    if (i < p.N && (j0 >> 4) < 2 * p.W) {
                        if (CH == 1) *reinterpret_cast<uint16_t*>(out_row + (j0 >> 3)) = (uint16_t)bits;
                        else out_row[(j0 >> 3) + part] = (uint8_t)(bits >> (8 * part));
                    }
                }
            }
            tile_seq += (uint32_t)w.z;
        }
        if (p.cand && qn) flush_q();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TF_TMEM_COLS);
}

}  // namespace tsc

extern "C" int64_t tsc_tf32_pr_floats(int64_t N, int32_t M) {
    const int64_t Mp = (M + 7) / 8 * 8, rows = (N + tsc::TF_ROWS - 1) / tsc::TF_ROWS * tsc::TF_ROWS;
    return rows * 3 * Mp;
}

static long long* g_ts_trace = nullptr;
// Measurement aid: a device buffer of 8 * 96 int64 that receives clock64 stamps of the first work item
// of CTA 0 (MMA warp: start, operands landed, accumulator free, issued; epilogue: start, accumulator
// full, tile done).  NULL (default) disables it.
extern "C" void tsc_set_trace_buffer(void* dev_ptr) { g_ts_trace = reinterpret_cast<long long*>(dev_ptr); }

extern "C" int64_t tsc_tf32_ct_floats(int64_t N) {
    return (N + tsc::TF_ROWS - 1) / tsc::TF_ROWS * tsc::TF_ROWS * 2;
}

namespace tsc {
template <bool F16>
static int launch_ts(const void* PA, const void* PB, const void* PR, const double* G, const double* sG, const float* CT,
                     int64_t N, int32_t M, const int32_t* items, int32_t n_items, double thr, uint32_t* sim_bits,
                     int32_t* cand_list, int64_t cand_stride, int32_t grid_ctas, void* stream) {
    if (n_items <= 0 || N <= 0) return 0;
    TsParams p;
    p.cand = reinterpret_cast<int2*>(cand_list);
    p.cand_stride = cand_stride;
    p.PA = reinterpret_cast<const unsigned char*>(PA);
    p.PB = reinterpret_cast<const unsigned char*>(PB);
    p.PR = reinterpret_cast<const unsigned char*>(PR);
    p.G = G; p.sG = sG; p.CT = CT;
    p.trace = g_ts_trace;
    p.items = reinterpret_cast<const int4*>(items);
    p.n_items = n_items;
    p.N = N;
    p.nkc = F16 ? (M + 15) / 16 * 2 : (M + 7) / 8 * 2;         // atoms padded to whole K blocks
    p.e_thr = (double)M * thr * thr * (1.0 + 1e-6);
    p.sim_bits16 = reinterpret_cast<uint16_t*>(sim_bits);
    p.W = num_blocks_padded(N);
    // grid_ctas < 0 selects an alternative configuration (tuning aid): -1 = A in TMEM, 2 groups, 8 columns per
    // TMEM load round; -2 = A in TMEM, 2 groups x 4; -3 = A in TMEM, 4 groups (two column halves per tile) x 4;
    // -4 = A in shared memory, 3 accumulator buffers, 3 groups x 4;  -5 = as -2 with two MMA warps;
    // -7 = A in TMEM, 2 groups each taking one column half of EVERY tile (SPLIT);
    // -8 = hybrid: the first 3 K blocks of A in TMEM (72 columns), the rest in shared memory, THREE accumulator
    //      buffers and three epilogue groups;  -9 = as -8 with three MMA warps, one per buffer
    //      (a warp must see every phase of the barriers it waits on: two warps on three buffers alias phases and hang);
    // -10 = as -8 with the two-stage FP32 epilogue (tf32_epilogue_tile_v5);  -11 = as -2 with that epilogue;
    // (measured and removed: issuing the second half of tile t's K blocks alternately with the first half of tile
    //  t+1's — six accumulator chains in flight instead of three, 45 instead of 56 cycles per MMA in
    //  tools/umma_probe.py — needs the NEXT tile's accumulator buffer before the current tile is committed, which
    //  leaves the epilogue less than one tile time to hand a buffer back: C3 3.26 ms with three buffers, 3.78 with
    //  two, against 2.31 / 2.32 without)
    // (measured on C3, FP16 operands: -2 2.72 ms, -3 3.08, -4 2.78, -5 2.88, -7 3.60 — an epilogue warp spends
    // ~600 cycles per tile whatever the number of columns it handles, so fewer columns per warp-tile lose;
    // reading the whole tile into 144
    // registers and releasing before any arithmetic: 4.06)
    // default: two-stage FP32 epilogue; three accumulator buffers next to a 3-K-block TMEM panel while the part of the
    // panel left in shared memory is small (<= 5 K blocks in all: M <= 80 with FP16 operands), otherwise the whole
    // panel (up to 9 K blocks) in TMEM and two buffers (20 000 x 150 atoms: 0.64 against 0.73 ms; equal at 80 atoms)
    const int cfg = grid_ctas < 0 ? -grid_ctas : (p.nkc / 2 <= 5 ? 10 : 11);
    if (grid_ctas < 0) grid_ctas = 0;
    const int kt_max = cfg == 4 ? 0 : (cfg == 8 || cfg == 9 || cfg == 10) ? TS_KT_HYB : TS_KT_MAX;
    const int nkb = p.nkc / 2, KT = nkb < kt_max ? nkb : kt_max;
    const size_t a_bytes = (size_t)3 * (p.nkc - 2 * KT) * TF_ROWS * 16, b_bytes = (size_t)p.nkc * TF_N * 16;
    const size_t q_bytes = (size_t)16 * TS_Q * sizeof(int2);            // candidate queues of up to 16 epilogue warps
    const size_t budget = 227 * 1024 - 512 - q_bytes;
    if (a_bytes + 2 * b_bytes > budget) return (int)cudaErrorInvalidValue;
    int nb = (int)((budget - a_bytes) / b_bytes);
    if (nb > TS_MAX_BSTAGES) nb = TS_MAX_BSTAGES;
    p.nb_stages = nb;
    const size_t smem = a_bytes + nb * b_bytes + 512 + q_bytes;
    auto kern = cfg == 1 ? rmsd_ts_kernel<1, 8, F16, 2> : cfg == 2 ? rmsd_ts_kernel<1, 4, F16, 2>
              : cfg == 3 ? rmsd_ts_kernel<2, 4, F16, 2> : cfg == 4 ? rmsd_ts_kernel<1, 4, F16, 3>
              : cfg == 5 ? rmsd_ts_kernel<1, 4, F16, 2, 2> : cfg == 8 ? rmsd_ts_kernel<1, 4, F16, 3, 1, false, TS_KT_HYB>
              : cfg == 9 ? rmsd_ts_kernel<1, 4, F16, 3, 3, false, TS_KT_HYB>
              : cfg == 10 ? rmsd_ts_kernel<1, 4, F16, 3, 1, false, TS_KT_HYB, 5>
              : cfg == 11 ? rmsd_ts_kernel<1, 4, F16, 2, 1, false, TS_KT_MAX, 5> : rmsd_ts_kernel<2, 4, F16, 2, 1, true>;
    const int threads = cfg <= 2 || cfg == 11 ? 320 : cfg == 3 ? 576 : cfg == 4 || cfg == 8 || cfg == 10 ? 448
                      : cfg == 5 ? 352 : cfg == 9 ? 512 : 320;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = grid_ctas > 0 ? grid_ctas : sms;
    if (grid > n_items) grid = n_items;
    kern<<<grid, threads, smem, (cudaStream_t)stream>>>(p);
    TSC_CHECK_LAUNCH();
    return 0;
}
}  // namespace tsc

extern "C" int tsc_rmsd_sim_tf32ts(const float* PA, const float* PB, const float* PR, const double* G,
                                   const double* sG, const float* CT, int64_t N, int32_t M, const int32_t* items,
                                   int32_t n_items, double thr, uint32_t* sim_bits, int32_t* cand_list,
                                   int64_t cand_stride, int32_t grid_ctas, void* stream) {
    return tsc::launch_ts<false>(PA, PB, PR, G, sG, CT, N, M, items, n_items, thr, sim_bits, cand_list, cand_stride,
                                 grid_ctas, stream);
}

// FP16-operand form (default): PA / PB / PR are the images tsc_pack_f16 writes.
extern "C" int tsc_rmsd_sim_f16ts(const void* PA, const void* PB, const void* PR, const double* G,
                                  const double* sG, const float* CT, int64_t N, int32_t M, const int32_t* items,
                                  int32_t n_items, double thr, uint32_t* sim_bits, int32_t* cand_list,
                                  int64_t cand_stride, int32_t grid_ctas, void* stream) {
    return tsc::launch_ts<true>(PA, PB, PR, G, sG, CT, N, M, items, n_items, thr, sim_bits, cand_list, cand_stride,
                                grid_ctas, stream);
}
