/*
 * tscode_b200.h — C-ABI of libtscode_b200.so: the B200-native conformer-ensemble hot path of
 * TSCoDe (RMSD pruning, clash screen, rigid-body pose transforms).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless marked [host]; arrays are row-major, FP64
 *     unless typed otherwise; `stream` is a cudaStream_t passed as void*;
 *   - every compute entry point is asynchronous on `stream` and returns 0 or a cudaError_t
 *     (tsc_error_string() decodes it); nothing allocates, nothing synchronises;
 *   - no torch / Python types cross this boundary.  The Python package binds it with ctypes
 *     (tscode_b200/_lib.py); INTEGRATION.md shows the binding a TSCoDe maintainer would add.
 *
 * Reference interfaces replaced (paths relative to the TSCoDe tree, tscode/...):
 *   rmsd_pruning.py:164   prune_conformers_rmsd   -> tsc_pack + tsc_rmsd_sim_tiles +
 *                                                    tsc_rmsd_verify + tsc_elim_*
 *   rmsd_pruning.py:6     rmsd_and_max_numba      -> tsc_rmsd_pairs
 *   rmsd_pruning.py:208   _rmsd_similarity        -> tsc_rmsd_pairs (broadcast_p = 1)
 *   numba_functions.py:60 compenetration_check    -> tsc_clash_structs / tsc_embed_clash
 *   embeds.py:961         get_embed               -> tsc_embed_gather (and fused in tsc_embed_clash)
 *   torsion_module.py:953 rotationally_corrected_rmsd / :1013 prune_conformers_rmsd_rot_corr
 *                                                 -> tsc_rotcorr_pairs + tsc_rotcorr_apply
 */
#ifndef TSCODE_B200_H
#define TSCODE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- bookkeeping ------------------------------------------------------------------------ */
int tsc_version(void);                         /* 100 = 0.1.0 */
const char* tsc_error_string(int code);        /* [host] */
int32_t tsc_device_sm_count(void);

/* Packed-ensemble geometry (see tscode_b200/csrc/tsc_common.cuh):
 *   conformer blocks of 32 (count rounded up to even = nb_pad), atom slabs of 20;
 *   packed[slab][block][xyz][32][20] doubles; sim rows have W = nb_pad 32-bit words. */
int64_t tsc_num_blocks_padded(int64_t N);
int32_t tsc_num_slabs(int32_t M);
int64_t tsc_packed_doubles(int64_t N, int32_t M);

/* ---- prune_conformers_rmsd ---------------------------------------------------------------- */
/* Heavy-atom gather (rmsd_pruning.py:178-179) + repack + squared norms.
 *   S (N, A, 3); heavy_idx (M) int32 = indices with atomnos != 1;
 *   packed: tsc_packed_doubles(N, M) doubles; G: nb_pad*32 doubles (sum |p|^2 per conformer). */
int tsc_pack(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M,
             double* packed, double* G, void* stream);

/* tsc_pack for the conformer blocks (32 rows) [block_begin, block_end) only — chunked upload pipelines. */
int tsc_pack_blocks(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M,
                    double* packed, double* G, int64_t block_begin, int64_t block_end, void* stream);

/* All-pairs similarity screen over a list of 32 x 64 pair tiles.
 *   tiles (n_tiles, 4) int32: {ib, jp, lb, 0} = I block ib, J blocks 2jp and 2jp+1, rows written
 *   at local row block lb of sim_bits.  For every owned row block ib the caller lists jp from
 *   ib/2 to nb_pad/2 - 1, so that all words >= ib of the row are written.
 *   sim_bits (n_local_blocks*32, W) uint32: bit j of row i set iff pair (i, j>i) may have
 *   rmsd < thr (screen; tsc_rmsd_verify makes the bits exact).
 *   variant 0 = FP64 tensor cores (DMMA), 1 = FP64 FMA pipe.  grid_ctas 0 = one CTA per SM. */
int tsc_rmsd_sim_tiles(const double* packed, const double* G, int64_t N, int32_t M,
                       const int32_t* tiles, int64_t n_tiles, double thr, uint32_t* sim_bits,
                       int32_t variant, int32_t grid_ctas, void* stream);

/* DEFAULT all-pairs pre-screen (rmsd_screen.cu): tcgen05 / TMEM, FP16 operands, FP32 accumulation, one
 * accumulator buffer per (tile, row of the covariances), three MMA chains, epilogue on T = S^T S.
 * Replaces the pair loop of rmsd_pruning.py:43-79 together with tsc_rmsd_verify: bits are a superset of the
 * similar pairs, every set bit is also appended to cand_list.
 *   mode 0: isotropic form — tiles of 48 conformers (MMAs 128 x 144 x 16, three buffers = one per row of the
 *           covariances), Samuelson's bound only;
 *   mode 1: tiles of 32 (MMAs 128 x 96 x 16, four buffers), Samuelson then the FP32 quartic sign test;
 *   mode 2: tiles of 32, quartic sign test for every pair (anisotropic ensembles);
 *   mode 3: mode 0 on tiles of 64 (MMAs 128 x 192 x 16, two buffers) — the previous default, kept for comparison.
 *   All are conservative (never lose a similar pair); they differ in speed and in how many candidates they leave.
 *   PA, PB, PR: tsc_screen_operand_bytes(N, M) bytes each; CT: tsc_screen_ct_floats(N) floats; G, sG: doubles for
 *   every padded row (tsc_screen_rows_padded(N): whole 128-row panels and whole tiles).  tsc_pack_screen writes them
 *   for conformers [row_begin, row_end) (row_begin a multiple of 8; row_end <= 0 = to the end incl. padding) for
 *   tiles of tile_j = 48 (mode 0), 32 (modes 1, 2) or 64 (mode 3).
 *   items (n_items, 4) int32 {panel, first j tile, tile count, local 32-row block of the panel's first row in
 *   sim_bits}, dealt round-robin to the CTAs; an item with count 0 ends a CTA's list.
 *   M <= tsc_screen_max_atoms(tile_j); above, use tsc_rmsd_sim_tiles.  grid_ctas 0 = one CTA per SM.  pace 0. */
int64_t tsc_screen_rows_padded(int64_t N);
int64_t tsc_screen_operand_bytes(int64_t N, int32_t M);
int64_t tsc_screen_ct_floats(int64_t N);
int32_t tsc_screen_max_atoms(int32_t tile_j);
int tsc_pack_screen(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M, void* PA, void* PB,
                    void* PR, double* G, double* sG, float* CT, int64_t row_begin, int64_t row_end, int32_t tile_j,
                    const double* frame, void* stream);
int tsc_rmsd_screen(const void* PA, const void* PB, const void* PR, const double* G, const double* sG,
                    const float* CT, int64_t N, int32_t M, const int32_t* items, int32_t n_items, double thr,
                    uint32_t* sim_bits, int32_t* cand_list, int64_t cand_stride, int32_t grid_ctas, int32_t mode,
                    int32_t pace, const double* frame, void* stream);
/*   frame (HOST memory, 12 doubles, or NULL): an orthogonal matrix Q (row-major) applied to every conformer and
 *   three scales t applied to the column-side image along Q's axes.  With sum_b 1 / (3 t_b^2) <= 1 the Frobenius norm
 *   of the scaled covariance still bounds lambda_max (Cauchy-Schwarz over the column norms), and in the principal-axes
 *   frame of the molecule with t_b^2 = (l1 + l2 + l3) / (3 l_b) that bound is as sharp for elongated and planar
 *   molecules as Samuelson's is for isotropic ones (rmsd_screen.cu, ScFrame).  NULL = identity = Samuelson.  The same
 *   frame must be given to tsc_pack_screen and tsc_rmsd_screen; both return cudaErrorInvalidValue (1) when Q is not
 *   orthogonal to 1e-13 or the scales violate the inequality — the conditions exclusion rests on. */
/* The reference's survivor choice inside one chunk of a grouping loop (torsion_module.py:1136-1152,
 * numba_functions.py:203-220, optimization_methods.py:341-355: `nx.Graph(matches)`, connected components, keep
 * group[0]) in native host code, with CPython's set / dict iteration orders restated exactly — which member is
 * group[0] depends on them (capi.cu).  mi, mj (n): the matches (chunk-relative, i < j, insertion order);
 * n_nodes_max: chunk length; rejects: the members not kept.  Returns their number (-1: bad arguments). */
int64_t tsc_host_cluster_rejects(const int32_t* mi, const int32_t* mj, int64_t n, int64_t n_nodes_max, int32_t* rejects);
/* The whole grouping loop on the host (torsion_module.py:1076-1152 and its TFD / MOI siblings): ladder of k values, gate
 * `k == 1 or gate * k < active`, k chunks per round (the last one ending at the ACTIVE count: the reference's quirk), per
 * chunk tsc_host_rotcorr_chunk + tsc_host_cluster_rejects.  final_mask: N bytes, all 1 on entry; scratch: 3 N int32;
 * T = 0: no rotor states (state / compact / off / ang_table unused).  Returns the number of chunks with matches. */
int64_t tsc_host_ladder_replay(int64_t N, const int64_t* ladder, int32_t n_ladder, int32_t gate,
                               const int64_t* first_hit, int64_t* reach, double* state, int32_t T,
                               const uint64_t* compact, const int64_t* off, const double* ang_table,
                               uint8_t* final_mask, int32_t* scratch);

/* Host-side plan of the screen (capi.cu; no GPU involved): tsc_host_sample_pairs fills K fixed pseudo-random pairs
 * i != j of [0, N); tsc_host_screen_plan takes the frame from structure `first` of the HOST array S (rows, A, 3) and
 * reports, for the pairs (pi[k], pj[k]) (row numbers of S), the fraction the weighted bound would leave undecided, and
 * the looseness `ratio` of the unweighted bound for that shape (1 = isotropic).  _host.screen_plan turns the two
 * numbers into the mode given to tsc_rmsd_screen.  Speed decisions only. */
void tsc_host_sample_pairs(int64_t N, int32_t K, int64_t* pi, int64_t* pj);
int tsc_host_screen_plan(const double* S, int32_t A, const int32_t* heavy_idx, int32_t M, int64_t first,
                         const int64_t* pi, const int64_t* pj, int32_t K, double thr, double* frame12, double* ratio,
                         double* undecided);
/*   cand_list (may be NULL): block of cand_stride int32 pairs, element 0 = header whose first int32 is the running
 *   count (zero it first); (local row, j) of every bit the screen sets is appended from element 1; a count outside
 *   [0, cand_stride - 1] means the list overflowed (tsc_rmsd_verify then scans the bit rows instead). */

/* Exact re-evaluation of every set bit, the way rmsd_and_max_numba does it (rmsd_pruning.py:6-41);
 * afterwards bit (i,j) == (rmsd < thr and maxdev < 2*thr)  (:75, :95).
 *   row_blocks (n_rb) int32: global block index of each local row block.
 *   stats (4) uint64, accumulated: candidates, confirmed, within 1e-6 A of a threshold,
 *   degenerate optimal rotation.
 *   pair_list (may be NULL): a block of `pair_stride` int32 pairs; element 0 is a header whose
 *   first int32 is the running pair count (zero it before the first call), confirmed pairs (i, j)
 *   are appended from element 1 (unordered).  A count > pair_stride - 1 means the list overflowed
 *   (pairs beyond the capacity are dropped; the bit rows stay complete).
 *   cand_list (may be NULL): the candidate list a tcgen05 screen wrote.  If its count is within the
 *   capacity the candidates are verified straight from the list (32 per warp), else — or with NULL, or a
 *   negative count — every set bit of the owned rows is found by scanning the bit rows. */
int tsc_rmsd_verify(const double* packed, int64_t N, int32_t M, const int32_t* row_blocks,
                    int32_t n_rb, double thr, uint32_t* sim_bits, uint64_t* stats, int32_t* pair_list,
                    int64_t pair_stride, const int32_t* cand_list, int64_t cand_stride, void* stream);
/* Incremental form: verifies only the candidate-list entries appended since the previous call and advances
 * progress[0] (device int32; zero it together with the list header).  For callers that interleave screen launches and
 * verification on one stream, so that only the last launch's candidates are left when the screen ends.  final_call != 0
 * also runs the bit-row scan that takes over when the list has overflowed (once, at the end). */
int tsc_rmsd_verify_incr(const double* packed, int64_t N, int32_t M, const int32_t* row_blocks, int32_t n_rb,
                         double thr, uint32_t* sim_bits, uint64_t* stats, int32_t* pair_list, int64_t pair_stride,
                         const int32_t* cand_list, int64_t cand_stride, int32_t* progress, int32_t final_call,
                         void* stream);

/* Batched rmsd_and_max_numba on explicit pairs: P, Q (n, M, 3) -> rmsd (n), maxdev (n).
 * broadcast_p = 1 compares P[0] with every Q[k] (_rmsd_similarity, rmsd_pruning.py:208-224). */
int tsc_rmsd_pairs(const double* P, const double* Q, int64_t n, int32_t M, int32_t broadcast_p,
                   double* rmsd, double* maxdev, void* stream);

/* Group-local de-duplication inside the cyclical embeds (embeds.py:714-718, 842-846): a pose that passed the clash
 * test is kept iff _rmsd_similarity(pose, kept poses of its group, rmsd_thr) is False (rmsd_pruning.py:208-224,
 * all atoms).  tsc_rmsd_pairs_idx evaluates the similarity of explicit index pairs (pi[k] = later pose = `ref`,
 * pj[k] = earlier pose) of one pose array S (P, M, 3); tsc_group_greedy replays the sequential keep/append logic,
 * one thread per group: members [g_begin[g], g_begin[g+1]) of `order` (pose indices in generation order), the
 * pairs of member m stored at sim[pair_base[m] + r], r = rank of the earlier member inside the group.
 * keep (P) uint8 must be zeroed first (non-members stay 0). */
int tsc_rmsd_pairs_idx(const double* S, const int32_t* pi, const int32_t* pj, int64_t n, int32_t M, double thr,
                       uint8_t* sim, void* stream);
int tsc_group_greedy(const int32_t* g_begin, int32_t n_groups, const int32_t* order, const int64_t* pair_base,
                     const uint8_t* sim, uint8_t* keep, void* stream);

/* One ladder round (rmsd_pruning.py:123-162) in three steps; cs = int(N // k) (:136).
 *   gate (device, 1 int32, or NULL): number of active structures BEFORE this round.  When given,
 *   every kernel first evaluates the reference's gate `k == 1 or 20*k < active` (:192) and does
 *   nothing if it is closed — the host can enqueue every candidate round without a readback.
 *   cachebits ((N+31)/32 words): bit s set iff key (first, s) is cached and `first` starts a
 *   chunk of this round;   key_first/key_second (N) int32 + n_keys (1) int32: the cache (:183,:204). */
int tsc_elim_cachebits(const int32_t* key_first, const int32_t* key_second, const int32_t* n_keys,
                       int64_t N, int64_t cs, int64_t k, uint32_t* cachebits, const int32_t* gate,
                       void* stream);
/*   row_state (N) int32, written for owned rows only: -2 inactive, -1 kept, >= 0 dropped (value =
 *   second element of the emitted cache key, first + j - i). */
int tsc_elim_round(const uint32_t* sim_bits, const int32_t* row_blocks, int32_t n_rb,
                   const uint32_t* active_words, const uint32_t* cachebits, int64_t N, int64_t cs,
                   int64_t k, int32_t* row_state, const int32_t* gate, void* stream);
/*   row_state -> active words, byte mask, emitted keys appended to the cache; n_active_out (must be
 *   zero on entry) receives the new active count, or a copy of *gate when the round was skipped. */
int tsc_elim_commit(const int32_t* row_state, int64_t N, int64_t cs, int64_t k,
                    uint32_t* active_words_out, uint8_t* mask_out, int32_t* key_first,
                    int32_t* key_second, int32_t* n_keys, const int32_t* gate, int32_t* n_active_out,
                    void* stream);

/* The whole k-ladder in ONE persistent cooperative launch on the pair lists tsc_rmsd_verify emits
 * (rmsd_pruning.py:186-204; same masks as the tsc_elim_cachebits/round/commit path).
 *   lists: n_lists blocks of `stride` int32 pairs laid out as tsc_rmsd_verify writes them (one block
 *   per rank after an all-gather; every rank runs the kernel on the complete set);
 *   gate: 20 (rmsd_pruning.py:192);  ws: tsc_elim_fused_ws_words(N) int32 of scratch;
 *   out: tsc_elim_fused_out_bytes(N) bytes = N mask bytes, padded to a multiple of 4, then 64 int32:
 *   [0] status (0 done, 1 a pair list overflowed -> use the bit-row kernels, -1 aborted),
 *   [1] rounds run, [2] survivors, [3] cache keys emitted, [8..8+rounds) the k of every round run.
 *   Returns cudaErrorInvalidValue for N > ~800 000 (bitmaps are staged in shared memory). */
int64_t tsc_elim_fused_ws_words(int64_t N);
int64_t tsc_elim_fused_out_bytes(int64_t N);
int tsc_elim_fused(const int32_t* lists, int32_t n_lists, int64_t stride, int64_t N, int32_t gate,
                   int32_t* ws, uint8_t* out, void* stream);
/* Several ranks without a collective: every rank copies its confirmed-pair list into block `rank` of every rank's list
 * array through NVLink peer mappings (symmetric memory: peer_lists[q] / peer_flags[q] = rank q's array / flag words as
 * mapped on this GPU, own rank included; device arrays of `world` addresses), sized on the device from the list's own
 * header, then raises flags[rank] = epoch on every rank (system-scope release).  tsc_elim_fused_p2p is tsc_elim_fused
 * on such an array: block l is read after flags[l] == epoch (status word 3 if a flag does not arrive within ~2 s).
 * Use two arrays alternately (a rank may be one call ahead of its peers). */
int tsc_pairs_push(const int32_t* local_list, int64_t stride, void* const* peer_lists, void* const* peer_flags,
                   int32_t rank, int32_t world, int32_t epoch, void* stream);
int tsc_elim_fused_p2p(const int32_t* lists, int32_t n_lists, int64_t stride, int64_t N, int32_t gate, int32_t* ws,
                       uint8_t* out, const int32_t* flags, int32_t epoch, void* stream);

/* ---- prune_conformers_tfd / prune_by_moment_of_inertia (run right before the RMSD prune, embedder.py:1325-1352) */
/* Torsion fingerprints (numba_functions.py:233-239, 258-268; dihedral of algebra.py:24-57): tf (N, Q) float32
 * degrees for quadruplets quads (Q, 4) int32 of the structures S (N, A, 3). */
int tsc_tfd_fingerprints(const double* S, int64_t N, int32_t A, const int32_t* quads, int32_t Q, float* tf,
                         void* stream);
/* first_hit[i] = first j > i with tfd_similarity(tf[i], tf[j], thresh) (numba_functions.py:241-256), N if none:
 * all the grouping loop (numba_functions.py:155-231) needs.  near_count (may be NULL): pairs with
 * |sum - thresh| < 1e-4 among those looked at. */
int tsc_tfd_scan(const float* tf, int64_t N, int32_t Q, double thresh, int32_t* first_hit, uint64_t* near_count,
                 void* stream);
/* Principal moments of inertia of the heavy atoms (algebra.py:166-187): moments (N, 3), ascending. */
int tsc_moi_moments(const double* S, int64_t N, int32_t A, const int32_t* heavy_idx, int32_t M,
                    const double* masses, double* moments, void* stream);
/* first_hit[i] = first j > i with all(|I_i - I_j| / I_i < max_deviation) (algebra.py:189-203), N if none. */
int tsc_moi_scan(const double* moments, int64_t N, double max_deviation, int32_t* first_hit, uint64_t* near_count,
                 void* stream);

/* Distance-constraint scores of P structures: score_abs32 = _score_embed_poses (numba_functions.py:273-288, sum of
 * |dist - target| accumulated in float32), error_signed = the error of fitness_check (optimization_methods.py:544-557,
 * signed sum in float64; the verdict is error < threshold).  cons (K, 2) int32 / targets (K) shared by all structures
 * (per_pose = 0) or (P, K, 2) / (P, K) (per_pose = 1); NaN target = no target (None).  Either output may be NULL. */
int tsc_constraint_scores(const double* S, int64_t P, int32_t A, const int32_t* cons, const double* targets,
                          int32_t K, int32_t per_pose, float* score_abs32, double* error_signed, void* stream);

/* ---- compenetration_check / get_embed -------------------------------------------------- */
/* Fused pose transform + clash screen (embeds.py:116-118 / 713-714 / 841-842).
 *   frag_lib: all fragments' conformers back to back; frag_off (F) int64 = offset in doubles of
 *   fragment k; conformer c of fragment k at frag_off[k] + c*3*n_atoms[k].  conf (P, F) int32,
 *   R (P, F, 3, 3), t (P, F, 3).  F in {2, 3}.  t2: d < thresh <=> d*d < t2 (host computes the
 *   exact image of thresh under correctly-rounded sqrt).  verdict (P) uint8 = 1 pass / 0 clash.
 *   near_count (1) uint64 or NULL: if given, early exit is disabled and pairs with
 *   |d - thresh| < 1e-9 are counted (parity reporting). */
int tsc_embed_clash(const double* frag_lib, const int64_t* frag_off, const int32_t* n_atoms, int32_t F,
                    int32_t A_total, const int32_t* conf, const double* R, const double* t, int64_t P,
                    double t2, double thresh, int64_t max_clashes, uint8_t* verdict,
                    uint64_t* near_count, void* stream);
/* compenetration_check on materialised structures S (P, A, 3) (embedder.py:1245-1248).
 *   F = 0 (ids NULL): intramolecular count of 0 < d < 0.5 over the full symmetric matrix
 *   (numba_functions.py:49-56, t2_half = image of 0.5); F = 2: only ids[0] is read; F = 3. */
int tsc_clash_structs(const double* S, int64_t P, int32_t A, const int32_t* ids, int32_t F, double t2,
                      double thresh, double t2_half, int64_t max_clashes, uint8_t* verdict,
                      uint64_t* near_count, void* stream);
/* get_embed (embeds.py:961-969) for poses keep_idx (n_keep) int64 (NULL = first n_keep poses):
 *   S_out (n_keep, A_total, 3). */
int tsc_embed_gather(const double* frag_lib, const int64_t* frag_off, const int32_t* n_atoms, int32_t F,
                     int32_t A_total, const int32_t* conf, const double* R, const double* t,
                     const int64_t* keep_idx, int64_t n_keep, double* S_out, void* stream);

/* Pose parameters of the string embed generated on the device (embeds.py:91-114):
 *   mol2.rotation = rot_mat_from_pointer(ref_vec, angle) @ rotation_matrix_from_vectors(mol_vec, -ref_vec)
 *   (first factor only for angle != 0; utils.py:183-208, algebra.py:325-344), mol2.position = p1 - rotation @ p2,
 * for the whole pose space conformers (c1, c2) x reactive centres (ai1, ai2) x angles in the reference's loop
 * order, P = n_conf1*n_conf2*n_c1*n_c2*n_ang.  centers1/vecs1 (n_conf1, n_c1, 3) = ra1.center / ra1.orb_vecs of
 * every conformer of molecule 1, centers2/vecs2 likewise; sin_half/cos_half (n_ang) = sin, cos of angle/2 in
 * radians evaluated on the host, nonzero (n_ang) uint8 = angle != 0; flip (9) = rot_mat_from_pointer([0,0,1], 180)
 * evaluated on the host.  Out: conf (P, 2) int32, R (P, 2, 3, 3), t (P, 2, 3) as tsc_embed_clash reads them. */
int tsc_string_embed_params(const double* centers1, const double* vecs1, const double* centers2,
                            const double* vecs2, int32_t n_conf1, int32_t n_conf2, int32_t n_c1, int32_t n_c2,
                            const double* sin_half, const double* cos_half, const uint8_t* nonzero,
                            int32_t n_ang, const double* flip, int32_t* conf, double* R, double* t, void* stream);

/* Pose parameters of the cyclical embeds generated on the device (embeds.py:657-709), G groups (combination of
 * conformers, pivots and polygon orientation) x C angle combinations (embedder.systematic_angles) x F molecules:
 *   A = align_vec_pair([end - start, direction], [pivot, mol_direction])   (algebra.py:258-282, once per group
 *   and molecule), axis = A @ axis_src, Sr = rot_mat_from_pointer(axis, angle), rotation = Sr @ A,
 *   position = A @ apm - Sr @ (A @ apm) + vmean - A @ pmean.
 * (G, F, ...) inputs: ref2 (2,3), tgt2 (2,3), axis_src (3) = rc0 - rc1 or pivot, apm (3) = atomic_pivot_mean,
 * vmean (3) = mean(vec_pair), pmean (3) = pivot.meanpoint, gconf int32 conformer index; combos (C, F) int32 index
 * into the angle table; sin_half / cos_half = host-evaluated sin, cos of angle/2; scratch G*F*18 doubles.
 * Out: conf (G*C, F) int32, R (G*C, F, 3, 3), t (G*C, F, 3); pose g*C + c. */
int tsc_cyclical_embed_params(const double* ref2, const double* tgt2, const double* axis_src, const double* apm,
                              const double* vmean, const double* pmean, const int32_t* gconf, int64_t G, int32_t F,
                              const int32_t* combos, int64_t C, const double* sin_half, const double* cos_half,
                              double* scratch, int32_t* conf, double* R, double* t, void* stream);

/* ---- prune_conformers_rmsd_rot_corr ---------------------------------------------------------- */
/* Rotor-corrected RMSD of every pair (i in [row_begin,row_end), j > i), stateless from the centred
 * structures Sc (N, A, 3)  (rotationally_corrected_rmsd, torsion_module.py:953-1011).
 *   heavy (A) uint8; T <= 10 rotors: tor_i2/tor_i3 (T) int32 = bond atoms (axis = x[i2]-x[i3], centre
 *   x[i3], utils.py:389-414); n_ang (T) <= 6; sin_half/cos_half (T, 6) of angle/2 (host-computed so the
 *   quaternion of algebra.py:325-344 is bit-identical); rot_mask (T, A) uint8 = _get_rotation_mask
 *   (:301-325); node_mask (T, A) uint8 = heavy atoms of the rotor's sub-graph (:964-977).
 *   sim_bits (N, (N+31)/32) uint32: rows [row_begin,row_end) are zeroed, then bit j of row i set iff
 *   rmsd < max_rmsd (:1118).  codes (N, N) uint64 or NULL (at most 20 rotors): best angle index of rotor t in bits
 *   [3t, 3t+3).  rmsd_out (N, N) or NULL.  near_count (1) uint64: pairs within 1e-6 A of max_rmsd. */
int tsc_rotcorr_pairs(const double* Sc, int64_t N, int32_t A, const uint8_t* heavy, int32_t T,
                      const int32_t* tor_i2, const int32_t* tor_i3, const int32_t* n_ang,
                      const double* sin_half, const double* cos_half, const uint8_t* rot_mask,
                      const uint8_t* node_mask, int64_t row_begin, int64_t row_end, double max_rmsd,
                      uint32_t* sim_bits, uint64_t* codes, double* rmsd_out, uint64_t* near_count,
                      void* stream);
/* [host] Inner loop of the grouping replay for one chunk [base, hi) of a ladder round, driven by the forward
 * scan's results (all pointers are HOST pointers): row i visits the columns (reach[i], min(first_hit[i], hi) - 1]
 * and first_hit[i] if inside the chunk; state[j] = (best_angle(i, j) + state[i]) mod 360 for every visited j
 * (torsion_module.py:1004-1008 as rotor-state algebra); matches are returned chunk-relative in the reference's
 * insertion order.  compact[off[i] + k] = 3-bit-per-rotor code of pair (i, i + 1 + k); ang_table (T, 6). */
int64_t tsc_host_rotcorr_chunk(int64_t base, int64_t hi, const int64_t* first_hit, int64_t* reach, double* state,
                               int32_t T, const uint64_t* compact, const int64_t* off, const double* ang_table,
                               int32_t* match_i, int32_t* match_j);
/* Forward scan for the stateless mode at large N: first_hit[i] = min{ j > i : rmsd(i, j) < max_rmsd } (N if
 * none) for rows row_begin, row_begin + row_stride, ... < row_end (row_stride = number of ranks when the rows are
 * dealt round-robin; 1 otherwise).  The grouping loop (torsion_module.py:1098-1125) stops at the first
 * similar later structure and caches dissimilar pairs, so first_hit plus the best-angle codes of the pairs
 * (i, j <= first_hit[i]) — written into the dense (N, N) codes / rmsd_out arrays, either may be NULL — is all
 * it needs.  One CTA per row, 8 pairs at a time, early exit.  row_counter: one int32 of scratch. */
int tsc_rotcorr_scan(const double* Sc, int64_t N, int32_t A, const uint8_t* heavy, int32_t T,
                     const int32_t* tor_i2, const int32_t* tor_i3, const int32_t* n_ang,
                     const double* sin_half, const double* cos_half, const uint8_t* rot_mask,
                     const uint8_t* node_mask, int64_t row_begin, int64_t row_end, int64_t row_stride,
                     double max_rmsd, int32_t* first_hit, uint64_t* codes, double* rmsd_out, uint64_t* near_count,
                     int32_t* row_counter, void* stream);
/* Stateful mode — exact emulation of the reference's in-place mutation (utils.py:412 through
 * torsion_module.py:984-1008), one row of the grouping loop (:1101-1125) at a time:
 *   tsc_rotcorr_row evaluates (i, js[k]) for k < n from the CURRENT structures cur (N, A, 3) and stages
 *   the rotor-corrected copy of js[k] in staged (n, A, 3) together with rmsd (n) and codes (n);
 *   tsc_rotcorr_commit writes the first n_accept staged copies back into cur (the structures the
 *   reference visited before its `break`).  Rotor descriptors as in tsc_rotcorr_pairs. */
int tsc_rotcorr_row(const double* cur, int64_t N, int32_t A, const uint8_t* heavy, int32_t T,
                    const int32_t* tor_i2, const int32_t* tor_i3, const int32_t* n_ang,
                    const double* sin_half, const double* cos_half, const uint8_t* rot_mask,
                    const uint8_t* node_mask, int64_t i, const int32_t* js, int32_t n, double* rmsd,
                    uint64_t* codes, double* staged, void* stream);
int tsc_rotcorr_commit(double* cur, const double* staged, const int32_t* js, int32_t n_accept, int32_t A,
                       void* stream);

/* Structures idx (n) int64 with rotor t turned by its accumulated angle (sin_half/cos_half (n, T)),
 * in torsion order about the current axis: what the reference's in-place mutation leaves behind
 * in the structures it returns (torsion_module.py:1004-1008, :1161).  out (n, A, 3). */
int tsc_rotcorr_apply(const double* Sc, int64_t n, int32_t A, const int64_t* idx, int32_t T,
                      const int32_t* tor_i2, const int32_t* tor_i3, const double* sin_half,
                      const double* cos_half, const uint8_t* rot_mask, double* out, void* stream);

/* [host] Work items of the default screen for a persistent grid of n_ctas CTAs (read by tsc_rmsd_screen; the rule is
 * tscode_b200/_host.py: build_items_balanced): one contiguous, equally expensive stretch of (panel, j tile) pairs per CTA,
 * one item {panel, first j tile, j tile count, local row block} per panel a stretch touches, laid out round by round
 * with stride n_ctas (empty items, count 0, end a CTA's list).  The reference has no counterpart: this schedules the
 * pair loop of rmsd_pruning.py:43-79 over the SMs.  row_blocks (n_rb): the rank's 32-row blocks, ascending;
 * panel_lo / panel_hi: only these 128-row panels (panel_hi < 0: all); tile_j: 32, 48 or 64 columns per j tile
 * (<= 0: tiles_per_panel tiles per panel); max_item > 0 cuts items (measurement aid).  out (cap, 4) int32 or NULL.
 * Returns the number of items (at most cap are written), -1 on bad arguments. */
int64_t tsc_host_screen_items(int64_t N, const int32_t* row_blocks, int64_t n_rb, int32_t n_ctas, int64_t panel_lo,
                              int64_t panel_hi, double item_cost, int32_t max_item, int32_t tiles_per_panel,
                              int32_t tile_j, int32_t* out, int64_t cap);

/* [host] Every structure minus its centroid (torsion_module.py:1023, the first statement of
 * prune_conformers_rmsd_rot_corr: `np.array([s - s.mean(axis=0) for s in structures])`), bit-identical to numpy: the A
 * rows are added in order, the sums divided by A, then subtracted.  S, out (N, A, 3) HOST doubles (out may alias S).
 * Structures are dealt to n_threads host threads.  Returns 0, or -1 on bad arguments. */
int32_t tsc_host_centre(const double* S, int64_t N, int32_t A, double* out, int32_t n_threads);

/* ---- ensemble I/O ------------------------------------------------------------------------ */
/* [host] Multi-frame XYZ text of n_frames structures (utils.py:114-126, write_xyz; embedder.py:996-1043): per frame
 * "<A>\n<title>\n" and per atom '%s     % .6f % .6f % .6f\n', byte-identical to the reference's Python formatting.
 * coords (n_frames, A, 3) HOST doubles; symbols: A zero-terminated strings at a stride of 4 bytes; titles: n_frames
 * zero-terminated strings back to back, or NULL for "temp".  out == NULL returns the byte count; otherwise the
 * text is written (cap bytes available; returns -needed if too small).  Frames are formatted by n_threads threads. */
int64_t tsc_host_write_xyz(const double* coords, int64_t n_frames, int32_t A, const char* symbols, const char* titles,
                           char* out, int64_t cap, int32_t n_threads);

/* [host] Multi-frame XYZ text -> coordinates: the input side of the same format (utils.py:128-135, read_xyz = cclib's
 * ccread; callers use .atomcoords and .atomnos: hypermolecule_class.py:163-168, operators.py:109, 169, 285).  Follows
 * cclib's XYZ reader: per frame an optional blank line, the atom count, a comment line, `count` lines "symbol x y z
 * [ignored ...]"; an incomplete last frame is dropped; numbers converted like Python's float() (correctly rounded).
 * text (len bytes, HOST); *n_atoms receives A; coords (max_frames, A, 3) doubles or NULL (count only); symbols: A x 4
 * bytes (zero-terminated, those of the last frame) or NULL; title_span: (offset, length) of each comment line or NULL.
 * Returns the number of frames; -1 bad arguments, -2 malformed frame, -3 frames of different sizes, -4 a coordinate
 * that is not a number, -5 max_frames too small. */
int64_t tsc_host_read_xyz(const char* text, int64_t len, int32_t* n_atoms, double* coords, int64_t max_frames,
                          char* symbols, int64_t* title_span, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif
