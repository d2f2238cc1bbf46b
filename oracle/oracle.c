/*
 * oracle.c — CPU restatement of TSCoDe's conformer-ensemble hot path.
 *
 * TEST INFRASTRUCTURE.  This file is the checker, not the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so.  Nothing under tscode_b200/ links, imports or calls it.
 *
 * Parity pinning: the reference ships no golden vectors for this path, so
 * the oracle is pinned against outputs of the LIVE reference frozen by
 * oracle/gen_golden.py into tests/golden/ (tests/test_oracle_golden.py checks every one).
 *
 * Each function cites the reference file:line (relative to the TSCoDe tree) it restates.
 * Written from the algorithm's description, in C, with its own 3x3 SVD (one-sided Jacobi)
 * in place of LAPACK gesdd.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * 3x3 SVD  A = U diag(S) V^T, S sorted descending, U and V orthogonal (det = +-1).
 * Stands in for np.linalg.svd at rmsd_pruning.py:19 and algebra.py:272.
 * ---------------------------------------------------------------------------------------- */
static void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

static double det3(const double m[9]) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
           m[2] * (m[3] * m[7] - m[4] * m[6]);
}

static void svd3(const double A[9], double U[9], double S[3], double V[9]) {
    double B[9];
    memcpy(B, A, sizeof(B));
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotated = 0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double al = 0, be = 0, ga = 0;
                for (int i = 0; i < 3; i++) {
                    al += B[3 * i + p] * B[3 * i + p];
                    be += B[3 * i + q] * B[3 * i + q];
                    ga += B[3 * i + p] * B[3 * i + q];
                }
                if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                rotated = 1;
                double zeta = (be - al) / (2.0 * ga);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 3; i++) {
                    double bp = B[3 * i + p], bq = B[3 * i + q];
                    B[3 * i + p] = c * bp - s * bq;
                    B[3 * i + q] = s * bp + c * bq;
                    double vp = V[3 * i + p], vq = V[3 * i + q];
                    V[3 * i + p] = c * vp - s * vq;
                    V[3 * i + q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double n[3];
    int ord[3] = {0, 1, 2};
    for (int j = 0; j < 3; j++)
        n[j] = sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
    for (int a = 0; a < 2; a++)
        for (int b = a + 1; b < 3; b++)
            if (n[ord[b]] > n[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double Vs[9], u[3][3];
    for (int j = 0; j < 3; j++) {
        int o = ord[j];
        S[j] = n[o];
        for (int i = 0; i < 3; i++) Vs[3 * i + j] = V[3 * i + o];
        for (int i = 0; i < 3; i++) u[j][i] = (n[o] > 0) ? B[3 * i + o] / n[o] : 0.0;
    }
    memcpy(V, Vs, sizeof(Vs));
    /* complete U when rank-deficient (columns with ~zero singular value) */
    double tiny = S[0] * 1e-14;
    if (S[0] <= 0) { /* zero matrix: U = I */
        u[0][0] = 1; u[0][1] = 0; u[0][2] = 0;
        u[1][0] = 0; u[1][1] = 1; u[1][2] = 0;
    } else if (S[1] <= tiny) { /* rank 1: any unit vector orthogonal to u0 */
        double e[3] = {0, 0, 0};
        int k = (fabs(u[0][0]) <= fabs(u[0][1]) && fabs(u[0][0]) <= fabs(u[0][2])) ? 0
                : (fabs(u[0][1]) <= fabs(u[0][2]) ? 1 : 2);
        e[k] = 1.0;
        cross3(u[0], e, u[1]);
        double nn = sqrt(u[1][0] * u[1][0] + u[1][1] * u[1][1] + u[1][2] * u[1][2]);
        for (int i = 0; i < 3; i++) u[1][i] /= nn;
    }
    if (S[0] <= 0 || S[2] <= tiny) cross3(u[0], u[1], u[2]);
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) U[3 * i + j] = u[j][i];
}

/* ------------------------------------------------------------------------------------------
 * rmsd_and_max_numba(p, q)                                        rmsd_pruning.py:6-41
 * Rotation-only Kabsch about the origin (no centring): C = p^T q (:15), SVD (:19), improper
 * fix by negating the last column of the left factor (:20-23), R = v @ w (:26), p @ R (:29),
 * rmsd = sqrt(sum(diff^2)/M) (:35), max_delta = max row norm (:39).
 * ---------------------------------------------------------------------------------------- */
void orc_kabsch_rotation(const double *p, const double *q, int M, double R[9]) {
    double C[9] = {0};
    for (int m = 0; m < M; m++)
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) C[3 * a + b] += p[3 * m + a] * q[3 * m + b];
    double U[9], S[3], V[9];
    svd3(C, U, S, V);
    /* numpy: v = U, w = V^T; d = det(v) * det(w) < 0 -> v[:, -1] = -v[:, -1] */
    if (det3(U) * det3(V) < 0.0)
        for (int i = 0; i < 3; i++) U[3 * i + 2] = -U[3 * i + 2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += U[3 * i + k] * V[3 * j + k];
            R[3 * i + j] = s;
        }
}

void orc_rmsd_and_max(const double *p, const double *q, int M, double *rmsd, double *maxdev) {
    double R[9];
    orc_kabsch_rotation(p, q, M, R);
    double ss = 0, mx = 0;
    for (int m = 0; m < M; m++) {
        double d2 = 0;
        for (int j = 0; j < 3; j++) {
            double r = p[3 * m] * R[j] + p[3 * m + 1] * R[3 + j] + p[3 * m + 2] * R[6 + j];
            double d = r - q[3 * m + j];
            d2 += d * d;
        }
        ss += d2;
        double nrm = sqrt(d2);
        if (nrm > mx) mx = nrm;
    }
    *rmsd = sqrt(ss / (double)M);
    *maxdev = mx;
}

/* sim(i,j) = rmsd < thr and maxdev < 2*thr            rmsd_pruning.py:75, :95 (strict <) */
static int sim_pair(const double *p, const double *q, int M, double thr) {
    double r, d;
    orc_rmsd_and_max(p, q, M, &r, &d);
    return (r < thr) && (d < 2.0 * thr);
}

/* All-pairs similarity bytes (and optionally rmsd / maxdev) for rows [row_begin,row_end),
 * columns j > i; everything else is written as 0.  Checker for the GPU sim bits. */
void orc_sim_rows(const double *H, long N, int M, double thr, long row_begin, long row_end,
                  unsigned char *sim_out, double *rmsd_out, double *maxdev_out) {
    long stride = 3L * M;
#pragma omp parallel for schedule(dynamic, 4)
    for (long i = row_begin; i < row_end; i++) {
        unsigned char *srow = sim_out + (i - row_begin) * N;
        for (long j = 0; j < N; j++) {
            double r = 0, d = 0;
            int s = 0;
            if (j > i) {
                orc_rmsd_and_max(H + i * stride, H + j * stride, M, &r, &d);
                s = (r < thr) && (d < 2.0 * thr);
            }
            srow[j] = (unsigned char)s;
            if (rmsd_out) rmsd_out[(i - row_begin) * N + j] = r;
            if (maxdev_out) maxdev_out[(i - row_begin) * N + j] = d;
        }
    }
}

/* Throughput leg for bench.py: evaluate n_pairs explicit (i,j) pairs with all threads. */
long orc_eval_pairs(const double *H, int M, double thr, const long *ii, const long *jj, long n_pairs) {
    long stride = 3L * M, hits = 0;
#pragma omp parallel for schedule(static) reduction(+ : hits)
    for (long k = 0; k < n_pairs; k++) hits += sim_pair(H + ii[k] * stride, H + jj[k] * stride, M, thr);
    return hits;
}

/* ------------------------------------------------------------------------------------------
 * The (first, second) cache: an open-addressing set of 64-bit keys.  The reference keeps a
 * numba typed List and tests membership by linear scan (rmsd_pruning.py:66); a set has the
 * same semantics.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t *slot; uint64_t cap, n; } keyset;
#define EMPTY_KEY UINT64_MAX
static void ks_init(keyset *k, uint64_t cap) {
    k->cap = 64; while (k->cap < cap * 2) k->cap <<= 1;
    k->slot = (uint64_t *)malloc(k->cap * sizeof(uint64_t));
    memset(k->slot, 0xff, k->cap * sizeof(uint64_t));
    k->n = 0;
}
static inline uint64_t ks_hash(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; return x; }
static int ks_has(const keyset *k, uint64_t key) {
    uint64_t h = ks_hash(key) & (k->cap - 1);
    while (k->slot[h] != EMPTY_KEY) { if (k->slot[h] == key) return 1; h = (h + 1) & (k->cap - 1); }
    return 0;
}
static void ks_add(keyset *k, uint64_t key);
static void ks_grow(keyset *k) {
    keyset n; ks_init(&n, k->cap);
    for (uint64_t i = 0; i < k->cap; i++) if (k->slot[i] != EMPTY_KEY) ks_add(&n, k->slot[i]);
    free(k->slot); *k = n;
}
static void ks_add(keyset *k, uint64_t key) {
    if ((k->n + 1) * 2 > k->cap) ks_grow(k);
    uint64_t h = ks_hash(key) & (k->cap - 1);
    while (k->slot[h] != EMPTY_KEY) { if (k->slot[h] == key) return; h = (h + 1) & (k->cap - 1); }
    k->slot[h] = key; k->n++;
}

/* ------------------------------------------------------------------------------------------
 * prune_conformers_rmsd on the heavy-atom array                   rmsd_pruning.py:164-206
 *   ladder over k (:186-188) gated by `k == 1 or 20*k < count_nonzero(mask)` (:192);
 *   one round = _similarity_mask_rmsd_group (:123-162): chunksize = int(N // k) (:136), last
 *   chunk takes the remainder (:141-144); inside a chunk every active row i walks the active
 *   later rows of the SAME chunk against the ROUND-START mask (:98-113, :57-60); the cache
 *   key is (first, first+1+offset-in-tail) = (first, first + j - i) (:65); a cache hit stops
 *   the walk and keeps i (:66-67); a similar pair appends its key and drops i (:75-77); the
 *   cache is extended only after the round (:204).
 * If sim_bytes != NULL it is an N*N byte matrix of precomputed sim(i,j) (i<j) and no RMSD is
 * evaluated: that is how the tests replay the elimination on top of GPU-computed bits.
 * Returns number of survivors; n_eval = pair evaluations actually performed (the reference
 * is lazy); rounds_out (optional, 18 slots) = k of every round run, 0-ended.
 * ---------------------------------------------------------------------------------------- */
static const double LADDER[18] = {5e5, 2e5, 1e5, 5e4, 2e4, 1e4, 5000, 2000, 1000, 500,
                                  200, 100, 50, 20, 10, 5, 2, 1};

long orc_prune_rmsd(const double *H, long N, int M, double thr, const unsigned char *sim_bytes,
                    unsigned char *mask_out, long long *n_eval, double *rounds_out) {
    long stride = 3L * M;
    unsigned char *mask = (unsigned char *)malloc(N > 0 ? N : 1);
    unsigned char *next = (unsigned char *)malloc(N > 0 ? N : 1);
    memset(mask, 1, N);
    keyset cache; ks_init(&cache, 1024);
    long long evals = 0;
    int nround = 0;
    for (int li = 0; li < 18; li++) {
        double k = LADDER[li];
        long active = 0;
        for (long i = 0; i < N; i++) active += mask[i];
        if (!(k == 1 || 20 * k < (double)active)) continue;
        if (rounds_out) rounds_out[nround++] = k;
        long K = (long)k;
        long cs = (long)floor((double)N / k);         /* int(len(structures) // k) */
        uint64_t *newkeys = (uint64_t *)malloc(sizeof(uint64_t) * (N > 0 ? N : 1));
        long n_new = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : evals)
        for (long c = 0; c < K; c++) {
            long first = c * cs, last = (c == K - 1) ? N : cs * (c + 1);
            for (long i = first; i < last; i++) {
                if (!mask[i]) { next[i] = 0; continue; }
                int keep = 1;
                for (long j = i + 1; j < last; j++) {
                    if (!mask[j]) continue;
                    uint64_t key = ((uint64_t)first << 32) | (uint64_t)(first + j - i);
                    if (ks_has(&cache, key)) break;                     /* :66-67 keep, stop */
                    int s;
                    if (sim_bytes) s = sim_bytes[i * N + j];
                    else { s = sim_pair(H + i * stride, H + j * stride, M, thr); evals++; }
                    if (s) {                                            /* :75-77 drop, emit */
                        keep = 0;
                        long slot;
#pragma omp atomic capture
                        slot = n_new++;
                        newkeys[slot] = key;
                        break;
                    }
                }
                next[i] = (unsigned char)keep;
            }
        }
        for (long t = 0; t < n_new; t++) ks_add(&cache, newkeys[t]);   /* :204 */
        free(newkeys);
        unsigned char *tmp = mask; mask = next; next = tmp;
    }
    if (rounds_out && nround < 18) rounds_out[nround] = 0;
    long surv = 0;
    for (long i = 0; i < N; i++) { mask_out[i] = mask[i]; surv += mask[i]; }
    if (n_eval) *n_eval = evals;
    free(mask); free(next); free(cache.slot);
    return surv;
}

/* _rmsd_similarity(ref, structures, rmsd_thr)                    rmsd_pruning.py:208-224 */
int orc_rmsd_similarity(const double *ref, const double *structs, long n, int A, double thr) {
    for (long s = 0; s < n; s++)
        if (sim_pair(ref, structs + s * 3L * A, A, thr)) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * compenetration_check(coords, ids, thresh, max_clashes)          numba_functions.py:59-105
 * with all_dists (algebra.py:98-157: d = sqrt(sum_k (a_k-b_k)^2), float64) and
 * count_clashes (numba_functions.py:49-56).
 *   ids == NULL : count over the FULL symmetric A x A matrix of (d < 0.5) & (d > 0)   (:71-72)
 *   F == 2      : only ids[0] is read; count d(m2,m1) < thresh                         (:74-81)
 *   else (3)    : cumulative count over (m2,m1), (m3,m2), (m1,m3) with early returns   (:85-105)
 * Returns int 0/1.
 * ---------------------------------------------------------------------------------------- */
static long count_close(const double *a, long na, const double *b, long nb, double thresh) {
    long c = 0;
    for (long i = 0; i < na; i++)
        for (long j = 0; j < nb; j++) {
            double dx = a[3 * i] - b[3 * j], dy = a[3 * i + 1] - b[3 * j + 1], dz = a[3 * i + 2] - b[3 * j + 2];
            double d = sqrt(dx * dx + dy * dy + dz * dz);
            c += (d < thresh);
        }
    return c;
}

int orc_compenetration_check(const double *coords, long A, const long *ids, int F, double thresh,
                             long max_clashes) {
    if (ids == NULL || F == 0) {
        long c = 0;
        for (long i = 0; i < A; i++)
            for (long j = 0; j < A; j++) {
                double dx = coords[3 * i] - coords[3 * j], dy = coords[3 * i + 1] - coords[3 * j + 1],
                       dz = coords[3 * i + 2] - coords[3 * j + 2];
                double d = sqrt(dx * dx + dy * dy + dz * dz);
                c += (d < 0.5) && (d > 0);
            }
        return c > max_clashes ? 0 : 1;
    }
    if (F == 2) {
        long n0 = ids[0];
        return count_close(coords + 3 * n0, A - n0, coords, n0, thresh) > max_clashes ? 0 : 1;
    }
    long n0 = ids[0], n1 = ids[1];
    const double *m1 = coords, *m2 = coords + 3 * n0, *m3 = coords + 3 * (n0 + n1);
    long n2 = A - n0 - n1, clashes = 0;
    clashes += count_close(m2, n1, m1, n0, thresh);
    if (clashes > max_clashes) return 0;
    clashes += count_close(m3, n2, m2, n1, thresh);
    if (clashes > max_clashes) return 0;
    clashes += count_close(m1, n0, m3, n2, thresh);
    if (clashes > max_clashes) return 0;
    return 1;
}

/* get_embed(mols, conf_ids): concatenate (R_k @ X_k.T).T + t_k    embeds.py:961-969
 * frag_lib: all fragments' conformers back to back; frag_off[k] = offset (in doubles) of
 * fragment k's conformer 0; conformer c of fragment k starts at frag_off[k] + c*3*n_atoms[k]. */
void orc_get_embed(const double *frag_lib, const long *frag_off, const int *n_atoms, int F,
                   const long *conf, const double *R, const double *t, double *out) {
    long o = 0;
    for (int k = 0; k < F; k++) {
        const double *X = frag_lib + frag_off[k] + conf[k] * 3L * n_atoms[k];
        const double *r = R + 9 * k, *tt = t + 3 * k;
        for (int a = 0; a < n_atoms[k]; a++, o++)
            for (int i = 0; i < 3; i++)
                out[3 * o + i] = (r[3 * i] * X[3 * a] + r[3 * i + 1] * X[3 * a + 1] + r[3 * i + 2] * X[3 * a + 2]) + tt[i];
    }
}

/* The generator inner loop: get_embed then compenetration_check per pose
 * (embeds.py:116-118, 713-714, 841-842), over P poses, OpenMP over poses. */
void orc_embed_clash_batch(const double *frag_lib, const long *frag_off, const int *n_atoms, int F,
                           const long *conf, const double *R, const double *t, long P, double thresh,
                           long max_clashes, unsigned char *verdict) {
    long A = 0, ids[8];
    for (int k = 0; k < F; k++) { A += n_atoms[k]; ids[k] = n_atoms[k]; }
#pragma omp parallel
    {
        double *pose = (double *)malloc(sizeof(double) * 3 * (A > 0 ? A : 1));
#pragma omp for schedule(static)
        for (long p = 0; p < P; p++) {
            orc_get_embed(frag_lib, frag_off, n_atoms, F, conf + p * F, R + p * F * 9, t + p * F * 3, pose);
            verdict[p] = (unsigned char)orc_compenetration_check(pose, A, ids, F, thresh, max_clashes);
        }
        free(pose);
    }
}

/* compenetration_refining's loop over materialised structures     embedder.py:1245-1248 */
void orc_clash_structs(const double *S, long P, long A, const long *ids, int F, double thresh,
                       long max_clashes, unsigned char *verdict) {
#pragma omp parallel for schedule(static)
    for (long p = 0; p < P; p++)
        verdict[p] = (unsigned char)orc_compenetration_check(S + p * 3 * A, A, ids, F, thresh, max_clashes);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
