// capi.cu — bookkeeping entry points of the C-ABI (see include/tscode_b200.h).
#include "tsc_common.cuh"

extern "C" int tsc_version(void) { return 100; }   // 0.1.0

extern "C" const char* tsc_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

extern "C" int64_t tsc_num_blocks_padded(int64_t N) { return tsc::num_blocks_padded(N); }
extern "C" int32_t tsc_num_slabs(int32_t M) { return tsc::num_slabs(M); }
extern "C" int64_t tsc_packed_doubles(int64_t N, int32_t M) {
    return (int64_t)tsc::num_slabs(M) * tsc::num_blocks_padded(N) * tsc::CHUNK_D;
}
extern "C" int32_t tsc_device_sm_count(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return sms;
}

// ------------------------------------------------------------------------------------------
// Host-side inner loop of the rot_corr grouping replay (tscode_b200/torsion_module.py,
// ladder_replay_scan): one chunk [base, hi) of one ladder round, rows in the reference's order
// (torsion_module.py:1098-1125).  Row i visits the columns (reach[i], min(first_hit[i], hi) - 1],
// which become cached, and first_hit[i] itself when it lies inside the chunk; every visited structure
// j is "mutated":  state[j] = (best_angle(i, j) + state[i]) mod 360  (:1004-1008 in rotor-state
// algebra).  Sequential by construction (state[i] may have been rewritten by an earlier row), a few
// flops per visit: plain C on the host.  All pointers are HOST pointers.
//   compact / off: best-angle codes of row i's pairs (i, i + 1 + k) at compact[off[i] + k], 3 bits per
//   rotor;  ang_table (T, 6) degrees;  match_i / match_j (capacity hi - base): chunk-relative matches
//   in the order the reference adds them.  Returns the number of matches.  T = 0 (no rotor states: the TFD
//   pruning, numba_functions.py:155-231, runs the same loop) ignores compact / off / ang_table / state.
// ------------------------------------------------------------------------------------------
#include <math.h>
#define TSC_HOST_MAX_T 21                 // rotcorr.cu: 3 bits of angle code per rotor in a 64-bit word
// (a + b) % 360.0 as Python evaluates it, for the values that occur (angles and states in [0, 360)): one exact
// subtraction instead of fmod; anything else takes fmod
static inline double add_mod360(double a, double b) {
    const double s = a + b;
    if (s >= 0.0 && s < 720.0) return s - (s >= 360.0 ? 360.0 : 0.0);      // x - 0.0 == x for every x >= 0
    double r = fmod(s, 360.0);
    if (r != 0.0 && r < 0.0) r += 360.0;                 // Python's % takes the sign of the divisor
    return r;
}
extern "C" int64_t tsc_host_rotcorr_chunk(int64_t base, int64_t hi, const int64_t* first_hit, int64_t* reach,
                                          double* state, int32_t T, const uint64_t* compact, const int64_t* off,
                                          const double* ang_table, int32_t* match_i, int32_t* match_j) {
    int64_t n_match = 0;
    double img[TSC_HOST_MAX_T * 8];         // img[t * 8 + v] = (angle v of rotor t + state of row i) mod 360
    if (T > TSC_HOST_MAX_T) return -1;
    for (int64_t i = base; i < hi; i++) {
        const int64_t p = first_hit[i];
        const int64_t new_hi = (p - 1 < hi - 1) ? p - 1 : hi - 1;
        const int64_t lo = reach[i] + 1;
        const bool hit = p < hi;
        const int64_t visits = (new_hi >= lo ? new_hi - lo + 1 : 0) + (hit ? 1 : 0);
        if (T > 0 && visits > 0) {
            const double* si = state + i * T;
            const uint64_t* ci = compact + off[i] - (i + 1);                    // code of (i, j) at ci[j]
            if (visits >= 8) {
                // every structure this row visits gets one of at most 6 images per rotor of row i's state: the
                // images are computed once per row (the same additions the per-visit form makes), the visits are
                // look-ups
                for (int t = 0; t < T; t++) {
                    for (int v = 0; v < 6; v++) img[t * 8 + v] = add_mod360(ang_table[t * 6 + v], si[t]);
                    img[t * 8 + 6] = img[t * 8 + 7] = 0.0;                       // codes 6, 7 do not occur
                }
                for (int64_t j = lo; j <= new_hi; j++) {
                    const uint64_t c = ci[j];
                    double* sj = state + j * T;
                    for (int t = 0; t < T; t++) sj[t] = img[t * 8 + ((c >> (3 * t)) & 7ull)];
                }
                if (hit) {
                    const uint64_t c = ci[p];
                    double* sj = state + p * T;
                    for (int t = 0; t < T; t++) sj[t] = img[t * 8 + ((c >> (3 * t)) & 7ull)];
                }
            } else {
                for (int64_t q = 0; q < visits; q++) {
                    const int64_t j = (q == visits - 1 && hit) ? p : lo + q;
                    const uint64_t c = ci[j];
                    double* sj = state + j * T;
                    for (int t = 0; t < T; t++) sj[t] = add_mod360(ang_table[t * 6 + ((c >> (3 * t)) & 7ull)], si[t]);
                }
            }
        }
        if (new_hi >= lo) reach[i] = new_hi;
        if (hit) {
            match_i[n_match] = (int32_t)(i - base);
            match_j[n_match] = (int32_t)(p - base);
            n_match++;
        }
    }
    return n_match;
}

// ------------------------------------------------------------------------------------------
// [host] Multi-frame XYZ text of an ensemble (tscode/utils.py:114-126, write_xyz, called once per structure by
// Embedder.write_structures, embedder.py:996-1043): per frame
//     "<n_atoms>\n<title>\n"  then per atom  '%s     % .6f % .6f % .6f\n' % (symbol, x, y, z)
// After the kernels, formatting 1e5-1e6 structures in Python ('%' per atom) is what dominates an embed run.
// Frames are formatted in parallel by a few host threads into their exact slots (every line has a computable
// length once the integer digits are known, so a first pass sizes the frames).  Numbers: correctly rounded
// like Python's / C's "% .6f" — the fast path scales by 1e6 and rounds to nearest; whenever the scaled value is
// within 1e-4 of a rounding boundary (where one binary rounding could flip the decimal one) snprintf decides.
//   coords (n_frames, A, 3) doubles; symbols: A zero-terminated strings packed at a stride of 4 bytes;
//   titles: n_frames zero-terminated strings back to back (or NULL -> "temp").  out == NULL: returns the size.
// ------------------------------------------------------------------------------------------
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>
namespace {
inline int fmt_fixed6(double x, char* o) {             // "% .6f": sign or space, digits, '.', 6 decimals
    if (!(fabs(x) < 1e15)) return snprintf(o, 400, "% .6f", x);      // inf / nan / huge
    const double ax = fabs(x), y = ax * 1e6;
    const double fl = floor(y), fr = y - fl;
    if (fabs(fr - 0.5) < 1e-4) return snprintf(o, 400, "% .6f", x);
    unsigned long long n = (unsigned long long)(fr > 0.5 ? fl + 1.0 : fl);
    const bool neg = signbit(x);                        // "-0.000000" for negative values that round to zero, like printf
    char tmp[32];
    int k = 0;
    for (int d = 0; d < 6; d++) { tmp[k++] = (char)('0' + n % 10); n /= 10; }
    tmp[k++] = '.';
    do { tmp[k++] = (char)('0' + n % 10); n /= 10; } while (n);
    int len = 0;
    o[len++] = neg ? '-' : ' ';
    while (k) o[len++] = tmp[--k];
    return len;
}
inline int64_t fmt_frame(const double* X, int A, const char* symbols, const char* title, char* o) {
    char* p = o;
    p += snprintf(p, 32, "%d\n", A);
    const size_t tl = strlen(title);
    memcpy(p, title, tl); p += tl; *p++ = '\n';
    for (int a = 0; a < A; a++) {
        const char* s = symbols + 4 * a;
        while (*s) *p++ = *s++;
        memcpy(p, "     ", 5); p += 5;
        p += fmt_fixed6(X[3 * a], p); *p++ = ' ';
        p += fmt_fixed6(X[3 * a + 1], p); *p++ = ' ';
        p += fmt_fixed6(X[3 * a + 2], p); *p++ = '\n';
    }
    return p - o;
}
}  // namespace

extern "C" int64_t tsc_host_write_xyz(const double* coords, int64_t n_frames, int32_t A, const char* symbols,
                                      const char* titles, char* out, int64_t cap, int32_t n_threads) {
    if (n_frames <= 0) return 0;
    std::vector<const char*> tptr((size_t)n_frames);
    {
        const char* t = titles;
        for (int64_t f = 0; f < n_frames; f++) {
            tptr[f] = titles ? t : "temp";
            if (titles) t += strlen(t) + 1;
        }
    }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    // one formatting pass: thread w formats the contiguous frame range [lo_w, hi_w) into its own buffer; the
    // buffers are then concatenated (out == NULL only reports the total)
    std::vector<std::vector<char>> bufs((size_t)n_threads);
    const int64_t per = (n_frames + n_threads - 1) / n_threads;
    {
        std::vector<std::thread> th;
        for (int w = 0; w < n_threads; w++)
            th.emplace_back([&, w]() {
                const int64_t lo = w * per, hi = (lo + per < n_frames) ? lo + per : n_frames;
                std::vector<char>& b = bufs[w];
                b.reserve((size_t)(hi > lo ? hi - lo : 0) * ((size_t)A * 44 + 32) + 4096);       // typical frame size
                size_t used = 0;
                for (int64_t f = lo; f < hi; f++) {
                    const size_t room = (size_t)A * 1300 + strlen(tptr[f]) + 64;     // 3 x "% .6f" of any double
                    if (b.size() < used + room) b.resize(used + room + (b.size() >> 1));
                    used += (size_t)fmt_frame(coords + f * (int64_t)A * 3, A, symbols, tptr[f], b.data() + used);
                }
                b.resize(used);
            });
        for (auto& t : th) t.join();
    }
    int64_t total = 0;
    for (auto& b : bufs) total += (int64_t)b.size();
    if (!out) return total;
    if (cap < total) return -total;
    {
        std::vector<std::thread> th;
        int64_t off = 0;
        for (int w = 0; w < n_threads; w++) {
            th.emplace_back([&, w, off]() { if (!bufs[w].empty()) memcpy(out + off, bufs[w].data(), bufs[w].size()); });
            off += (int64_t)bufs[w].size();
        }
        for (auto& t : th) t.join();
    }
    return total;
}

// ------------------------------------------------------------------------------------------
// [host] Centring of every structure on its centroid (torsion_module.py:1023,
//     structures = np.array([s - s.mean(axis=0) for s in structures])
// the first statement of prune_conformers_rmsd_rot_corr).  numpy reduces axis 0 of an (A, 3) array by adding the
// rows in order, then divides the three sums by A; the same operations in the same order here (no reassociation, no
// reciprocal), so the result is bit-identical — tests/test_capi_and_host.py compares with numpy.  Native because on
// 20 000 x 63 atoms the numpy statement is a third of the whole call; structures are dealt to a few host threads.
// ------------------------------------------------------------------------------------------
extern "C" int32_t tsc_host_centre(const double* S, int64_t N, int32_t A, double* out, int32_t n_threads) {
    if (N < 0 || A < 1 || !S || !out) return N == 0 ? 0 : -1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (N * (int64_t)A < 65536) n_threads = 1;
    const int64_t per = (N + n_threads - 1) / n_threads;
    auto work = [&](int64_t lo, int64_t hi) {
        const double cnt = (double)A;
        for (int64_t s = lo; s < hi; s++) {
            const double* x = S + s * (int64_t)A * 3;
            double* o = out + s * (int64_t)A * 3;
            double sx = x[0], sy = x[1], sz = x[2];
            for (int a = 1; a < A; a++) { sx += x[3 * a]; sy += x[3 * a + 1]; sz += x[3 * a + 2]; }
            const double mx = sx / cnt, my = sy / cnt, mz = sz / cnt;
            for (int a = 0; a < A; a++) { o[3 * a] = x[3 * a] - mx; o[3 * a + 1] = x[3 * a + 1] - my; o[3 * a + 2] = x[3 * a + 2] - mz; }
        }
    };
    if (n_threads == 1) { work(0, N); return 0; }
    std::vector<std::thread> th;
    for (int w = 0; w < n_threads; w++) {
        const int64_t lo = w * per, hi = (lo + per < N) ? lo + per : N;
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto& t : th) t.join();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Host-side plan of the default screen (rmsd_screen.cu, ScFrame): principal-axes frame and column weights from the
// first structure, and — from a fixed pseudo-random sample of pairs — how often the weighted Samuelson bound would
// leave a pair undecided, which is what decides between the screen's forms (_host.screen_plan states the rule).
// Native because it sits on the latency path of every prune call (numpy: ~2 ms; this: ~0.1 ms).  A speed device only.
// ------------------------------------------------------------------------------------------
namespace {
// eigen-decomposition of a symmetric 3x3 matrix by cyclic Jacobi: a -> diagonal, v rows = eigenvectors
void jacobi3(double a[3][3], double v[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off <= 1e-300 || off <= 1e-17 * (fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]))) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) {                 // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {                 // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {                 // rows of v = eigenvectors
                    const double vpk = v[p][k], vqk = v[q][k];
                    v[p][k] = c * vpk - s * vqk;
                    v[q][k] = s * vpk + c * vqk;
                }
            }
    }
}
}  // namespace

// pi, pj (K): fixed pseudo-random pairs i != j of [0, N) (splitmix64 of the pair number: the same on every rank)
extern "C" void tsc_host_sample_pairs(int64_t N, int32_t K, int64_t* pi, int64_t* pj) {
    for (int32_t k = 0; k < K; k++) {
        uint64_t z = 0x9E3779B97F4A7C15ull * (uint64_t)(2 * k + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        uint64_t y = 0x9E3779B97F4A7C15ull * (uint64_t)(2 * k + 2);
        y = (y ^ (y >> 30)) * 0xBF58476D1CE4E5B9ull; y = (y ^ (y >> 27)) * 0x94D049BB133111EBull; y ^= y >> 31;
        const int64_t i = N > 0 ? (int64_t)(z % (uint64_t)N) : 0;
        const int64_t j = N > 1 ? (i + 1 + (int64_t)(y % (uint64_t)(N - 1))) % N : 0;
        pi[k] = i; pj[k] = j;
    }
}

// S: host array (rows, A, 3); `first` = row of the structure the frame is taken from; pairs (pi[k], pj[k]) index rows
// of S.  Out: frame12 (Q row-major, t), ratio = looseness of the unweighted bound for the first structure's shape,
// undecided = fraction of the sampled pairs the weighted bound does not exclude.  Returns 0, or 1 for bad arguments.
extern "C" int tsc_host_screen_plan(const double* S, int32_t A, const int32_t* heavy_idx, int32_t M, int64_t first,
                                    const int64_t* pi, const int64_t* pj, int32_t K, double thr, double* frame12,
                                    double* ratio, double* undecided) {
    if (!S || !heavy_idx || !frame12 || !ratio || !undecided || M <= 0) return 1;
    for (int k = 0; k < 12; k++) frame12[k] = (k < 9) ? (k % 4 == 0 ? 1.0 : 0.0) : 1.0;
    *ratio = INFINITY;
    *undecided = 0.0;
    double Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, t[3] = {1, 1, 1};
    {
        const double* x = S + first * (int64_t)A * 3;
        double a[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, v[3][3];
        bool finite = true;
        for (int m = 0; m < M; m++) {
            const double* c = x + (int64_t)heavy_idx[m] * 3;
            for (int p = 0; p < 3; p++) {
                finite = finite && isfinite(c[p]);
                for (int q = 0; q < 3; q++) a[p][q] += c[p] * c[q];
            }
        }
        double l0[3] = {a[0][0], a[1][1], a[2][2]};
        const double tot0 = l0[0] + l0[1] + l0[2];
        if (finite && isfinite(tot0) && tot0 > 0.0) {
            jacobi3(a, v);
            double lam[3] = {a[0][0], a[1][1], a[2][2]};
            const double tot = lam[0] + lam[1] + lam[2];
            *ratio = sqrt(3.0 * (lam[0] * lam[0] + lam[1] * lam[1] + lam[2] * lam[2])) / tot;
            if (*ratio > 1.0005) {
                double w[3], lsum = 0.0;
                for (int b = 0; b < 3; b++) { lam[b] = lam[b] > 1e-4 * tot ? lam[b] : 1e-4 * tot; lsum += lam[b]; }
                for (int b = 0; b < 3; b++) w[b] = lsum / lam[b] * (1.0 + 1e-9);
                for (int rep = 0; rep < 2; rep++) {           // Gram-Schmidt to the library's 1e-13
                    for (int r = 0; r < 3; r++) {
                        for (int p = 0; p < r; p++) {
                            const double d = v[r][0] * v[p][0] + v[r][1] * v[p][1] + v[r][2] * v[p][2];
                            for (int k = 0; k < 3; k++) v[r][k] -= d * v[p][k];
                        }
                        const double n = sqrt(v[r][0] * v[r][0] + v[r][1] * v[r][1] + v[r][2] * v[r][2]);
                        for (int k = 0; k < 3; k++) v[r][k] /= n;
                    }
                }
                for (int r = 0; r < 3; r++) {
                    for (int k = 0; k < 3; k++) { Q[r][k] = v[r][k]; frame12[3 * r + k] = v[r][k]; }
                    t[r] = sqrt(w[r] / 3.0);
                    frame12[9 + r] = t[r];
                }
            }
        }
    }
    int und = 0;
    for (int32_t k = 0; k < K; k++) {
        const double* p = S + pi[k] * (int64_t)A * 3;
        const double* q = S + pj[k] * (int64_t)A * 3;
        double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, gi = 0.0, gj = 0.0, gw = 0.0;
        for (int m = 0; m < M; m++) {
            const double* a = p + (int64_t)heavy_idx[m] * 3;
            const double* b = q + (int64_t)heavy_idx[m] * 3;
            double ar[3], br[3];
            for (int r = 0; r < 3; r++) {
                ar[r] = Q[r][0] * a[0] + Q[r][1] * a[1] + Q[r][2] * a[2];
                br[r] = t[r] * (Q[r][0] * b[0] + Q[r][1] * b[1] + Q[r][2] * b[2]);
            }
            gi += a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
            gj += b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
            gw += br[0] * br[0] + br[1] * br[1] + br[2] * br[2];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) C[r][c] += ar[r] * br[c];
        }
        double f = 0.0;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) f += C[r][c] * C[r][c];
        const double lf = 0.5 * (gi + gj) - 0.5 * M * thr * thr - 1.7320508075688772 * 1.05e-3 * sqrt(gi * (gw > gj ? gw : gj));
        und += !(lf > 0.0 && 3.00004 * f < lf * lf);
    }
    *undecided = K > 0 ? (double)und / K : 0.0;
    return 0;
}

// ------------------------------------------------------------------------------------------
// The reference's survivor choice inside one chunk of a grouping loop (torsion_module.py:1136-1152,
// numba_functions.py:203-220, optimization_methods.py:341-355):
//     matches = set of (i, j) tuples;  G = nx.Graph(matches);  for every connected component keep `group[0]`
// in native code.  WHICH member of a component is group[0] depends on iteration orders of CPython sets and dicts, so
// this restates them exactly (CPython 3.8+ / networkx 3.x; torsion_module._cluster_rejects_fast is the same in Python
// and is checked against networkx itself; tests/test_capi_and_host.py compares all three on random graphs):
//   * a set is an open-addressing table of (hash, key): slot i = hash & mask, then up to 9 linear probes, then
//     i = (5 i + 1 + (perturb >>= 5)) & mask; it grows when 5 fill >= 3 mask, to the first power of two above 4 used
//     (2 used beyond 50 000 entries), re-inserting in table order; iteration is table order;
//   * hash(int) = the value; hash((a, b)) = the xxHash-style tuple hash of Objects/tupleobject.c;
//   * Graph.add_edges_from inserts nodes in order of first appearance in the edge iteration (dict order), adjacency
//     in insertion order; components are found in node order by a breadth-first search whose `seen` is a set;
//   * G.subgraph(c).nodes iterates a NEW set filled from c element by element when 2 |c| < |G|, else the nodes of G in
//     order, filtered by membership in c (coreviews.FilterAtlas.__iter__).
// It replaced 0.2 s of Python per BASELINE configs[3] prune (3 900 chunks).
// ------------------------------------------------------------------------------------------
namespace {
// Storage is kept between uses (reset() only rewinds to CPython's initial 8-slot table) and a slot is one small record
// (a BASELINE configs[3] prune builds ~1e5 of these sets and probes them ~1e6 times).  Two key types: PairKey = a tuple
// (a, b) with its tuple hash, IntKey = a non-negative int, whose hash is its value.  a == -1 marks a free slot.
struct PairKey {
    uint64_t h; int32_t a, b;
    bool free_slot() const { return a < 0; }
    bool same(const PairKey& o) const { return h == o.h && a == o.a && b == o.b; }
    uint64_t hash() const { return h; }
};
struct IntKey {
    int32_t a;
    bool free_slot() const { return a < 0; }
    bool same(const IntKey& o) const { return a == o.a; }
    uint64_t hash() const { return (uint64_t)(int64_t)a; }
};
template <class Key>
struct PySetModel {
    std::vector<Key> cur, alt;
    size_t mask = 7, fill = 0;
    static void prepare(std::vector<Key>& t, size_t n) {
        if (t.size() < n) t.resize(n);
        memset((void*)t.data(), 0xFF, n * sizeof(Key));          // every a = -1
    }
    PySetModel() { prepare(cur, 8); }
    void reset() { mask = 7; fill = 0; prepare(cur, 8); }
    size_t size_used() const { return fill; }
    bool slot_used(size_t s) const { return !cur[s].free_slot(); }
    const Key& key(size_t s) const { return cur[s]; }
    static void insert_clean(std::vector<Key>& t, size_t nmask, const Key& k) {
        const uint64_t hash = k.hash();
        size_t i = (size_t)hash & nmask;
        uint64_t perturb = hash;
        while (true) {
            size_t j = i;
            if (t[j].free_slot()) { t[j] = k; return; }
            if (i + 9 <= nmask)
                for (j = i + 1; j <= i + 9; j++)
                    if (t[j].free_slot()) { t[j] = k; return; }
            perturb >>= 5;
            i = (i * 5 + 1 + (size_t)perturb) & nmask;
        }
    }
    void resize(size_t minused) {
        size_t newsize = 8;
        while (newsize <= minused) newsize <<= 1;
        prepare(alt, newsize);
        for (size_t s = 0; s <= mask; s++)
            if (!cur[s].free_slot()) insert_clean(alt, newsize - 1, cur[s]);
        cur.swap(alt);
        mask = newsize - 1;
    }
    bool add(const Key& k) {                                     // true if the key was new
        const uint64_t hash = k.hash();
        size_t i = (size_t)hash & mask;
        uint64_t perturb = hash;
        while (true) {
            int probes = (i + 9 <= mask) ? 9 : 0;
            size_t j = i;
            bool placed = false;
            while (true) {
                if (cur[j].free_slot()) { cur[j] = k; placed = true; break; }
                if (cur[j].same(k)) return false;
                if (probes-- == 0) break;
                j++;
            }
            if (placed) break;
            perturb >>= 5;
            i = (i * 5 + 1 + (size_t)perturb) & mask;
        }
        fill++;
        if (fill * 5 >= mask * 3) resize(fill > 50000 ? fill * 2 : fill * 4);
        return true;
    }
};
inline uint64_t py_hash_int(int32_t v) { return (uint64_t)(int64_t)v; }          // (v >= 0 here; hash(-1) is -2 in CPython)
inline uint64_t py_hash_pair(int32_t x, int32_t y) {
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P5 = 2870177450012600261ULL;
    uint64_t acc = P5;
    const uint64_t lanes[2] = {py_hash_int(x), py_hash_int(y)};
    for (int k = 0; k < 2; k++) {
        acc += lanes[k] * P2;
        acc = (acc << 31) | (acc >> 33);
        acc *= P1;
    }
    acc += 2ULL ^ (P5 ^ 3527539ULL);
    return acc == (uint64_t)-1 ? 1546275796ULL : acc;
}
// per-thread working storage of tsc_host_cluster_rejects, kept between calls
struct ClusterScratch {
    PySetModel<PairKey> matches;
    PySetModel<IntKey> seen, nodes;
    std::vector<int32_t> node_of;             // dense index -> node, in order of first appearance
    std::vector<int32_t> idx_of;              // node -> dense index, -1 outside a call
    std::vector<int32_t> head, tail;          // per dense index: first / last adjacency entry (-1: none)
    std::vector<int32_t> adj_to, adj_next;    // adjacency entries in insertion order, chained per node
    std::vector<uint8_t> seen_all, in_comp;
    std::vector<int32_t> level, next_level, comp;
};
}  // namespace

// mi, mj (n): the chunk's matches (i_rel < j_rel, chunk-relative, in insertion order); n_nodes_max: the chunk length
// (every index < n_nodes_max).  rejects (>= number of distinct nodes): the members that are NOT kept.  Returns their
// number, or -1 for bad arguments.
extern "C" int64_t tsc_host_cluster_rejects(const int32_t* mi, const int32_t* mj, int64_t n, int64_t n_nodes_max,
                                            int32_t* rejects) {
    if (n < 0 || n_nodes_max <= 0 || (n > 0 && (!mi || !mj || !rejects))) return -1;
    if (n == 0) return 0;
    for (int64_t k = 0; k < n; k++)
        if (mi[k] < 0 || mj[k] < 0 || mi[k] >= n_nodes_max || mj[k] >= n_nodes_max) return -1;
    static thread_local ClusterScratch W;
    PySetModel<PairKey>& matches = W.matches;
    matches.reset();
    for (int64_t k = 0; k < n; k++) matches.add(PairKey{py_hash_pair(mi[k], mj[k]), mi[k], mj[k]});
    // Graph(matches): nodes in order of first appearance, adjacency in insertion order
    if (W.idx_of.size() < (size_t)n_nodes_max) W.idx_of.resize((size_t)n_nodes_max, -1);
    std::vector<int32_t>& node_of = W.node_of;
    std::vector<int32_t>& idx_of = W.idx_of;
    node_of.clear(); W.head.clear(); W.tail.clear(); W.adj_to.clear(); W.adj_next.clear();
    auto node_index = [&](int32_t v) {
        if (idx_of[v] < 0) { idx_of[v] = (int32_t)node_of.size(); node_of.push_back(v); W.head.push_back(-1); W.tail.push_back(-1); }
        return idx_of[v];
    };
    auto append = [&](int32_t iu, int32_t v) {
        const int32_t e = (int32_t)W.adj_to.size();
        W.adj_to.push_back(v); W.adj_next.push_back(-1);
        if (W.tail[iu] < 0) W.head[iu] = e; else W.adj_next[W.tail[iu]] = e;
        W.tail[iu] = e;
    };
    for (size_t s = 0; s <= matches.mask; s++) {
        if (!matches.slot_used(s)) continue;
        const int32_t u = matches.key(s).a, v = matches.key(s).b;
        const int32_t iu = node_index(u);
        const int32_t iv = node_index(v);
        bool dup = false;                                        // adj[u][v] = None is idempotent
        for (int32_t e = W.head[iu]; e >= 0; e = W.adj_next[e]) dup = dup || W.adj_to[e] == v;
        if (!dup) { append(iu, v); if (u != v) append(iv, u); }
    }
    const size_t n_nodes = node_of.size();
    W.seen_all.assign(n_nodes, 0); W.in_comp.assign(n_nodes, 0);
    int64_t n_rej = 0;
    std::vector<int32_t>&level = W.level, &next_level = W.next_level, &comp = W.comp;
    PySetModel<IntKey>&seen = W.seen, &nodes = W.nodes;
    for (size_t q0 = 0; q0 < n_nodes; q0++) {
        if (W.seen_all[q0]) continue;
        const int32_t v0 = node_of[q0];
        seen.reset();
        seen.add(IntKey{v0});
        W.seen_all[q0] = 1;
        next_level.assign(1, v0);
        while (!next_level.empty()) {
            level.swap(next_level);
            next_level.clear();
            for (int32_t x : level)
                for (int32_t e = W.head[idx_of[x]]; e >= 0; e = W.adj_next[e]) {
                    const int32_t w = W.adj_to[e];
                    uint8_t& in_seen = W.seen_all[idx_of[w]];        // `w not in seen` without probing the set model
                    if (!in_seen) { in_seen = 1; seen.add(IntKey{w}); next_level.push_back(w); }
                }
        }
        // nodes = set(iter(seen)); first = next(iter(nodes)) if 2 |nodes| < |G| else first node of G in nodes
        comp.clear();
        const bool small = 2 * seen.size_used() < n_nodes;
        if (small) nodes.reset();
        for (size_t s = 0; s <= seen.mask; s++)
            if (seen.slot_used(s)) {
                const int32_t x = seen.key(s).a;
                if (small) nodes.add(IntKey{x});
                comp.push_back(x);
                W.in_comp[idx_of[x]] = 1;
            }
        int32_t first = -1;
        if (small) {
            for (size_t s = 0; s <= nodes.mask && first < 0; s++)
                if (nodes.slot_used(s)) first = nodes.key(s).a;
        } else {
            for (size_t q = 0; q < n_nodes && first < 0; q++)
                if (W.in_comp[q]) first = node_of[q];
        }
        for (int32_t x : comp) {
            W.in_comp[idx_of[x]] = 0;
            if (x != first) rejects[n_rej++] = x;
        }
    }
    for (int32_t v : node_of) idx_of[v] = -1;
    return n_rej;
}

// The whole grouping loop of torsion_module.py:1076-1152 (and of its TFD / MOI siblings) on the host: for every k of the
// ladder whose gate is open (k == 1 or gate * k < active structures; gate = 5 there), the k chunks in order — the last
// one ends at the number of ACTIVE structures, the reference's quirk (:1093-1094) —, per chunk the first-hit walk with
// rotor-state algebra (tsc_host_rotcorr_chunk) and the survivor choice (tsc_host_cluster_rejects).  The Python loop
// around the two spent 25 us per chunk on 3 900 chunks of BASELINE configs[3]: 0.09 of 0.13 s.
//   final_mask (N bytes, all 1 on entry); scratch: 3 N int32.  Returns the number of chunks that had matches, -1 on
//   bad arguments.
extern "C" int64_t tsc_host_ladder_replay(int64_t N, const int64_t* ladder, int32_t n_ladder, int32_t gate,
                                          const int64_t* first_hit, int64_t* reach, double* state, int32_t T,
                                          const uint64_t* compact, const int64_t* off, const double* ang_table,
                                          uint8_t* final_mask, int32_t* scratch) {
    if (N <= 0 || !ladder || !first_hit || !reach || !final_mask || !scratch || T > TSC_HOST_MAX_T ||
        (T > 0 && (!state || !compact || !off || !ang_table)))
        return -1;
    int32_t* mi = scratch;
    int32_t* mj = scratch + N;
    int32_t* rej = scratch + 2 * N;
    int64_t busy = 0;
    for (int32_t q = 0; q < n_ladder; q++) {
        const int64_t k = ladder[q];
        int64_t num_active = 0;
        for (int64_t i = 0; i < N; i++) num_active += final_mask[i] != 0;
        if (!(k == 1 || (int64_t)gate * k < num_active)) continue;
        const int64_t d = N / k;
        for (int64_t step = 0; step < k; step++) {
            const int64_t base = d * step;
            const int64_t len = (step == k - 1) ? (num_active > base ? num_active - base : 0) : d;
            if (len <= 1) continue;
            const int64_t n = tsc_host_rotcorr_chunk(base, base + len, first_hit, reach, state, T, compact, off, ang_table, mi, mj);
            if (n <= 0) continue;
            const int64_t nr = tsc_host_cluster_rejects(mi, mj, n, len, rej);
            if (nr < 0) return -1;
            for (int64_t r = 0; r < nr; r++) final_mask[base + rej[r]] = 0;
            busy++;
        }
    }
    return busy;
}

// ------------------------------------------------------------------------------------------
// [host] Work items of the default screen for a persistent grid (rmsd_screen.cu reads them; the rule is stated in
// tscode_b200/_host.py: build_items_balanced, which the tests compare this with entry by entry): the (panel, j tile)
// pairs of the owned 128-row panels, panel after panel, are cut into one contiguous stretch of equal cost per CTA
// (tiles + item_cost tiles per item start), one item {panel, first j tile, j tile count, local row block} per panel a
// stretch touches; laid out round by round with stride n_ctas, CTAs with fewer items than the longest list getting
// empty items (count 0).  Native because every NEW ensemble size pays it before its first launch (the lists are
// cached per size): 12 ms of Python for 50 000 structures — nine lists: the whole ensemble and the eight upload
// chunks —, more than two whole prunes.
//   row_blocks (n_rb): the rank's 32-row blocks, ascending (a panel is owned when its first block is);
//   panel_hi < 0: all panels;  tile_j <= 0: `tiles_per_panel` j tiles per panel (the FP64 variants' 16-column tiles),
//   else tiles of tile_j columns, panel p starting at tile 128 p / tile_j;  max_item > 0 cuts items (measurement aid).
//   out: (cap, 4) int32 or NULL.  Returns the number of items (also when out is NULL or cap is too small: at most cap
//   items are written), -1 on bad arguments.
// ------------------------------------------------------------------------------------------
extern "C" int64_t tsc_host_screen_items(int64_t N, const int32_t* row_blocks, int64_t n_rb, int32_t n_ctas,
                                         int64_t panel_lo, int64_t panel_hi, double item_cost, int32_t max_item,
                                         int32_t tiles_per_panel, int32_t tile_j, int32_t* out, int64_t cap) {
    if (N < 0 || n_rb < 0 || (n_rb > 0 && !row_blocks) || n_ctas < 1 || (tile_j <= 0 && tiles_per_panel < 1)) return -1;
    struct Item { int32_t p, j, cnt, lb; };
    const int64_t PB = 4;                                        // 32-row blocks per 128-row panel
    const int64_t n_pan = (N + 127) / 128;
    const int64_t tpp = tile_j > 0 ? (128 / tile_j > 1 ? 128 / tile_j : 1) : tiles_per_panel;
    const int64_t njt = tile_j > 0 ? (n_pan * 128 + tile_j - 1) / tile_j : n_pan * tpp;
    auto first = [&](int64_t p) { return tile_j > 0 ? (128 * p) / tile_j : tpp * p; };
    if (panel_hi < 0) panel_hi = n_pan;
    std::vector<int64_t> pan_p, pan_lb;
    int64_t total = 0;
    for (int64_t lb = 0; lb < n_rb; lb++) {
        const int64_t ib = row_blocks[lb];
        if (ib < 0) return -1;
        if (ib % PB == 0 && panel_lo <= ib / PB && ib / PB < panel_hi) {
            pan_p.push_back(ib / PB); pan_lb.push_back(lb);
            total += njt - first(ib / PB);
        }
    }
    if (total == 0) return 0;
    const int64_t n_bins = std::max<int64_t>(1, std::min<int64_t>(n_ctas, total / tpp));
    std::vector<Item> flat;                                      // items in the order they are cut; bin ids ascend
    std::vector<int64_t> bin_of;
    double left = (double)total + item_cost * (double)((int64_t)pan_p.size() + n_bins);   // cost still to hand out
    int64_t b = 0;
    double budget = left / (double)n_bins;                       // what the current CTA may still take
    for (size_t q = 0; q < pan_p.size(); q++) {
        int64_t j = first(pan_p[q]);
        const int64_t end = njt;
        while (j < end) {
            const int64_t room = (int64_t)(budget - item_cost);
            if (room < 4 && b + 1 < n_bins) {                    // not worth starting an item here: next CTA
                b++;
                budget = left / (double)(n_bins - b);            // re-balance over the CTAs that are left
                continue;
            }
            const int64_t take = (b + 1 == n_bins) ? end - j : std::max<int64_t>(1, std::min<int64_t>(end - j, room));
            flat.push_back(Item{(int32_t)pan_p[q], (int32_t)j, (int32_t)take, (int32_t)pan_lb[q]});
            bin_of.push_back(b);
            j += take;
            budget -= (double)take + item_cost;
            left -= (double)take + item_cost;
        }
    }
    // non-empty bins in order; optionally the same stretches cut into short items
    std::vector<std::vector<Item>> bins;
    int64_t prev = -1;
    for (size_t e = 0; e < flat.size(); e++) {
        if (bin_of[e] != prev) { bins.emplace_back(); prev = bin_of[e]; }
        if (max_item > 0)
            for (int32_t o = 0; o < flat[e].cnt; o += max_item)
                bins.back().push_back(Item{flat[e].p, flat[e].j + o, std::min<int32_t>(max_item, flat[e].cnt - o), flat[e].lb});
        else
            bins.back().push_back(flat[e]);
    }
    size_t rounds = 0;
    for (auto& x : bins) rounds = std::max(rounds, x.size());
    const Item pad{(int32_t)pan_p[0], (int32_t)first(pan_p[0]), 0, (int32_t)pan_lb[0]};
    // one round: one entry per CTA, any grid works; several: stride n_ctas exactly, so that the entries a CTA visits
    // are the ones meant for it and an empty entry really ends its list
    size_t n_bins_out = bins.size();
    if (rounds > 1 && (size_t)n_ctas > n_bins_out) n_bins_out = (size_t)n_ctas;
    int64_t n_items = 0;
    for (size_t r = 0; r < rounds; r++) {
        size_t last = 0;
        for (size_t q = 0; q < bins.size(); q++) if (bins[q].size() > r) last = q;
        const size_t width = (r + 1 < rounds) ? n_bins_out : last + 1;
        for (size_t q = 0; q < width; q++) {
            const Item& it = (q < bins.size() && bins[q].size() > r) ? bins[q][r] : pad;
            if (out && n_items < cap) { int32_t* o = out + 4 * n_items; o[0] = it.p; o[1] = it.j; o[2] = it.cnt; o[3] = it.lb; }
            n_items++;
        }
    }
    return n_items;
}

// ------------------------------------------------------------------------------------------
// [host] Multi-frame XYZ text -> coordinates (the input side of SURVEY 8(f)-4: utils.py:128-135, read_xyz, a wrapper
// of cclib's ccread; hypermolecule_class.py:163-168 and operators.py:109, 169, 285 use .atomcoords and .atomnos of the
// result).  cclib is a third-party dependency absent from the reference tree; this follows the published algorithm
// of its XYZ reader (cclib/io/xyzreader.py, 1.7-1.8): per frame an optional single blank line, a line whose first
// token is the atom count, a comment line, then `count` lines of at least four whitespace-separated tokens — symbol,
// x, y, z, anything further ignored; the text may end anywhere (an incomplete last frame is dropped); the symbols
// reported are those of the last complete frame.  Numbers are converted like Python's float(): correctly rounded
// (exact fast path for up to 19 digits and |exponent| <= 22, otherwise strtod in the C locale).
// A first serial pass finds the frames (line ends only), then n_threads host threads tokenise and convert them.
//   coords: (max_frames, A, 3) doubles or NULL (with symbols NULL as well: frames counted, atom lines not checked); symbols: A x 4 bytes, zero-terminated, or NULL;
//   title_span: 2 int64 per frame (offset, length of the comment line) or NULL.
// Returns the number of complete frames; -1 bad arguments, -2 malformed frame (cclib: AssertionError / ValueError),
// -3 frames with different atom counts, -4 a coordinate that is not a number, -5 max_frames too small.
// ------------------------------------------------------------------------------------------
#include <locale.h>
#include <atomic>
namespace {
inline bool xyz_space(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f' || c == '\r' || c == '\n'; }
// next line [b, e) of [p, end): split at \n, \r\n or \r; returns false at the end of the text
template <bool BARE_CR>
inline bool xyz_next_line(const char*& p, const char* end, const char*& b, const char*& e) {
    if (p >= end) return false;
    b = p;
    if (BARE_CR) {
        while (p < end && *p != '\n' && *p != '\r') p++;
        e = p;
        if (p < end) { if (*p == '\r' && p + 1 < end && p[1] == '\n') p += 2; else p++; }
    } else {                                 // only \n and \r\n occur: a \r left at the end of a line is whitespace
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        e = nl ? nl : end;
        p = nl ? nl + 1 : end;
    }
    return true;
}
inline bool xyz_next_token(const char*& p, const char* e, const char*& tb, const char*& te) {
    while (p < e && xyz_space(*p)) p++;
    if (p >= e) return false;
    tb = p;
    while (p < e && !xyz_space(*p)) p++;
    te = p;
    return true;
}
inline bool xyz_blank(const char* b, const char* e) { while (b < e && xyz_space(*b)) b++; return b >= e; }
inline bool xyz_ieq(const char* b, const char* e, const char* word) {
    for (; b < e && *word; b++, word++) if ((*b | 0x20) != *word) return false;
    return b == e && !*word;
}
bool xyz_parse_double(const char* b, const char* e, double* out) {
    static const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                                   1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* p = b;
    bool neg = false;
    if (p < e && (*p == '+' || *p == '-')) { neg = *p == '-'; p++; }
    if (xyz_ieq(p, e, "inf") || xyz_ieq(p, e, "infinity")) { *out = neg ? -INFINITY : INFINITY; return true; }
    if (xyz_ieq(p, e, "nan")) { *out = neg ? -NAN : NAN; return true; }
    uint64_t mant = 0;
    int n_sig = 0, n_dig = 0;
    int64_t exp10 = 0;
    bool exact = true;
    for (; p < e && *p >= '0' && *p <= '9'; p++) {
        n_dig++;
        if (n_sig < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant) n_sig++; } else { exact = false; }
    }
    if (p < e && *p == '.') {
        p++;
        for (; p < e && *p >= '0' && *p <= '9'; p++) {
            n_dig++;
            if (n_sig < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant) n_sig++; exp10--; } else { exact = false; }
        }
    }
    if (n_dig == 0) return false;
    if (p < e && (*p == 'e' || *p == 'E')) {
        p++;
        bool eneg = false;
        if (p < e && (*p == '+' || *p == '-')) { eneg = *p == '-'; p++; }
        if (p >= e || *p < '0' || *p > '9') return false;
        int64_t ex = 0;
        for (; p < e && *p >= '0' && *p <= '9'; p++) if (ex < 100000) ex = ex * 10 + (*p - '0');
        exp10 += eneg ? -ex : ex;
    }
    if (p != e) return false;
    if (exact && mant < (1ull << 53) && exp10 >= -22 && exp10 <= 22) {        // one correctly rounded operation
        double v = (double)mant;
        v = exp10 >= 0 ? v * P10[exp10] : v / P10[-exp10];
        *out = neg ? -v : v;
        return true;
    }
    static locale_t c_loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    char tmp[512];
    const size_t n = (size_t)(e - b);
    std::vector<char> big;
    char* s = tmp;
    if (n >= sizeof(tmp)) { big.resize(n + 1); s = big.data(); }
    memcpy(s, b, n); s[n] = 0;
    char* endp = nullptr;
    *out = strtod_l(s, &endp, c_loc);
    return endp == s + n;
}
struct XyzFrame { const char* atoms; const char* title_b; const char* title_e; };
}  // namespace

template <bool BARE_CR>
static int64_t read_xyz_impl(const char* text, int64_t len, int32_t* n_atoms, double* coords, int64_t max_frames,
                             char* symbols, int64_t* title_span, int32_t n_threads) {
    const char* p = text;
    const char* end = text + len;
    std::vector<XyzFrame> frames;
    int64_t A = -1;
    // serial pass: frame boundaries only (count line, comment line, `count` lines skipped)
    while (true) {
        const char *b, *e, *tb, *te;
        if (!xyz_next_line<BARE_CR>(p, end, b, e)) break;
        if (xyz_blank(b, e) && !xyz_next_line<BARE_CR>(p, end, b, e)) break;  // one optional blank line
        const char* q = b;
        if (!xyz_next_token(q, e, tb, te)) return -2;                        // assert len(tokens) >= 1
        int64_t natom = 0;
        {
            const char* d = tb;
            if (d < te && (*d == '+' || *d == '-')) { if (*d == '-') return -2; d++; }
            if (d >= te) return -2;
            for (; d < te; d++) { if (*d < '0' || *d > '9' || natom > 100000000) return -2; natom = natom * 10 + (*d - '0'); }
        }
        if (natom < 1) return -2;
        XyzFrame f;
        if (!xyz_next_line<BARE_CR>(p, end, f.title_b, f.title_e)) break;    // comment line
        if (!BARE_CR && f.title_e > f.title_b && f.title_e[-1] == '\r') f.title_e--;
        f.atoms = p;
        bool complete = true;
        for (int64_t a = 0; a < natom; a++)
            if (!xyz_next_line<BARE_CR>(p, end, b, e)) { complete = false; break; }
        if (!complete) break;                                                // the text ended inside a frame: dropped
        if (A >= 0 && natom != A) return -3;
        A = natom;
        frames.push_back(f);
    }
    const int64_t n_frames = (int64_t)frames.size();
    *n_atoms = (int32_t)(A < 0 ? 0 : A);
    if (title_span)
        for (int64_t f = 0; f < n_frames && f < max_frames; f++) {
            title_span[2 * f] = frames[f].title_b - text;
            title_span[2 * f + 1] = frames[f].title_e - frames[f].title_b;
        }
    if (n_frames == 0 || (!coords && !symbols)) return n_frames;           // count only: the atom lines are not looked at
    if (coords && max_frames < n_frames) return -5;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (n_frames * A < 4096) n_threads = 1;
    // parallel pass: every atom line tokenised (>= 4 tokens or the frame is malformed), numbers converted
    std::atomic<int> bad{0};
    auto work = [&](int64_t lo, int64_t hi) {
        double sink[3];
        for (int64_t f = lo; f < hi && !bad.load(std::memory_order_relaxed); f++) {
            const char* q = frames[f].atoms;
            for (int64_t a = 0; a < A; a++) {
                const char *b, *e, *tb, *te;
                xyz_next_line<BARE_CR>(q, end, b, e);
                const char* w = b;
                if (!xyz_next_token(w, e, tb, te)) { bad.store(2); return; }              // symbol
                if (symbols && f == n_frames - 1) {
                    if (te - tb > 3) { bad.store(2); return; }                            // no element symbol is that long
                    memset(symbols + 4 * a, 0, 4);
                    memcpy(symbols + 4 * a, tb, (size_t)(te - tb));
                }
                double* o = coords ? coords + (f * A + a) * 3 : sink;
                for (int c = 0; c < 3; c++) {
                    if (!xyz_next_token(w, e, tb, te)) { bad.store(2); return; }          // assert len(tokens) >= 4
                    if (!xyz_parse_double(tb, te, o + c)) { bad.store(4); return; }
                }
            }
        }
    };
    if (n_threads == 1) work(0, n_frames);
    else {
        const int64_t per = (n_frames + n_threads - 1) / n_threads;
        std::vector<std::thread> th;
        for (int w = 0; w < n_threads; w++) {
            const int64_t lo = w * per, hi = std::min<int64_t>(lo + per, n_frames);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& t : th) t.join();
    }
    return bad.load() ? -(int64_t)bad.load() : n_frames;
}

extern "C" int64_t tsc_host_read_xyz(const char* text, int64_t len, int32_t* n_atoms, double* coords, int64_t max_frames,
                                     char* symbols, int64_t* title_span, int32_t n_threads) {
    if (!text || len < 0 || !n_atoms) return -1;
    bool bare_cr = false;                                        // a \r that is not part of \r\n: old Mac line ends
    for (const char* r = (const char*)memchr(text, '\r', (size_t)len); r && !bare_cr;
         r = (const char*)memchr(r + 1, '\r', (size_t)(text + len - r - 1)))
        bare_cr = r + 1 >= text + len || r[1] != '\n';
    return bare_cr ? read_xyz_impl<true>(text, len, n_atoms, coords, max_frames, symbols, title_span, n_threads)
                   : read_xyz_impl<false>(text, len, n_atoms, coords, max_frames, symbols, title_span, n_threads);
}
