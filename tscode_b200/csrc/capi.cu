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
extern "C" int64_t tsc_host_rotcorr_chunk(int64_t base, int64_t hi, const int64_t* first_hit, int64_t* reach,
                                          double* state, int32_t T, const uint64_t* compact, const int64_t* off,
                                          const double* ang_table, int32_t* match_i, int32_t* match_j) {
    int64_t n_match = 0;
    for (int64_t i = base; i < hi; i++) {
        const int64_t p = first_hit[i];
        const int64_t new_hi = (p - 1 < hi - 1) ? p - 1 : hi - 1;
        const int64_t lo = reach[i] + 1;
        const double* si = state + i * T;
        const uint64_t* ci = T > 0 ? compact + off[i] - (i + 1) : nullptr;      // code of (i, j) at ci[j]
        for (int64_t j = lo; T > 0 && j <= new_hi; j++) {
            const uint64_t c = ci[j];
            double* sj = state + j * T;
            for (int t = 0; t < T; t++) sj[t] = fmod(ang_table[t * 6 + ((c >> (3 * t)) & 7ull)] + si[t], 360.0);
        }
        if (new_hi >= lo) reach[i] = new_hi;
        if (p < hi) {
            const uint64_t c = T > 0 ? ci[p] : 0ull;
            double* sj = state + p * T;
            for (int t = 0; t < T; t++) sj[t] = fmod(ang_table[t * 6 + ((c >> (3 * t)) & 7ull)] + si[t], 360.0);
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
