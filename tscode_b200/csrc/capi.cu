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
                                          double* state, int32_t T, const uint32_t* compact, const int64_t* off,
                                          const double* ang_table, int32_t* match_i, int32_t* match_j) {
    int64_t n_match = 0;
    for (int64_t i = base; i < hi; i++) {
        const int64_t p = first_hit[i];
        const int64_t new_hi = (p - 1 < hi - 1) ? p - 1 : hi - 1;
        const int64_t lo = reach[i] + 1;
        const double* si = state + i * T;
        const uint32_t* ci = T > 0 ? compact + off[i] - (i + 1) : nullptr;      // code of (i, j) at ci[j]
        for (int64_t j = lo; T > 0 && j <= new_hi; j++) {
            const uint32_t c = ci[j];
            double* sj = state + j * T;
            for (int t = 0; t < T; t++) sj[t] = fmod(ang_table[t * 6 + ((c >> (3 * t)) & 7u)] + si[t], 360.0);
        }
        if (new_hi >= lo) reach[i] = new_hi;
        if (p < hi) {
            const uint32_t c = T > 0 ? ci[p] : 0u;
            double* sj = state + p * T;
            for (int t = 0; t < T; t++) sj[t] = fmod(ang_table[t * 6 + ((c >> (3 * t)) & 7u)] + si[t], 360.0);
            match_i[n_match] = (int32_t)(i - base);
            match_j[n_match] = (int32_t)(p - base);
            n_match++;
        }
    }
    return n_match;
}
