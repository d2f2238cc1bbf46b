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
