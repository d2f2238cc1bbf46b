#!/bin/bash
# builds tools/probes/libtsc_probe.so (measurement aids; not part of the product library)
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -o libtsc_probe.so tmem_probe.cu screen_trace.cu peaks.cu
