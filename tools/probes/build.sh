#!/bin/bash
# builds tools/probes/libtsc_probe.so (measurement aids; not part of the product library)
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -o libtsc_probe.so tmem_probe.cu screen_trace.cu peaks.cu
# grid-size sweep of the fused ladder (tools/elim_probe.py)
for g in 16 32 64 96; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DTSC_EF_GRID=$g -o libtsc_elim_g$g.so ../../tscode_b200/csrc/eliminate.cu
done
# the previous form of the list verification (4 lanes per candidate), for A/B runs (tools/verify_probe.py)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DTSC_VERIFY_PER_LANE -o libtsc_verify_perlane.so ../../tscode_b200/csrc/rmsd_verify.cu
