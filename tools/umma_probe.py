#!/usr/bin/env python
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200._lib import lib, check, ptr, stream_ptr
out = torch.zeros(2, dtype=torch.int64, device="cuda")
reps = 2000
for tm in (0, 1):
    for N, nsets in ((16, -3), (32, -3), (48, -1), (48, -3), (48, -6), (48, -9), (64, -3), (64, -6), (80, -3), (96, -3), (128, -3),
                     (144, -3), (192, -2), (256, -1)):
        check(lib().tsc_bench_umma(N, nsets, reps, tm, ptr(out), stream_ptr()), "umma"); torch.cuda.synchronize()
        check(lib().tsc_bench_umma(N, nsets, reps, tm, ptr(out), stream_ptr()), "umma"); torch.cuda.synchronize()
        cyc = int(out[0].item()); n = reps * (-nsets if nsets < 0 else nsets * 3)
        print(f"A_in_tmem={tm} N={N:3d} independent accumulators={(-nsets if nsets < 0 else nsets*3):2d}: {cyc/n:7.1f} cycles/MMA  ({128*N*8/(cyc/n):7.0f} MAC/clk/SM)", flush=True)
