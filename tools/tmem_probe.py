#!/usr/bin/env python
"""tcgen05.ld fragment layouts and read rates, tcgen05.mma chain timings (tools/probes/libtsc_probe.so).
Measurement aid for the (conformer, component)-row tiling of the screen; prints a report."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tscode_b200._lib import check, ptr, stream_ptr  # noqa: E402

P = C.CDLL(os.path.join(ROOT, "tools", "probes", "libtsc_probe.so"))
vp, i32 = C.c_void_p, C.c_int32
P.tsc_probe_ld_layout.argtypes = [vp, vp]
P.tsc_probe_ld_rate.argtypes = [i32, i32, i32, i32, vp, vp, vp]
P.tsc_bench_umma.argtypes = [i32, i32, i32, i32, vp, vp]

out = torch.zeros(5 * 2 * 128 * 4, dtype=torch.int32, device="cuda")
check(P.tsc_probe_ld_layout(ptr(out), stream_ptr()), "layout")
torch.cuda.synchronize()
o = out.cpu().numpy().view(np.uint32).reshape(5, 2, 128, 4)
names = ["16x256b.x1", "16x128b.x1", "16x64b.x1", "16x256b.x2 (2nd repetition)", "16x32bx2.x1 (+8)"]
for s in range(5):
    for half in range(2):
        print(f"== {names[s]}  lane offset {16 * half}: thread -> [(lane, col) per register], warp 0 and warp 1 thread 0..7")
        for t in list(range(0, 32)) + list(range(32, 40)):
            regs = ["(%3d,%2d)" % (v >> 16, v & 0xffff) if v != 0xffffffff else "   -    " for v in o[s, half, t]]
            print(f"   t{t:3d}: " + " ".join(regs))

cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
sink = torch.zeros(2, dtype=torch.int32, device="cuda")
mode_names = ["32x32b.x4 (4 per wait)", "16x256b.x4 x2 halves", "16x256b.x2 x2 halves", "32x32b.x16"]
for warps in (4, 8, 16):
    for mode in range(4):
        for ncols in (96, 192, 256):
            reps = 2000
            for _ in range(2):
                check(P.tsc_probe_ld_rate(ncols, reps, mode, warps, ptr(cyc), ptr(sink), stream_ptr()), "rate")
                torch.cuda.synchronize()
            c = int(cyc[0].item()) / reps
            # every warp reads ncols columns of its 32-lane quarter; warps/4 warps share a quarter (they read the SAME data)
            byts = warps * 32 * ncols * 4
            print(f"ld rate: warps={warps:2d} mode={mode_names[mode]:24s} ncols={ncols:3d}: {c:8.1f} cycles per pass "
                  f"({byts / c:7.1f} B/clk/SM delivered to registers)", flush=True)

o2 = torch.zeros(2, dtype=torch.int64, device="cuda")
reps = 2000
for tm in (0, 1):
    for N, nsets in ((144, -1), (144, -2), (144, -3), (192, -1), (192, -2), (208, -1), (208, -2), (224, -1), (224, -2), (240, -1),
                     (256, -1), (96, -2), (96, -4), (128, -1), (128, -2)):
        try:
            check(P.tsc_bench_umma(N, nsets, reps, tm, ptr(o2), stream_ptr()), "umma"); torch.cuda.synchronize()
            check(P.tsc_bench_umma(N, nsets, reps, tm, ptr(o2), stream_ptr()), "umma"); torch.cuda.synchronize()
        except Exception as e:
            print("umma", N, nsets, "failed", e)
            continue
        n = reps * (-nsets)
        c = int(o2[0].item()) / n
        print(f"umma: A_in_tmem={tm} N={N:3d} chains={-nsets}: {c:7.1f} cycles/MMA ({128 * N * 8 / c:7.0f} MAC/clk/SM, tf32 K=8)", flush=True)
