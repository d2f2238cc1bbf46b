#!/usr/bin/env python
"""clock64 timeline of CTA 0 of the default screen (measurement build in tools/probes/libtsc_probe.so):
per unit (tile, row a): MMA thread {loop top, B tile landed, buffer free, MMAs issued + committed},
epilogue warp 0 {top, accumulator full, loads landed}, epilogue warp 15 {buffer released}.
python tools/screen_trace.py N M pace"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tscode_b200._lib import check, ptr, stream_ptr  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble  # noqa: E402

P = C.CDLL(os.path.join(ROOT, "tools", "probes", "libtsc_probe.so"))
vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
P.tsc_rmsd_screen.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, vp, i32, f64, vp, vp, i64, i32, i32, i32, vp, vp]
P.tsc_screen_set_trace.argtypes = [vp]

N, M = int(sys.argv[1]), int(sys.argv[2])
for pace in [int(x) for x in sys.argv[3].split(",")]:           # here: the screen mode (0, 1, 2)
    S = gen_ensemble(3, N, M, N // 10)
    pr = RmsdPruner(S, np.full(M, 6), 0.5, variant="screen", screen_mode=pace)
    pr.pack()
    trace = torch.zeros(192 * 8, dtype=torch.int64, device="cuda")
    for rep in range(2):
        pr.cand_list[0].fill_(0)
        P.tsc_screen_set_trace(ptr(trace) if rep == 1 else None)
        check(P.tsc_rmsd_screen(ptr(pr.PA), ptr(pr.PB), ptr(pr.PR), ptr(pr.G), ptr(pr.sG), ptr(pr.CT), pr.N, pr.M, ptr(pr.items),
                                pr.n_items, pr.thr, ptr(pr.sim_bits), ptr(pr.cand_list), pr.cand_stride, 0, pace, 0, pr._frame_ptr(), stream_ptr()), "screen")
        torch.cuda.synchronize()
    P.tsc_screen_set_trace(None)
    tr = trace.cpu().numpy().reshape(192, 8)
    t0 = tr[0, 0]
    print(f"== mode {pace}: unit: MMA[top, B landed, buffer free, issued] | EPI[top, full, loads landed, released(w15)]  (cycles since start; deltas)")
    prev_rel = None
    for u in range(40, 76):
        r = tr[u] - t0
        print(f"u{u:3d} a={u % 3} buf={u % 4}: MMA {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} (wait buf {r[2] - r[1]:5d}, issue {r[3] - r[2]:5d}) | "
              f"EPI {r[4]:7d} {r[5]:7d} {r[6]:7d} {r[7]:7d} (wait full {r[5] - r[4]:5d}, load {r[6] - r[5]:5d}, rel {r[7] - r[6]:5d}; full-after-issue {r[5] - r[3]:5d})")
    per_unit = (tr[150, 7] - tr[30, 7]) / 120.0
    print(f"   steady state: {per_unit:.1f} cycles per unit, {3 * per_unit:.1f} per tile")
