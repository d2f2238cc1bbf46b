#!/usr/bin/env python
"""Where the fused ladder's time goes (eliminate.cu): per-phase globaltimer stamps of the product kernel on BASELINE
configs[2], and the kernel time for several grid sizes (measurement builds tools/probes/libtsc_elim_g*.so).
python tools/elim_probe.py [N]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tscode_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
S = gen_ensemble(3, N, 80, N // 10)
pr = RmsdPruner(torch.from_numpy(S).cuda(), np.full(80, 6), 0.5)
mask = pr.run().cpu().numpy()
print("digest", mask_digest(mask), "rounds", pr.rounds, "pairs", int(pr.pair_list[0, 0]))
info = pr.fused_out[pr._info_off:pr._info_off + 256].view(torch.int32).cpu().numpy()
st = info[32:62]
print("stamps (us since kernel start):", [round(x / 1e3, 1) for x in st if x > 0])
vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64


def time_lib(L, name):
    L.tsc_elim_fused.argtypes = [vp, i32, i64, i64, i32, vp, vp, vp]
    ts = []
    for rep in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(L.tsc_elim_fused(ptr(pr.pair_list), 1, pr.pair_stride, pr.N, 20, ptr(pr.fused_ws), ptr(pr.fused_out), stream_ptr()), name)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    m = pr.fused_out[:pr.N].to(torch.bool).cpu().numpy()
    print(f"{name:10s}: {min(ts):7.1f} us (median {sorted(ts)[len(ts) // 2]:.1f}) mask ok {np.array_equal(m, mask)}")


time_lib(lib(), "product")
for g in (96, 64, 32, 16):
    p = os.path.join(ROOT, "tools", "probes", f"libtsc_elim_g{g}.so")
    if os.path.exists(p):
        time_lib(C.CDLL(p), f"grid {g}")
