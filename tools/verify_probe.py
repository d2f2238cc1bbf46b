#!/usr/bin/env python
"""A/B of the two list-verification kernels (rmsd_verify.cu): the product's coalesced form against the previous
4-lanes-per-candidate form (measurement build tools/probes/libtsc_verify_perlane.so) on the same screened bits, for a
range of atom counts; both against each other bit for bit, with timing.  python tools/verify_probe.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tscode_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble  # noqa: E402

vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
OLD = C.CDLL(os.path.join(ROOT, "tools", "probes", "libtsc_verify_perlane.so"))
for L in (OLD,):
    L.tsc_rmsd_verify.argtypes = [vp, i64, i32, vp, i32, f64, vp, vp, vp, i64, vp, i64, vp]


def run(L, pr, bits0):
    pr.sim_bits.copy_(bits0)
    pr.pair_list[0].zero_()
    pr.stats.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(L.tsc_rmsd_verify(ptr(pr.packed), pr.N, pr.M, ptr(pr.row_blocks), pr.n_rb, pr.thr, ptr(pr.sim_bits), ptr(pr.stats),
                            ptr(pr.pair_list), pr.pair_stride, ptr(pr.cand_list), pr.cand_stride, stream_ptr()), "verify")
    e1.record()
    torch.cuda.synchronize()
    return pr.sim_bits.clone(), pr.stats.cpu().numpy().copy(), int(pr.pair_list[0, 0]), e0.elapsed_time(e1)


cases = [(600, m) for m in (1, 5, 12, 19, 20, 21, 25, 31, 32, 33, 40, 63, 64, 65, 80, 96, 97, 128, 129, 150, 160, 161, 192)]
cases += [(20000, 40), (50000, 80), (20000, 150)]
for N, M in cases:
    S = gen_ensemble(7, N, M, max(2, N // 10), scale=np.array([5.0, 2.0, 1.0]))
    pr = RmsdPruner(torch.from_numpy(S).cuda(), np.full(M, 6), 0.5)
    pr.pack()
    pr.screen()
    torch.cuda.synchronize()
    bits0 = pr.sim_bits.clone()
    a = run(lib(), pr, bits0)
    b = run(OLD, pr, bits0)
    a = run(lib(), pr, bits0)
    b = run(OLD, pr, bits0)
    rows = pr.n_rb * 32
    diff = int((a[0][:rows] != b[0][:rows]).sum())
    print(f"N={N} M={M}: words differing {diff}  stats new {a[1].tolist()} old {b[1].tolist()}  pairs {a[2]} / {b[2]}  "
          f"ms new {a[3]:.3f} old {b[3]:.3f}  {'ok' if diff == 0 and a[2] == b[2] else 'BAD'}", flush=True)
