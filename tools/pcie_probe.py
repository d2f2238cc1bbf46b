#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth on this box (the e2e roofline's denominator)."""
import json
import sys
import time

import torch

dev = torch.device("cuda:0")
res = {}
for mb in (1, 12, 96, 512):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[f"{name}_{mb}MB_GBps"] = round(n / best / 1e6, 2)
    # the same 96 MB in 8 chunks on a side stream (what the pipelined upload does)
    if mb == 96:
        s = torch.cuda.Stream()
        best = 1e9
        for _ in range(10):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.cuda.stream(s):
                for c in range(8):
                    lo, hi = n * c // 8, n * (c + 1) // 8
                    d[lo:hi].copy_(h[lo:hi], non_blocking=True)
            s.synchronize()
            best = min(best, time.perf_counter() - t0)
        res["h2d_96MB_8chunks_wall_GBps"] = round(n / best / 1e9, 2)
    # pageable source
    if mb == 96:
        hp = torch.empty(n, dtype=torch.uint8); hp.fill_(1)
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hp); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        res["h2d_96MB_pageable_GBps"] = round(n / best / 1e9, 2)
print(json.dumps(res))
