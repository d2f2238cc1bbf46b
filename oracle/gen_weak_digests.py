#!/usr/bin/env python
"""Mask digests of the weak-scaling ensembles bench.py reports beside the strong-scaling headline (pair count
proportional to the number of GPUs: N = floor(50000 sqrt(w) / 128) * 128 conformers x 80 atoms, N/10 clusters).
TEST INFRASTRUCTURE: produced by the C oracle (oracle/oracle.c, pinned against the live reference at 10k / 20k / 50k),
not by the reference itself — labelled so in tests/golden/prune_masks_weak.json."""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_c  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

# usage: python oracle/gen_weak_digests.py [worlds, default "2 4 8"]; entries already in the file are kept
path = os.path.join(ROOT, "tests", "golden", "prune_masks_weak.json")
out = json.load(open(path)) if os.path.exists(path) else {}
worlds = [int(x) for x in sys.argv[1:]] or [2, 4, 8]
for w in worlds:
    if str(w) in out:
        continue
    N = int(50000 * math.sqrt(w) / 128) * 128
    S = gen_ensemble(3, N, 80, N // 10)
    t0 = time.perf_counter()
    m, ne, rounds = oracle_c.prune_heavy(S, 0.5)
    out[str(w)] = dict(world=w, N=N, M=80, n_clusters=N // 10, seed=3, thr=0.5, survivors=int(m.sum()), digest=mask_digest(m),
                       pairs_evaluated=int(ne), wall_s=round(time.perf_counter() - t0, 1), source="oracle/oracle.c (C port)")
    print(out[str(w)], flush=True)
    json.dump(out, open(path, "w"), indent=1)
