#!/usr/bin/env python
"""Does the unmodified reference (baseline/_ref, numba) run on this box?  Times it on C3-derived samples."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n_thr = len(os.sched_getaffinity(0))
os.environ["NUMBA_NUM_THREADS"] = str(n_thr)
os.environ["OMP_NUM_THREADS"] = str(n_thr)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np  # noqa: E402

res = {"affinity": n_thr}
try:
    import numba
    from tscode.rmsd_pruning import prune_conformers_rmsd
    res["numba"] = numba.__version__
    res["numba_threads"] = numba.get_num_threads()
    res["threading_layer_pref"] = numba.config.THREADING_LAYER
    from tscode_b200.synth import gen_ensemble, mask_digest
    S = gen_ensemble(0, 1000, 40, 100)
    t0 = time.perf_counter(); _, m = prune_conformers_rmsd(S, np.full(40, 6), 0.5); res["c1_first_call_s"] = round(time.perf_counter() - t0, 2)
    t0 = time.perf_counter(); _, m = prune_conformers_rmsd(S, np.full(40, 6), 0.5); res["c1_warm_s"] = round(time.perf_counter() - t0, 3)
    res["c1_digest_ok"] = mask_digest(m) == "94e7a6f4441ed28b"
    try:
        res["threading_layer"] = numba.threading_layer()
    except Exception as e:
        res["threading_layer"] = repr(e)
    full = gen_ensemble(3, 50000, 80, 5000)
    a80 = np.full(80, 6)
    for n in (4000, 10000) + ((50000,) if "--full" in sys.argv else ()):
        t0 = time.perf_counter(); _, m = prune_conformers_rmsd(full[:n], a80, 0.5); dt = time.perf_counter() - t0
        res[f"first_{n}_s"] = round(dt, 2); res[f"first_{n}_survivors"] = int(m.sum())
        if n == 50000:
            res["full_digest_ok"] = mask_digest(m) == "478bc29df1e239da"
except Exception as e:
    res["error"] = repr(e)
print(json.dumps(res))
