#!/usr/bin/env python
"""Per-call latency of the SCALAR drop-ins (one structure per call: upload + launch + read-back) next to the numba
originals from baseline/_ref on the same box — the measurement behind install_into's default of NOT patching them
(VERDICT r1 item 5 / ADVICE).  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ.setdefault("TSCODE_REFERENCE", os.path.join(ROOT, "baseline", "_ref"))

from tscode_b200 import numba_functions as nf, optimization_methods as om, rmsd_pruning as rp  # noqa: E402
from tscode_b200.synth import gen_ensemble, gen_poses, materialise_poses  # noqa: E402


def per_call(fn, n):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


frags, conf, R, t = gen_poses(0, 64, (50, 50))
pose = materialise_poses(frags, conf, R, t, [0])[0]
ids = np.array([50, 50])
S = gen_ensemble(0, 40, 80, 4)


class Mol:
    pass


mols = [Mol(), Mol()]
for k, m in enumerate(mols):
    m.atomcoords, m.rotation, m.position = frags[k], R[0, k], t[0, k]
cons, tg = [(3, 60), (7, 71)], [2.5, None]
res = {"unit": "microseconds per call"}
res["b200"] = {
    "compenetration_check (2 x 50 atoms)": per_call(lambda: nf.compenetration_check(pose, ids, 1.5, 0), 300),
    "get_embed (2 x 50 atoms)": per_call(lambda: nf.get_embed(mols, conf[0]), 300),
    "rmsd_and_max_numba (80 atoms)": per_call(lambda: rp.rmsd_and_max_numba(S[0], S[1]), 300),
    "_rmsd_similarity (vs 36 structures)": per_call(lambda: rp._rmsd_similarity(S[0], list(S[1:37]), 1.0), 300),
    "fitness_check (2 constraints)": per_call(lambda: om.fitness_check(pose, cons, tg, 5.0), 300),
    "compenetration_check_batch, 100 000 poses, per pose": per_call(
        lambda: nf.compenetration_check_batch(materialise_poses(*gen_poses(0, 100000, (50, 50))[0:4])[:100000], ids), 1) / 1e5
    if False else None,
}
try:
    import ref_harness
    ref_harness.install(full=True)
    from tscode.embeds import get_embed
    from tscode.numba_functions import compenetration_check
    from tscode.optimization_methods import fitness_check
    from tscode.rmsd_pruning import _rmsd_similarity, rmsd_and_max_numba
    res["numba_reference"] = {
        "compenetration_check (2 x 50 atoms)": per_call(lambda: compenetration_check(pose, ids, 1.5, 0), 2000),
        "get_embed (2 x 50 atoms)": per_call(lambda: get_embed(mols, conf[0]), 2000),
        "rmsd_and_max_numba (80 atoms)": per_call(lambda: rmsd_and_max_numba(S[0], S[1]), 2000),
        "_rmsd_similarity (vs 36 structures)": per_call(lambda: _rmsd_similarity(S[0], list(S[1:37]), 1.0), 300),
        "fitness_check (2 constraints)": per_call(lambda: fitness_check(pose, cons, tg, 5.0), 2000),
    }
except Exception as e:
    res["numba_reference"] = {"unavailable": repr(e)}
# the batched forms the patched loops use instead
import torch  # noqa: E402
P = 100000
fr, cf, Rr, tt = gen_poses(0, P, (50, 50))
Sp = materialise_poses(fr, cf, Rr, tt)
nf.compenetration_check_batch(Sp, ids)
t0 = time.perf_counter(); v = nf.compenetration_check_batch(Sp, ids); dt = time.perf_counter() - t0
res["batched"] = {"compenetration_check_batch: 100 000 materialised poses incl. H2D of 240 MB, us per pose": dt / P * 1e6,
                  "passes": int(v.sum())}
print(json.dumps(res, indent=1))
