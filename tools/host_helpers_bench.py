#!/usr/bin/env python
"""Timings of the library's HOST helpers against the Python / numpy statements they replaced (no GPU needed):
work lists of the screen, centring, the rot_corr replay on a clustered synthetic first-hit array, XYZ text out and in.
Prints one JSON object; `python tools/host_helpers_bench.py > profiles/r02_host_helpers.json`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tscode_b200 import _host, torsion_module as tm  # noqa: E402
from tscode_b200.rmsd_pruning import _upload_bounds  # noqa: E402
from tscode_b200.utils import parse_xyz, xyz_text  # noqa: E402


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); ts.append((time.perf_counter() - t) * 1e3)
    return min(ts), r


out = {"host": {"cpus": os.cpu_count()}}

# work lists of one new ensemble size (the whole ensemble + the upload chunks)
N = 50000
rb = _host.owned_row_blocks(N, 0, 1)
bounds = _upload_bounds(N)
spans = [(0, None)] + [(bounds[c], bounds[c + 1]) for c in range(len(bounds) - 1)]
t_nat, a = best(lambda: [_host.build_screen_items(N, rb, 148, panel_lo=lo, panel_hi=hi, tile_j=48) for lo, hi in spans])
t_py, b = best(lambda: [_host.build_items_balanced(N, rb, 148, lo, hi, 2.5 * 32.0 / 48, tile_j=48) for lo, hi in spans], 2)
out["screen_work_lists_50000"] = {"lists": len(spans), "native_ms": t_nat, "python_ms": t_py,
                                  "equal": all(np.array_equal(x, y) for x, y in zip(a, b))}

# centring (torsion_module.py:1023)
S = np.random.default_rng(0).normal(size=(20000, 63, 3))
t_nat, c1 = best(lambda: tm.centre_structures(S))
t_py, c2 = best(lambda: S - S.mean(axis=1, keepdims=True))
out["centre_20000x63"] = {"native_ms": t_nat, "numpy_ms": t_py, "bit_identical": bool(np.array_equal(c1, c2))}

# replay of the grouping loop on a clustered first-hit array (the shape of BASELINE configs[3]: 84 clusters)
N, T = 20000, 5
rng = np.random.default_rng(0)
cl = rng.integers(0, 84, N)
first = np.full(N, N, dtype=np.int64)
last = {}
for i in range(N - 1, -1, -1):
    if cl[i] in last:
        first[i] = last[cl[i]]
    last[cl[i]] = i
ln = np.clip(np.minimum(first, N - 1) - np.arange(N), 0, None)
off = (np.cumsum(ln) - ln).astype(np.int64)
compact = np.zeros(int(ln.sum()), dtype=np.uint64)
for t in range(T):
    compact |= rng.integers(0, 3, size=compact.size).astype(np.uint64) << np.uint64(3 * t)
table = np.zeros((T, 6)); table[:, :3] = [0, 120, 240]


def lookup(i, js):
    cc = compact[off[i] + (np.asarray(js) - i - 1)]
    return np.stack([table[t][((cc >> np.uint64(3 * t)) & np.uint64(7)).astype(np.int64)] for t in range(T)], axis=-1)


lookup.T, lookup.compact, lookup.off, lookup.table = T, compact, off, table
tm.ladder_replay_scan(first.copy(), N, lookup, native=True)
t_nat, (m1, s1) = best(lambda: tm.ladder_replay_scan(first.copy(), N, lookup, native=True))
t_py, (m2, s2) = best(lambda: tm.ladder_replay_scan(first.copy(), N, lookup, native=False), 1)
out["rotcorr_replay_20000_5rotors"] = {"pairs_visited": int(ln.sum()), "survivors": int(m1.sum()), "native_ms": t_nat,
                                       "python_ms": t_py, "equal": bool(np.array_equal(m1, m2) and np.array_equal(s1, s2))}

# XYZ text out and in
S = np.random.default_rng(1).normal(size=(20000, 100, 3)) * 3
at = np.random.default_rng(2).choice([1, 6, 7, 8], size=100)
t_w, txt = best(lambda: xyz_text(S, at))
sym = ["X", "H", "", "", "", "", "C", "N", "O"]
t0 = time.perf_counter()
py = "".join("100\ntemp\n" + "".join('%s     % .6f % .6f % .6f\n' % (sym[at[i]], a[0], a[1], a[2]) for i, a in enumerate(s))
             for s in S[:2000])
t_wpy = (time.perf_counter() - t0) * 1e3 * 10
t_r, mol = best(lambda: parse_xyz(txt))


def python_reader(text):                      # the per-line split / float() loop a Python XYZ reader runs
    it = iter(text.splitlines())
    frames = []
    for line in it:
        n = int(line.split()[0]); next(it)
        frames.append([[float(x) for x in next(it).split()[1:4]] for _ in range(n)])
    return np.array(frames)


t0 = time.perf_counter(); ref = python_reader(xyz_text(S[:2000], at).decode()); t_rpy = (time.perf_counter() - t0) * 1e3 * 10
out["xyz_20000x100"] = {"bytes": len(txt), "write_native_ms": t_w, "write_python_ms_extrapolated_from_a_tenth": t_wpy,
                        "write_equal_on_the_tenth": txt[:len(py)] == py.encode(),
                        "read_native_ms": t_r, "read_python_ms_extrapolated_from_a_tenth": t_rpy,
                        "read_equal_on_the_tenth": bool(np.array_equal(ref, mol.atomcoords[:ref.shape[0]]))}
print(json.dumps(out, indent=1))
