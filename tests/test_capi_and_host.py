"""CPU-only checks: the C-ABI library builds, loads and exports exactly what include/*.h
declares (no compute calls); host-side bookkeeping (tiles, sharding, ladder, thresholds)."""
import ctypes
import math
import os
import re
import subprocess

import numpy as np
import pytest

from tscode_b200 import _host, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "tscode_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsc_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_header_symbols():
    from tscode_b200.csrc import build
    so = build.build()
    assert os.path.exists(so)
    L = ctypes.CDLL(so)
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tscode_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert _lib.lib().tsc_version() == 100
    # geometry helpers are host-only and must agree with the numpy mirror
    for N in (1, 31, 32, 33, 64, 1000, 50000):
        assert _lib.lib().tsc_num_blocks_padded(N) == _host.num_blocks_padded(N)
        for M in (1, 20, 21, 80):
            assert _lib.lib().tsc_packed_doubles(N, M) == _host.packed_doubles(N, M)


def test_sass_uses_dmma_and_bulk_tma():
    """The hot kernel must really be on the FP64 tensor pipe and stage through bulk TMA."""
    from tscode_b200.csrc import build
    so = build.build()
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "DMMA.8x8x4" in sass
    assert "UBLKCP" in sass
    assert "DFMA" in sass


def test_sass_of_the_default_screen_uses_tcgen05_tmem_and_bulk_copies():
    """The default screen is a tcgen05 kernel: UTCHMMA (tcgen05.mma kind::f16), TMEM loads / stores (LDTM / STTM),
    tcgen05.commit (UTCBAR), bulk copies (UBLKCP), packed FP32 arithmetic (FFMA2) and the staged column terms
    (LDGSTS = cp.async) — for every instantiated form."""
    from tscode_b200.csrc import build
    so = build.build()
    for mode, J in ((0, 48), (1, 32), (2, 32), (0, 64)):
        fn = f"_ZN3tsc18rmsd_screen_kernelILi{mode}ELi{J}EEEvNS_8ScParamsE"
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, so], capture_output=True, text=True).stdout
        for mnemonic in ("UTCHMMA", "LDTM.x8", "STTM.x4", "UTCBAR", "UBLKCP", "FFMA2"):
            assert mnemonic in sass, (mode, J, mnemonic)
        if mode == 0:
            assert "LDGSTS" in sass
        if mode != 0:
            assert "FMNMX3.NAN" in sass and "MUFU.SQRT" in sass        # the quartic stage's lean decision and root


def test_product_has_no_oracle_or_cpu_fallback():
    for root, _, files in os.walk(os.path.join(ROOT, "tscode_b200")):
        for f in files:
            if f.endswith(".py"):
                s = open(os.path.join(root, f)).read()
                assert "import oracle" not in s and "from oracle" not in s, f
    import torch
    if not torch.cuda.is_available():
        from tscode_b200.rmsd_pruning import prune_conformers_rmsd
        from tscode_b200.numba_functions import compenetration_check
        with pytest.raises(RuntimeError):
            prune_conformers_rmsd(np.zeros((4, 3, 3)), np.full(3, 6))
        with pytest.raises(RuntimeError):
            compenetration_check(np.zeros((4, 3)), np.array([2, 2]))


@pytest.mark.parametrize("N", [1, 2, 31, 32, 33, 64, 65, 1000, 1037, 4096])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_tiles_cover_upper_triangle(N, world):
    nb, nbp = _host.num_blocks(N), _host.num_blocks_padded(N)
    covered = np.zeros((nb, nbp), bool)
    total = 0
    for rank in range(world):
        rb = _host.owned_row_blocks(N, rank, world)
        tiles = _host.build_tiles(N, rb)
        total += len(tiles)
        assert np.all(tiles[:, 2] < max(len(rb), 1))
        for ib, jp, lb, _ in tiles:
            assert rb[lb] == ib
            assert not covered[ib, 2 * jp] and not covered[ib, 2 * jp + 1]
            covered[ib, 2 * jp] = covered[ib, 2 * jp + 1] = True
    for ib in range(nb):
        assert covered[ib, ib:].all()            # every word >= ib of every row block is written
    assert total == sum(nbp // 2 - ib // 2 for ib in range(nb))
    # work items of the tcgen05 screen, for both tile widths: every j tile right of a panel's first row exactly once;
    # one contiguous stretch per CTA, dealt round-robin; empty items only at the end of a CTA's list (the kernel stops
    # at the first one)
    for tile_j in (32, 48, 64):
        first = lambda p: (128 * p) // tile_j                 # noqa: E731  the tile that holds the panel's first column
        njt = (((N + 127) // 128) * 128 + tile_j - 1) // tile_j
        for n_ctas in (148, 5):
            seen = set()
            for rank in range(world):
                rb = _host.owned_row_blocks(N, rank, world)
                items = _host.build_screen_items(N, rb, n_ctas, tile_j=tile_j)
                g = min(n_ctas, len(items))
                for b in range(g):
                    mine = items[b::g]
                    live = mine[:, 2] > 0
                    assert not live[np.argmin(live):].any() or live.all()
                for p, j0, cnt, lb in items:
                    assert rb[lb] == 4 * p and cnt >= 0
                    for jt in range(j0, j0 + cnt):
                        assert (p, jt) not in seen
                        seen.add((p, jt))
            assert seen == {(p, jt) for p in range((N + 127) // 128) for jt in range(first(p), njt)}
        # ... and restricted to panel ranges (the sub-launches of the pipelined upload) the pieces tile the whole
        seen = set()
        rb = _host.owned_row_blocks(N, 0, 1)
        n_panels = (N + 127) // 128
        cuts = sorted({0, n_panels // 3, n_panels // 2, n_panels})
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            for p, j0, cnt, lb in _host.build_screen_items(N, rb, 148, panel_lo=lo, panel_hi=hi, tile_j=tile_j):
                assert lo <= p < hi or cnt == 0
                for jt in range(j0, j0 + cnt):
                    assert (p, jt) not in seen
                    seen.add((p, jt))
        assert seen == {(p, jt) for p in range(n_panels) for jt in range(first(p), njt)}


def test_row_sharding_is_balanced():
    N = 50000
    loads = [len(_host.build_tiles(N, _host.owned_row_blocks(N, r, 8))) for r in range(8)]
    assert max(loads) / min(loads) < 1.005
    items = [_host.build_screen_items(N, _host.owned_row_blocks(N, r, 8), 148)[:, 2].sum() for r in range(8)]
    assert max(items) / min(items) < 1.005
    # inside a rank: cost (tiles + 2.5 per item) per CTA of the persistent grid, full launch and upload sub-launches
    def cta_loads(items, g=148, cost=2.5):
        return np.array([items[b::g, 2].sum() + cost * (items[b::g, 2] > 0).sum() for b in range(g)])
    for world in (1, 8):
        rb = _host.owned_row_blocks(N, 0, world)
        for tile_j in (32, 48, 64):
            loads = cta_loads(_host.build_screen_items(N, rb, 148, tile_j=tile_j), cost=2.5 * 32 / tile_j)
            assert loads.max() / loads.mean() < 1.02, (world, tile_j, loads.max() / loads.mean())
    rb = _host.owned_row_blocks(N, 0, 1)
    n_panels = (N + 127) // 128
    for c in range(8):
        loads = cta_loads(_host.build_screen_items(N, rb, 148, panel_lo=n_panels * c // 8, panel_hi=n_panels * (c + 1) // 8))
        assert loads.max() / loads.mean() < (1.06 if c < 6 else 1.25), (c, loads.max() / loads.mean())   # (the last chunks are tiny)


def test_ladder_schedule_matches_reference_rule():
    # rmsd_pruning.py:186-192 with a mask that never shrinks
    assert _host.run_ladder(1000, lambda k, cs: 1000) == [20, 10, 5, 2, 1]
    assert _host.run_ladder(777, lambda k, cs: 777) == [20, 10, 5, 2, 1]
    assert _host.run_ladder(1037, lambda k, cs: 1037) == [50, 20, 10, 5, 2, 1]
    assert _host.run_ladder(50000, lambda k, cs: 50000)[0] == 2000
    assert _host.run_ladder(1, lambda k, cs: 1) == [1]
    seen = []
    _host.run_ladder(1037, lambda k, cs: seen.append((k, cs)) or 1037)
    assert seen[0] == (50, 20) and seen[-1] == (1, 1037)
    # a shrinking mask skips rounds
    it = iter([100, 100, 100, 100])
    assert _host.run_ladder(1000, lambda k, cs: next(it)) == [20, 2, 1]


@pytest.mark.parametrize("thr", [1.5, 0.5, 1.4, 2.0, 1.7, 1e-3, 3.3333333333333335, 0.1])
def test_sqrt_threshold_image(thr):
    t2 = _host.sqrt_threshold_image(thr)
    assert math.sqrt(t2) >= thr and math.sqrt(math.nextafter(t2, -math.inf)) < thr
    rng = np.random.default_rng(0)
    x = thr * thr * (1 + (rng.random(20000) - 0.5) * 1e-14)
    assert np.array_equal(np.sqrt(x) < thr, x < t2)


@pytest.mark.parametrize("N,density,seed", [(64, 0.05, 1), (777, 0.002, 2), (1037, 0.01, 3), (2500, 0.0005, 4),
                                            (45, 0.3, 6), (300, 0.0, 7), (33, 1.0, 8), (1, 0.0, 9), (2, 1.0, 10)])
def test_pairlist_ladder_model_matches_oracle(N, density, seed):
    """The pair-list formulation the fused ladder kernel implements (first similar active partner per
    row by atomicMin, cache bitmap scanned up to it) is the reference ladder."""
    from oracle import oracle_c
    rng = np.random.default_rng(seed)
    sim = np.triu(rng.random((N, N)) < density, 1)
    ref, _, rounds = oracle_c.prune_heavy(np.zeros((N, 1, 3)), 0.5, sim_bytes=sim.astype(np.uint8))
    pairs = np.argwhere(sim)
    rng.shuffle(pairs)
    mask, ran = _host.ladder_pairlist_model(pairs, N)
    assert ran == [int(k) for k in rounds]
    assert np.array_equal(mask, ref)


@pytest.mark.parametrize("N,dens,seed", [(60, 0.05, 1), (300, 0.01, 2), (300, 0.2, 3), (777, 0.003, 4), (100, 0.0, 6),
                                         (50, 1.0, 7)])
def test_rotcorr_scan_replay_equals_dense_replay(N, dens, seed):
    """The grouping loop of prune_conformers_rmsd_rot_corr driven by first hits (what the forward-scan
    kernel returns) gives the same mask and the same rotor states as the replay on a full matrix —
    both in its Python form and with the per-chunk inner loop in C (tsc_host_rotcorr_chunk)."""
    from tscode_b200.torsion_module import ladder_replay, ladder_replay_scan
    rng = np.random.default_rng(seed)
    sim = np.triu(rng.random((N, N)) < dens, 1)
    T = 3
    table = np.zeros((T, 6)); table[:, :3] = [0, 120, 240]
    codes = rng.integers(0, 3, size=(N, N, T))
    ang = table[np.arange(T)[None, None, :], codes]
    m1, s1 = ladder_replay(sim, N, ang)
    first = np.array([(np.flatnonzero(sim[i])[0] if sim[i].any() else N) for i in range(N)])
    lengths = np.maximum(np.minimum(first, N - 1) - np.arange(N), 0)
    off = np.cumsum(lengths) - lengths
    packed = (codes[..., 0] | (codes[..., 1] << 3) | (codes[..., 2] << 6)).astype(np.uint32)
    compact = np.concatenate([packed[i, i + 1:i + 1 + lengths[i]] for i in range(N)]) if lengths.sum() else np.zeros(0, np.uint32)

    def lookup(i, js):
        cc = compact[off[i] + (np.asarray(js) - i - 1)]
        return np.stack([table[t][(cc >> (3 * t)) & 7] for t in range(T)], axis=-1)
    lookup.T, lookup.compact, lookup.off, lookup.table = T, compact, off.astype(np.int64), table
    for native in (False, True):
        m2, s2 = ladder_replay_scan(first, N, lookup, native=native)
        assert np.array_equal(m1, m2) and np.array_equal(s1, s2), native


@pytest.mark.parametrize("N,n_clusters,seed", [(2500, 12, 0), (1200, 200, 1), (3000, 1, 2)])
def test_native_replay_on_clustered_first_hits(N, n_clusters, seed):
    """The shape BASELINE configs[3] has — a few clusters, every row's first hit = the next member of its cluster —
    with 2-, 3- and 6-fold rotors: the native replay (tsc_host_ladder_replay: per-row image tables for rows with many
    visits, direct sums for short ones, the survivor choice on re-used set models) against the Python loop, mask and
    rotor states bit for bit; several ladder rounds run (N > 5 k for k up to 500 / 200)."""
    from tscode_b200.torsion_module import ladder_replay_scan
    rng = np.random.default_rng(seed)
    cl = rng.integers(0, n_clusters, N)
    first = np.full(N, N, dtype=np.int64)
    last = {}
    for i in range(N - 1, -1, -1):
        if cl[i] in last:
            first[i] = last[cl[i]]
        last[cl[i]] = i
    n_ang = [2, 3, 6, 3, 1]
    T = len(n_ang)
    table = np.zeros((T, 6))
    for t, n in enumerate(n_ang):
        table[t, :n] = np.arange(n) * (360.0 / n)
    lengths = np.maximum(np.minimum(first, N - 1) - np.arange(N), 0)
    off = (np.cumsum(lengths) - lengths).astype(np.int64)
    compact = np.zeros(max(int(lengths.sum()), 1), dtype=np.uint64)
    for t, n in enumerate(n_ang):
        compact[:int(lengths.sum())] |= rng.integers(0, n, size=int(lengths.sum())).astype(np.uint64) << np.uint64(3 * t)

    def lookup(i, js):
        cc = compact[off[i] + (np.asarray(js) - i - 1)]
        return np.stack([table[t][((cc >> np.uint64(3 * t)) & np.uint64(7)).astype(np.int64)] for t in range(T)], axis=-1)
    lookup.T, lookup.compact, lookup.off, lookup.table = T, compact, off, table
    m_py, s_py = ladder_replay_scan(first.copy(), N, lookup, native=False)
    m_c, s_c = ladder_replay_scan(first.copy(), N, lookup, native=True)
    assert np.array_equal(m_py, m_c) and np.array_equal(s_py, s_c)
    assert 0 < m_c.sum() < N and (s_c >= 0).all() and (s_c < 360).all() and len(np.unique(s_c)) > 2


def test_native_xyz_text_is_byte_identical_to_reference_formatting():
    """(f)-4: tsc_host_write_xyz against the reference's per-atom '%s     % .6f % .6f % .6f\\n' (utils.py:114-126),
    including half-way decimals, negative zero, huge values, inf / nan, and per-frame titles."""
    import io
    from tscode_b200.utils import _SYMBOLS, write_xyz, xyz_text

    def ref_write(coords, atomnos, title='temp'):                 # utils.py:114-126 verbatim (pt -> symbol table)
        string = ''
        string += str(len(coords))
        string += f'\n{title}\n'
        for i, atom in enumerate(coords):
            string += '%s     % .6f % .6f % .6f\n' % (_SYMBOLS[atomnos[i]], atom[0], atom[1], atom[2])
        return string
    rng = np.random.default_rng(0)
    S = rng.normal(size=(300, 37, 3)) * np.array([1, 10, 1000])
    S[0, 0] = [0.0, -0.0, -1e-9]; S[0, 1] = [0.5e-6, 1.5e-6, 2.5e-6]; S[0, 2] = [123456.7890125, -0.0000005, 1e15]
    S[0, 3] = [np.inf, -np.inf, np.nan]; S[0, 4] = [1e300, -1e300, 1e-300]
    S[1, :, 0] = np.round(S[1, :, 0], 6) + 0.5e-6
    at = rng.choice([1, 6, 7, 8, 17, 35], size=37)
    want = ''.join(ref_write(s, at, f"conf {k}") for k, s in enumerate(S)).encode()
    for nt in (1, 3, 16):
        assert xyz_text(S, at, [f"conf {k}" for k in range(len(S))], n_threads=nt) == want
    assert xyz_text(S[:2], at) == ''.join(ref_write(s, at) for s in S[:2]).encode()
    buf = io.StringIO()
    write_xyz(S[5], at, buf, title='x')
    assert buf.getvalue() == ref_write(S[5], at, 'x')


def _deal_like_the_kernel(items, n_ctas):
    """rmsd_ts_kernel's item loop: grid = min(n_ctas, n_items) CTAs, CTA b visits entries b, b + grid, ... and stops
    at the first empty one.  Returns how often every (panel, j tile) pair is processed."""
    n = len(items)
    g = min(n_ctas, n)
    seen = {}
    for b in range(g):
        for it in range(b, n, g):
            p, j0, cnt, lb = items[it]
            if cnt == 0:
                break
            for jt in range(j0, j0 + cnt):
                seen[(p, jt)] = seen.get((p, jt), 0) + 1
    return seen


@pytest.mark.parametrize("N", [1, 129, 650, 900, 1500, 2000, 2500, 4096, 4100, 5000, 8191, 12345])
def test_balanced_items_dealt_like_the_kernel_cover_every_tile_once(N):
    """The balanced work list relies on the kernel's round-robin dealing AND on 'an empty entry ends the CTA's list':
    simulate exactly that for whole launches and for the sub-launches of the pipelined upload, several ranks and
    grid sizes (incl. grids larger than the number of stretches: the layout stride must then be the grid)."""
    from tscode_b200.rmsd_pruning import _upload_bounds
    n_panels = (N + 127) // 128
    pb = _upload_bounds(N)
    ranges = [(0, None)] + [(pb[c], pb[c + 1]) for c in range(len(pb) - 1)]
    for world in (1, 2, 8):
        for rank in range(min(world, 2)):
            rb = _host.owned_row_blocks(N, rank, world)
            own = [int(ib // 4) for ib in rb if ib % 4 == 0]
            for n_ctas in (148, 132, 5):
                for lo, hi in ranges:
                    for tile_j in (32, 48, 64):
                        njt = (n_panels * 128 + tile_j - 1) // tile_j
                        h = n_panels if hi is None else hi
                        want = {(p, jt) for p in own if lo <= p < h for jt in range((128 * p) // tile_j, njt)}
                        items = _host.build_screen_items(N, rb, n_ctas, panel_lo=lo, panel_hi=hi, tile_j=tile_j)
                        seen = _deal_like_the_kernel(items, n_ctas)
                        assert set(seen) == want and all(v == 1 for v in seen.values()), (N, world, rank, n_ctas, lo, hi, tile_j)


def test_cluster_survivor_choice_equals_networkx():
    """torsion_module._cluster_rejects_fast restates how the reference picks the survivor of every cluster — nx.Graph(set
    of matches) -> connected_components -> tuple(subgraph.nodes)[0] (torsion_module.py:1136-1152, numba_functions.py,
    optimization_methods.py:341-355) — with plain dicts and sets; it must give networkx's answer for sets (whose
    iteration order decides), lists and sorted lists of edges, trees, cliques and long chains alike."""
    import random
    from tscode_b200 import torsion_module as tm
    rnd = random.Random(7)
    n_checked = 0
    for trial in range(1500):
        m = rnd.choice((2, 3, 5, 8, 9, 16, 33, 100, 700, 5000))
        kind = trial % 4
        es = set()
        if kind == 0:                                   # random sparse
            for _ in range(rnd.randrange(1, 2 * m)):
                a, b = rnd.randrange(m), rnd.randrange(m)
                if a != b:
                    es.add((min(a, b), max(a, b)))
        elif kind == 1:                                 # first-hit structure: every row points to one later row
            for a in range(m - 1):
                if rnd.random() < 0.6:
                    es.add((a, rnd.randrange(a + 1, m)))
        elif kind == 2:                                 # chains
            start = rnd.randrange(m)
            for a in range(start, min(m - 1, start + rnd.randrange(1, 40))):
                es.add((a, a + 1))
        else:                                           # a few dense clusters
            for _ in range(3):
                nodes = rnd.sample(range(m), min(m, rnd.randrange(2, 7)))
                for x in nodes:
                    for y in nodes:
                        if x < y:
                            es.add((x, y))
        if not es:
            continue
        for E in (es, list(es), sorted(es), sorted(es, reverse=True)):
            assert sorted(tm._cluster_rejects_fast(E)) == sorted(tm._cluster_rejects_nx(E)), (trial, kind)
            n_checked += 1
    assert n_checked > 3000
    assert tm._cluster_rejects([(0, 1), (1, 2), (5, 6)]) is not None and tm._CLUSTER_IMPL[0] is tm._cluster_rejects_fast


def test_screen_frame_is_orthogonal_and_weights_satisfy_the_inequality():
    """_host.screen_frame: whatever the first structure looks like, Q must be orthogonal to the library's 1e-13 and
    sum_b 1 / (3 t_b^2) <= 1 — the two facts the weighted Samuelson bound rests on (rmsd_screen.cu, ScFrame); an
    isotropic blob gives the identity exactly; the weighted bound never falls below lambda_max."""
    rng = np.random.default_rng(0)
    shapes = ([3, 3, 3], [6, 2, 1], [4, 4, 0.5], [8, 1, 1], [5, 5, 1e-6], [1e-3, 1, 1e3], [10, 0, 0])
    for sc in shapes:
        for M in (3, 17, 80):
            X = rng.normal(size=(M, 3)) * np.array(sc, dtype=np.float64)
            th = rng.normal(size=3)
            R, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            X = X @ R.T + 0.0 * th
            fr, ratio = _host.screen_frame(X)
            Q, t = fr[:9].reshape(3, 3), fr[9:]
            assert np.abs(Q @ Q.T - np.eye(3)).max() <= 1e-13
            assert (1.0 / (3.0 * t * t)).sum() <= 1.0 and (t > 1e-3).all() and (t < 1e3).all()
            # the bound itself: for random pairs of perturbed copies, sqrt(3) ||Q S Q^T diag(t)||_F >= sum of singular values
            for _ in range(20):
                P = X + rng.normal(size=X.shape) * 0.7
                Y = X + rng.normal(size=X.shape) * 0.7
                S = (P @ Q.T).T @ ((Y @ Q.T) * t)
                plain = P.T @ Y
                assert np.sqrt(3.0) * np.linalg.norm(S) >= np.linalg.svd(plain, compute_uv=False).sum() * (1 - 1e-12)
    fr, ratio = _host.screen_frame(np.eye(3) * 2.0)
    assert np.array_equal(fr, np.concatenate([np.eye(3).ravel(), np.ones(3)])) and abs(ratio - 1.0) < 1e-12
    assert _host.screen_mode_for(rng.normal(size=(80, 3)) * 3.0) == 0
    assert _host.screen_mode_for(rng.normal(size=(80, 3)) * np.array([6.0, 2.0, 1.0])) == 1


def test_library_rejects_a_frame_that_would_break_the_bound():
    """tsc_pack_screen / tsc_rmsd_screen check the frame before anything touches the GPU: a non-orthogonal Q or
    weights with sum 1 / w > 1 return cudaErrorInvalidValue (1)."""
    L = _lib.lib()
    good = np.concatenate([np.eye(3).ravel(), np.ones(3)])
    bad_q = good.copy(); bad_q[1] = 1e-6
    bad_t = good.copy(); bad_t[9:] = [0.9, 1.0, 1.0]
    nan_t = good.copy(); nan_t[10] = np.nan
    for fr in (bad_q, bad_t, nan_t):
        assert L.tsc_pack_screen(None, 10, 4, None, 4, None, None, None, None, None, None, 0, 0, 32, fr.ctypes.data, None) == 1
        assert L.tsc_rmsd_screen(None, None, None, None, None, None, 10, 4, None, 1, 0.5, None, None, 0, 0, 0, 0,
                                 fr.ctypes.data, None) == 1


def test_native_screen_plan_agrees_with_its_numpy_statement():
    """capi.cu: tsc_host_screen_plan against _host.screen_plan (numpy) — same undecided fraction, same mode, and a
    frame the library accepts; the sampled pairs are the same on every call and never pair a structure with itself."""
    from tscode_b200.synth import gen_ensemble
    rng = np.random.default_rng(1)
    for scale, M, N in ((3.0, 80, 3000), (np.array([6.0, 2.0, 1.0]), 80, 3000), (np.array([4.0, 4.0, 0.5]), 30, 2000),
                        (3.0, 12, 500)):
        S = gen_ensemble(5, N, M, max(2, N // 10), scale=scale)
        A = M + 3
        full = np.zeros((N, A, 3))
        heavy = np.sort(rng.choice(A, size=M, replace=False)).astype(np.int32)
        full[:, heavy] = S
        full[:, np.setdiff1d(np.arange(A), heavy)] = rng.normal(size=(N, A - M, 3)) * 50.0     # "hydrogens": ignored
        pi, pj = _host.sample_pair_indices(N)
        pi2, pj2 = _host.sample_pair_indices(N)
        assert np.array_equal(pi, pi2) and np.array_equal(pj, pj2) and (pi != pj).all()
        assert pi.min() >= 0 and pj.min() >= 0 and pi.max() < N and pj.max() < N and pi.size == _host.SAMPLE_PAIRS
        fr, mode, und = _host.screen_plan_native(full, heavy, 0.5, 0, pi, pj)
        fr2, mode2, und2 = _host.screen_plan(S[0], S[pi], S[pj], 0.5)
        assert mode == mode2 and abs(und - und2) < 1e-12, (mode, mode2, und, und2)
        Q, t = fr[:9].reshape(3, 3), fr[9:]
        assert np.abs(Q @ Q.T - np.eye(3)).max() <= 1e-13 and (1.0 / (3.0 * t * t)).sum() <= 1.0
        assert np.allclose(np.sort(t), np.sort(fr2[9:]), rtol=1e-9)
    assert _host.sample_pair_indices(1)[0].size == 0
    assert _host.plan_mode(1.0, 0.0) == 0 and _host.plan_mode(1.5, 0.0) == 1 and _host.plan_mode(1.0, 0.02) == 1
    assert _host.plan_mode(1.0, 0.5) == 2


def test_native_xyz_reader_equals_the_cclib_reader_model():
    """(f)-4, input side: tsc_host_read_xyz / utils.parse_xyz against the restated algorithm of the XYZ reader behind
    the reference's read_xyz (oracle_np.xyz_reader_model; cclib itself is absent: parity against it is unpinned) — on
    the reference's own output format (write_xyz), on free-form files (tabs, extra columns, blank lines between frames,
    CRLF and bare-CR line ends, no final newline, exponents, signs, inf / nan), an incomplete last frame, and the
    malformed cases the reader rejects; numbers bit-identical to Python's float()."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle_np
    from tscode_b200.utils import _SYMBOLS, parse_xyz, read_xyz, xyz_text

    def check(text, n_threads=None):
        want_c, want_sym, want_com = oracle_np.xyz_reader_model(text)
        got = parse_xyz(text, n_threads=n_threads)
        n = want_c.shape[0]
        assert got.atomcoords.shape[0] == n
        if n:
            assert got.atomcoords.shape == want_c.shape
            assert np.array_equal(got.atomcoords.view(np.uint64), want_c.view(np.uint64))      # bit for bit, -0.0 and nan too
            assert [_SYMBOLS[z] for z in got.atomnos] == want_sym
            assert got.metadata["comments"] == want_com[:n]
        return got

    rng = np.random.default_rng(5)
    S = rng.normal(size=(300, 37, 3)) * np.array([1, 10, 1000])
    S[0, 0] = [0.0, -0.0, -1e-9]; S[0, 1] = [0.5e-6, 1.5e-6, 2.5e-6]; S[0, 2] = [123456.7890125, -0.0000005, 1e15]
    S[0, 3] = [np.inf, -np.inf, np.nan]; S[0, 4] = [1e300, -1e300, 1e-300]
    at = rng.choice([1, 6, 7, 8, 17, 35], size=37)
    text = xyz_text(S, at, [f"conf {k}  E = {k * 0.1:.4f}" for k in range(len(S))]).decode()
    for nt in (1, 5):
        got = check(text, nt)
    assert np.array_equal(got.atomnos, at) and got.natom == 37
    fin = np.isfinite(S) & (np.abs(S) < 1e6)
    assert np.abs(got.atomcoords[fin] - S[fin]).max() <= 0.50001e-6             # what '% .6f' keeps
    big = xyz_text(rng.normal(size=(3000, 12, 3)), np.full(12, 6)).decode()       # enough frames for several threads
    check(big, 7)
    free = ("\n  3 atoms\nwater  # comment with 4 tokens 1 2\nO\t0.0 0 1e-3 extra columns ignored\n"
            "H   +0.757   .586  -0.0\nH  -7.57E-1 5.86e+00 1.\n"
            "\n3\n\nO 1 2 3\nH 4 5 6\nH 123456789012345678901234567890e-25 0.1000000000000000055511151231257827 4.9e-324\n"
            "3\nthird\nO inf -Infinity nan\nH 1e400 -1e400 1e-400\nCl 2.2250738585072011e-308 17976931348623157e292 9007199254740993")
    got = check(free)
    assert got.atomcoords.shape == (3, 3, 3) and list(got.atomnos) == [8, 1, 17]
    check(free.replace("\n", "\r\n"))
    check(free.replace("\n", "\r"))
    check(free + "\n")
    check(free + "\n\n3\nincomplete frame: dropped\nO 0 0 0\nH 0 0 1")          # the text ends inside a frame
    check(free + "\n3")                                                           # ... or right after a count line
    check("")
    check("\n")
    assert parse_xyz(b"1\nx\nH 0 0 0\n").atomcoords.shape == (1, 1, 3)
    # random decimal strings: every digit count, exponents beyond the exact fast path, leading / trailing zeros
    import random
    rnd = random.Random(9)
    toks = []
    for _ in range(3000):
        nd = rnd.randrange(1, 25)
        digits = "".join(rnd.choice("0123456789") for _ in range(nd))
        cut = rnd.randrange(0, nd + 1)
        t = rnd.choice(["", "-", "+"]) + digits[:cut] + rnd.choice([".", ""] if cut == nd and cut else ["."]) + digits[cut:]
        if t.strip("+-") in (".", ""):
            t += "0"
        if rnd.random() < 0.5:
            t += rnd.choice("eE") + rnd.choice(["", "-", "+"]) + str(rnd.randrange(0, 330))
        toks.append(t)
    body = "".join(f"C {toks[3 * a]} {toks[3 * a + 1]} {toks[3 * a + 2]}\n" for a in range(1000))
    check("1000\nfuzz\n" + body)
    for bad, exc in [("x\nc\nH 0 0 0\n", ValueError), ("-1\nc\n", ValueError), ("2\nc\nH 0 0 0\nH 0 0\n", ValueError),
                     ("1\nc\nH 0 0 zero\n", ValueError), ("1\nc\nH 0 0 1e\n", ValueError), ("1\nc\nH 0 0 0x10\n", ValueError),
                     ("1\nc\nH 0 0 0\n2\nc\nH 0 0 0\nH 0 0 1\n", ValueError), ("1\nc\nQq 0 0 0\n", KeyError)]:
        with pytest.raises(exc):
            parse_xyz(bad)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "ens.xyz")
        with open(fn, "w") as f:
            f.write(text)
        mol = read_xyz(fn)
        assert np.array_equal(mol.atomcoords.view(np.uint64), got_bits(text, oracle_np)) and np.array_equal(mol.atomnos, at)
        with open(fn, "w") as f:
            f.write("")
        with pytest.raises(AssertionError):
            read_xyz(fn)


def got_bits(text, oracle_np):
    return oracle_np.xyz_reader_model(text)[0].view(np.uint64)


def test_native_screen_items_equal_the_python_rule():
    """capi.cu: tsc_host_screen_items (what _host.build_screen_items calls) against the rule as _host.build_items_balanced
    states it, entry by entry: every tile width, whole ensembles and upload chunks, 1 / 2 / 8 ranks, grids of 1 / 16 /
    148 CTAs, sizes from 1 structure to 2e5; and through the raw entry point the forms build_screen_items does not use
    (tiles_per_panel layout, max_item cutting, a too small output buffer, bad arguments)."""
    from tscode_b200.rmsd_pruning import _upload_bounds
    L = _lib.lib()
    n_cases = 0
    for N in list(range(1, 700, 61)) + [1000, 4097, 10000, 50000, 75892, 200000]:
        for world in (1, 2, 8):
            for rank in sorted({0, world - 1}):
                rb = _host.owned_row_blocks(N, rank, world)
                b = _upload_bounds(N)
                for tile_j in (32, 48, 64):
                    for n_ctas in (148, 16, 1):
                        for lo, hi in [(0, None)] + [(b[c], b[c + 1]) for c in range(len(b) - 1)]:
                            got = _host.build_screen_items(N, rb, n_ctas, panel_lo=lo, panel_hi=hi, tile_j=tile_j)
                            want = _host.build_items_balanced(N, rb, n_ctas, lo, hi, 2.5 * 32.0 / tile_j, tile_j=tile_j)
                            assert got.dtype == np.int32 and got.shape == want.shape and np.array_equal(got, want), \
                                (N, world, rank, tile_j, n_ctas, lo, hi)
                            n_cases += 1
    assert n_cases > 3000

    def raw(N, rb, n_ctas, lo, hi, cost, max_item, tpp, tile_j, cap=None):
        rb = np.ascontiguousarray(rb, dtype=np.int32)
        n = int(L.tsc_host_screen_items(N, rb.ctypes.data, rb.size, n_ctas, lo, hi, cost, max_item, tpp, tile_j, None, 0))
        out = np.full((max(n if cap is None else cap, 1), 4), -7, dtype=np.int32)
        n2 = int(L.tsc_host_screen_items(N, rb.ctypes.data, rb.size, n_ctas, lo, hi, cost, max_item, tpp, tile_j,
                                         out.ctypes.data, n if cap is None else cap))
        assert n2 == n
        return n, out
    for N in (130, 5000, 30000):
        rb = _host.owned_row_blocks(N, 0, 1)
        for tpp in (8, 4):                                       # tile_j None: tiles_per_panel tiles per panel
            n, out = raw(N, rb, 148, 0, -1, 3.0, 0, tpp, 0)
            want = _host.build_items_balanced(N, rb, 148, tiles_per_panel=tpp)
            assert n == want.shape[0] and np.array_equal(out[:n], want)
        for max_item in (1, 5):
            n, out = raw(N, rb, 148, 0, -1, 2.5, max_item, 8, 32)
            want = _host.build_items_balanced(N, rb, 148, item_cost=2.5, max_item=max_item, tile_j=32)
            assert n == want.shape[0] and np.array_equal(out[:n], want)
        n, out = raw(N, rb, 148, 0, -1, 2.5, 0, 8, 48, cap=3)     # short buffer: count reported, only cap items written
        want = _host.build_items_balanced(N, rb, 148, item_cost=2.5, tile_j=48)
        assert n == want.shape[0] and np.array_equal(out[:3], want[:3])
    rb = _host.owned_row_blocks(1000, 0, 1)
    assert raw(1000, rb, 148, 5, 5, 2.5, 0, 8, 32)[0] == 0          # empty panel range
    assert L.tsc_host_screen_items(1000, rb.ctypes.data, rb.size, 0, 0, -1, 2.5, 0, 8, 32, None, 0) == -1
    assert L.tsc_host_screen_items(1000, None, 3, 148, 0, -1, 2.5, 0, 8, 32, None, 0) == -1
    assert L.tsc_host_screen_items(0, None, 0, 148, 0, -1, 2.5, 0, 8, 32, None, 0) == 0


def test_native_centring_is_bit_identical_to_numpy():
    """capi.cu: tsc_host_centre == `np.array([s - s.mean(axis=0) for s in structures])` (torsion_module.py:1023) to
    the last bit, for sizes from one atom to more than a thousand, widely different magnitudes and offsets, any number
    of threads, in place as well; torsion_module.centre_structures routes regular inputs to it and ragged / empty
    ones to numpy."""
    from tscode_b200 import torsion_module as tm
    L = _lib.lib()
    rng = np.random.default_rng(11)
    for N, A in [(1, 1), (3, 2), (7, 8), (300, 129), (50, 1025), (1000, 200), (4000, 63)]:
        S = rng.normal(size=(N, A, 3)) * rng.choice([1e-6, 1e-3, 1.0, 1e3]) + rng.normal(size=(N, 1, 3)) * 10.0
        ref = np.array([s - s.mean(axis=0) for s in S])
        for nt in (1, 3, 16):
            out = np.full_like(S, np.nan)
            assert L.tsc_host_centre(S.ctypes.data, N, A, out.ctypes.data, nt) == 0
            assert np.array_equal(out, ref), (N, A, nt)
        inplace = S.copy()
        assert L.tsc_host_centre(inplace.ctypes.data, N, A, inplace.ctypes.data, 4) == 0
        assert np.array_equal(inplace, ref)
        assert np.array_equal(tm.centre_structures(S), ref)
        assert np.array_equal(tm.centre_structures(S.astype(np.float32)),
                              np.array([s - s.mean(axis=0) for s in S.astype(np.float32).astype(np.float64)]))
        assert np.array_equal(tm.centre_structures(S[:, ::-1][:, :, ::-1]),                 # non-contiguous view
                              np.array([s - s.mean(axis=0) for s in S[:, ::-1][:, :, ::-1]]))
        assert np.array_equal(tm.centre_structures(list(S)), ref)
    assert L.tsc_host_centre(None, 0, 5, None, 1) == 0 and L.tsc_host_centre(None, 4, 5, None, 1) == -1
    assert tm.centre_structures(np.zeros((0, 5, 3))).shape[0] == 0
    inf = np.ones((2, 3, 3)); inf[0, 1, 0] = np.inf; inf[1, 2, 2] = np.nan
    with np.errstate(invalid="ignore"):
        assert np.array_equal(tm.centre_structures(inf), np.array([s - s.mean(axis=0) for s in inf]), equal_nan=True)


def test_native_cluster_survivor_choice_equals_python_and_networkx():
    """capi.cu: tsc_host_cluster_rejects restates CPython's set / dict iteration orders (tuple hash, open-addressing
    probe sequence, growth policy) to pick the same `group[0]` per connected component as the reference's
    nx.Graph(set_of_tuples) does.  Against the Python restatement on 1 500 random match sets (sizes across several
    table growths, both branches of the subgraph-view iteration) and against networkx itself on a tenth of them."""
    import random
    from tscode_b200 import torsion_module as tm
    L = _lib.lib()

    def c_rejects(mi, mj, nmax):
        mi = np.ascontiguousarray(mi, dtype=np.int32)
        mj = np.ascontiguousarray(mj, dtype=np.int32)
        rej = np.empty(max(2 * len(mi), 1), dtype=np.int32)
        n = L.tsc_host_cluster_rejects(mi.ctypes.data, mj.ctypes.data, len(mi), nmax, rej.ctypes.data)
        assert n >= 0
        return sorted(rej[:n].tolist())

    rnd = random.Random(7)
    cases = 0
    for trial in range(1500):
        m = rnd.choice((3, 8, 30, 200, 1000, 6000))
        mi, mj, seen = [], [], set()
        if rnd.random() < 0.6:                       # as the replay produces them: rows ascending, one match per row
            dens = rnd.random()
            for i in range(m - 1):
                if rnd.random() < dens:
                    mi.append(i); mj.append(rnd.randrange(i + 1, m))
        else:
            for _ in range(rnd.randrange(1, 3 * m if m < 1000 else m)):
                a, b = rnd.randrange(m), rnd.randrange(m)
                if a != b and (min(a, b), max(a, b)) not in seen:
                    seen.add((min(a, b), max(a, b))); mi.append(min(a, b)); mj.append(max(a, b))
        if not mi:
            continue
        matches = set(zip(mi, mj))
        got = c_rejects(mi, mj, m)
        assert got == sorted(tm._cluster_rejects_fast(matches)), (m, len(mi))
        if trial % 10 == 0:
            assert got == sorted(tm._cluster_rejects_nx(matches)), (m, len(mi))
        cases += 1
    assert cases > 1000
    assert L.tsc_host_cluster_rejects(None, None, 0, 5, None) == 0
    bad = np.array([7], dtype=np.int32)
    assert L.tsc_host_cluster_rejects(bad.ctypes.data, bad.ctypes.data, 1, 5, bad.ctypes.data) == -1
