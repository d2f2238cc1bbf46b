"""Host-side bookkeeping of the RMSD prune that needs no GPU: packed-layout geometry, tile
lists, row sharding, the k-ladder schedule.  Pure numpy so the CPU test-suite covers it."""
from __future__ import annotations

import math

import numpy as np

CB = 32          # conformers per block   (tsc_common.cuh)
KS = 20          # atoms per slab

#: the reference's ladder of chunk counts, tscode/rmsd_pruning.py:186-188
LADDER = (5e5, 2e5, 1e5, 5e4, 2e4, 1e4, 5000, 2000, 1000, 500, 200, 100, 50, 20, 10, 5, 2, 1)


def num_blocks(N: int) -> int:
    return (N + CB - 1) // CB


def num_blocks_padded(N: int) -> int:
    nb = num_blocks(N)
    nb += nb & 1
    return max(nb, 2)


def num_slabs(M: int) -> int:
    return (M + KS - 1) // KS


def packed_doubles(N: int, M: int) -> int:
    return num_slabs(M) * num_blocks_padded(N) * 3 * CB * KS


PANEL_BLOCKS = 4      # 128-row panels (the tcgen05 variant's UMMA M) = 4 row blocks of 32


def owned_row_blocks(N: int, rank: int = 0, world: int = 1) -> np.ndarray:
    """Row sharding: row block ib has ~(nb - ib) column blocks of work, so contiguous
    ranges would be badly imbalanced.  128-row panels are dealt to the ranks in snake order
    (0..w-1, w-1..0, ...), which cancels the linear decrease of work per panel (< 0.5 % imbalance)."""
    ib = np.arange(num_blocks(N), dtype=np.int32)
    panel = ib // PANEL_BLOCKS
    grp, pos = panel // world, panel % world
    owner = np.where(grp % 2 == 0, pos, world - 1 - pos)
    return ib[owner == rank]


def build_screen_items(N: int, row_blocks: np.ndarray, n_ctas: int, panel_lo: int = 0, panel_hi: int | None = None,
                       item_cost: float = 2.5, tile_j: int = 32) -> np.ndarray:
    """Work items of the default screen (rmsd_screen.cu): as build_items_balanced, with j tiles of `tile_j`
    (32, 48 or 64) conformers; panel p starts at the tile that holds its first column, 128 p // tile_j.  Built by the
    library's host code (capi.cu: tsc_host_screen_items — the same rule, compared with build_items_balanced entry by
    entry in the tests): every new ensemble size needs nine such lists before its first launch, 12 ms of Python for
    50 000 structures against 0.1 ms."""
    from ._lib import lib
    rb = np.ascontiguousarray(row_blocks, dtype=np.int32)
    args = (int(N), rb.ctypes.data, int(rb.size), int(n_ctas), int(panel_lo), -1 if panel_hi is None else int(panel_hi),
            float(item_cost * 32.0 / tile_j), 0, 8, int(tile_j))
    n = int(lib().tsc_host_screen_items(*args, None, 0))
    if n < 0:
        raise ValueError("tsc_host_screen_items: bad arguments")
    items = np.empty((n, 4), dtype=np.int32)
    if n and int(lib().tsc_host_screen_items(*args, items.ctypes.data, n)) != n:
        raise RuntimeError("tsc_host_screen_items: inconsistent item count")
    return items


def screen_frame(first_heavy: np.ndarray):
    """Frame and column weights of the default screen for an ensemble, from its first structure (rmsd_screen.cu, ScFrame):
    returns (frame, ratio) with frame = 12 float64 [Q row-major (rows = principal axes of the heavy atoms' second-moment
    tensor), t_b = sqrt(w_b / 3)], w_b = (l1 + l2 + l3) / l_b (eigenvalues floored at 1e-4 of their sum, weights then
    inflated so that sum 1 / w_b <= 1 holds in floating point), and ratio = sqrt(3 sum l^2) / sum l, the looseness of
    the UNweighted Samuelson bound for this shape (1 = isotropic).  For an isotropic molecule Q = I, t = 1 exactly.
    A speed device only: any orthogonal Q and any weights with sum 1 / w_b <= 1 are sound (checked by the library)."""
    ident = np.concatenate([np.eye(3).ravel(), np.ones(3)])
    X = np.asarray(first_heavy, dtype=np.float64).reshape(-1, 3)
    if X.shape[0] == 0 or not np.isfinite(X).all():
        return ident, np.inf
    lam, vec = np.linalg.eigh(X.T @ X)
    tot = float(lam.sum())
    if not np.isfinite(tot) or tot <= 0.0:
        return ident, np.inf
    ratio = float(np.sqrt(3.0 * (lam ** 2).sum())) / tot
    if ratio <= 1.0005:                                          # isotropic to the last digit that matters: plain Samuelson
        return ident, ratio
    lam = np.maximum(lam, 1e-4 * tot)
    w = lam.sum() / lam * (1.0 + 1e-9)
    Q = vec.T.copy()
    for _ in range(2):                                           # re-orthonormalise to the library's 1e-13 (Gram-Schmidt)
        Q[0] /= np.linalg.norm(Q[0])
        Q[1] -= Q[0] * (Q[1] @ Q[0]); Q[1] /= np.linalg.norm(Q[1])
        Q[2] -= Q[0] * (Q[2] @ Q[0]) + Q[1] * (Q[2] @ Q[1]); Q[2] /= np.linalg.norm(Q[2])
    assert float((1.0 / w).sum()) <= 1.0
    return np.concatenate([Q.ravel(), np.sqrt(w / 3.0)]), ratio


def screen_mode_for(first_heavy: np.ndarray) -> int:
    """Form of the default screen from the shape of the first structure alone (rmsd_screen.cu; a speed decision only:
    every form is conservative): near-isotropic molecules (looseness of the unweighted Samuelson bound <= 1.03) take the
    cheapest form, mode 0: the bound only, on 48-wide tiles; everything else mode 1: the bound in the frame of
    screen_frame, then the FP32 quartic sign test for the groups of 64 pairs in which it left a pair undecided.
    screen_plan refines this with a sample of pairs."""
    return 0 if screen_frame(first_heavy)[1] <= 1.03 else 1


SAMPLE_PAIRS = 128


def plan_mode(ratio: float, undecided: float) -> int:
    """The rule of screen_plan: sample fraction u of pairs the weighted bound leaves undecided, looseness `ratio` of the
    unweighted bound for the first structure's shape -> form of the screen."""
    if undecided <= 0.005 and ratio <= 1.03:
        return 0
    return 1 if undecided <= 0.03 else 2


def sample_pair_indices(N: int, k: int = SAMPLE_PAIRS):
    """Fixed pseudo-random pairs (i != j) for screen_plan: the same on every rank and every call
    (capi.cu: tsc_host_sample_pairs)."""
    from ._lib import lib
    k = int(k) if N >= 2 else 0
    pi, pj = np.empty(k, np.int64), np.empty(k, np.int64)
    if k:
        lib().tsc_host_sample_pairs(int(N), k, pi.ctypes.data, pj.ctypes.data)
    return pi, pj


def screen_plan_native(S: np.ndarray, heavy_idx: np.ndarray, thr: float, first: int, pi: np.ndarray, pj: np.ndarray):
    """(frame, mode, undecided fraction) like screen_plan, computed by the library's host code on the caller's array S
    (rows, A, 3) float64 C-contiguous without copying it (capi.cu: tsc_host_screen_plan; ~0.1 ms against ~2 ms of
    numpy: it is on the latency path of every prune call)."""
    from ._lib import lib
    assert S.dtype == np.float64 and S.flags.c_contiguous and S.ndim == 3 and S.shape[2] == 3
    h = np.ascontiguousarray(heavy_idx, dtype=np.int32)
    pi = np.ascontiguousarray(pi, dtype=np.int64)
    pj = np.ascontiguousarray(pj, dtype=np.int64)
    frame, out = np.empty(12), np.empty(2)
    rc = lib().tsc_host_screen_plan(S.ctypes.data, int(S.shape[1]), h.ctypes.data, int(h.size), int(first),
                                    pi.ctypes.data, pj.ctypes.data, int(pi.size), float(thr), frame.ctypes.data,
                                    out[0:].ctypes.data, out[1:].ctypes.data)
    if rc != 0:
        raise ValueError("tsc_host_screen_plan: bad arguments")
    return frame, plan_mode(float(out[0]), float(out[1])), float(out[1])


def screen_plan(first_heavy: np.ndarray, P: np.ndarray, Qs: np.ndarray, thr: float):
    """(frame, mode, undecided fraction): screen_frame plus a look at a sample of pairs (P[k], Qs[k]) (heavy atoms, as
    given).  How often the weighted bound fails to exclude a pair depends on the ENSEMBLE, not on the first structure:
    conformers of one molecule are excluded almost always (then the bound alone, mode 0, or the bound with the
    quartic as a backstop, mode 1, is fastest), but e.g. embedded poses of several fragments in many arrangements are
    not, and a failed bound costs a group of 64 pairs the quartic on top (mode 2, the quartic for every pair, is
    then cheaper).  Sample fraction u of pairs the bound leaves undecided (similar pairs included):
        u <= 0.5 % and near-isotropic shape -> 0;   u <= 3 % -> 1;   else 2      (plan_mode)
    A speed decision only.  (numpy statement of screen_plan_native, which is what the product calls.)"""
    frame, ratio = screen_frame(first_heavy)
    P = np.asarray(P, dtype=np.float64)
    Qs = np.asarray(Qs, dtype=np.float64)
    if P.shape[0] == 0:
        return frame, plan_mode(ratio, 0.0), 0.0
    Q, t = frame[:9].reshape(3, 3), frame[9:]
    M = P.shape[1]
    C = np.einsum("kma,kmb->kab", P, Qs)                      # plain covariances
    Sw = np.einsum("xa,kab,yb->kxy", Q, C, Q) * t             # in the frame, columns scaled
    f = (Sw ** 2).sum((1, 2))
    Gi, Gj = (P ** 2).sum((1, 2)), (Qs ** 2).sum((1, 2))
    Gw = (((Qs @ Q.T) * t) ** 2).sum((1, 2))                  # squared norm of the scaled column conformer
    lf = 0.5 * (Gi + Gj) - 0.5 * M * thr * thr - np.sqrt(3.0) * 1.05e-3 * np.sqrt(Gi * np.maximum(Gw, Gj))
    und = float(np.mean(~((lf > 0) & (3.00004 * f < lf * lf))))
    return frame, plan_mode(ratio, und), und


def build_items_balanced(N: int, row_blocks: np.ndarray, n_ctas: int, panel_lo: int = 0,
                              panel_hi: int | None = None, item_cost: float = 3.0, max_item: int = 0,
                              tiles_per_panel: int = 8, tile_j: int | None = None) -> np.ndarray:
    """Work items of the tcgen05 pre-screen for a persistent grid of `n_ctas` CTAs, each of which takes the array
    entries b, b + n_ctas, b + 2 n_ctas, ...: rows {panel, first j tile, j tile count, local row block of the panel};
    panel p (128 rows) needs the j tiles from tiles_per_panel * p (tile_j given: from 128 p // tile_j) to the padded
    end (optionally only the panels in [panel_lo, panel_hi)).  Partitioned linearly: the (panel, j tile) pairs are laid out panel after panel and
    every CTA gets one contiguous stretch of equal cost (tiles + `item_cost` tiles per item start: pipeline drain
    and panel rows -> TMEM, measured ~2 400 cycles), i.e. one item per panel its stretch touches.  Plain
    round-robin dealing of 128-tile chunks left the busiest CTA of C3 5 % above the mean, and 20 % in the eight
    sub-launches of the pipelined upload.  CTAs whose stretch touches fewer panels than the longest list get
    empty items (count 0, a valid panel) in the last rounds."""
    rb = np.asarray(row_blocks, dtype=np.int64)
    n_pan = (N + 127) // 128
    if tile_j is None:
        tpp = int(tiles_per_panel)
        njt = n_pan * tpp
        first = lambda p: tpp * p                                  # noqa: E731
    else:
        tpp = max(1, 128 // int(tile_j))
        njt = (n_pan * 128 + int(tile_j) - 1) // int(tile_j)      # tiles up to the padded width of the bit rows
        first = lambda p: (128 * p) // int(tile_j)                # noqa: E731
    if panel_hi is None:
        panel_hi = (N + 127) // 128
    panels = [(int(ib // PANEL_BLOCKS), lb) for lb, ib in enumerate(rb)
              if ib % PANEL_BLOCKS == 0 and panel_lo <= ib // PANEL_BLOCKS < panel_hi]
    total = sum(njt - first(p) for p, _ in panels)
    if total == 0:
        return np.zeros((0, 4), np.int32)
    n_bins = max(1, min(n_ctas, total // tpp))
    bins = [[] for _ in range(n_bins)]
    left = float(total + item_cost * (len(panels) + n_bins))      # cost still to hand out (upper estimate)
    b, budget = 0, left / n_bins                                  # current CTA and what it may still take
    for p, lb in panels:
        j, end = first(p), njt
        while j < end:
            room = int(budget - item_cost)
            if room < 4 and b + 1 < n_bins:                       # not worth starting an item here: next CTA
                b += 1
                budget = left / (n_bins - b)                      # re-balance over the CTAs that are left
                continue
            take = end - j if b + 1 == n_bins else max(1, min(end - j, room))
            bins[b].append((p, j, take, lb))
            j += take
            budget -= take + item_cost
            left -= take + item_cost
    bins = [x for x in bins if x]
    if max_item > 0:                                 # (measurement aid) the same stretches, cut into short items
        bins = [[(p, j + o, min(max_item, cnt - o), lb) for p, j, cnt, lb in x for o in range(0, cnt, max_item)] for x in bins]
    rounds = max(len(x) for x in bins)
    pad_p, pad_lb = panels[0]
    # The kernel's stride is its grid = min(n_ctas, number of entries).  One round: one entry per CTA, any grid works.
    # Several rounds: the array is laid out with stride n_ctas exactly (CTAs without work get an empty first entry), so
    # that the entries a CTA visits are the ones meant for it and an empty entry really ends its list.
    if rounds > 1:
        bins += [[] for _ in range(n_ctas - len(bins))]
    items = []
    for r in range(rounds):
        last = max(q for q in range(len(bins)) if len(bins[q]) > r)
        for q in range(len(bins) if r + 1 < rounds else last + 1):
            items.append(bins[q][r] if len(bins[q]) > r else (pad_p, first(pad_p), 0, pad_lb))
    return np.asarray(items, dtype=np.int32).reshape(-1, 4)


def screen_rows_padded(N: int) -> int:
    """Rows the screen's operand images, G, sG and CT are padded to: whole 128-row panels and whole tiles of 32 / 48 / 64
    conformers (rmsd_screen.cu: tsc_screen_rows_padded)."""
    return ((N + 383) // 384) * 384


def build_tiles(N: int, row_blocks: np.ndarray) -> np.ndarray:
    """(n_tiles, 4) int32 rows {ib, jp, lb, 0}: for each owned row block ib (local index lb) the
    J block pairs jp = ib//2 .. nb_pad/2 - 1 (every word >= ib of the row gets written)."""
    njp = num_blocks_padded(N) // 2
    rb = np.asarray(row_blocks, dtype=np.int64)
    if rb.size == 0:
        return np.zeros((0, 4), np.int32)
    first = rb // 2
    counts = njp - first
    total = int(counts.sum())
    lb = np.repeat(np.arange(rb.size, dtype=np.int64), counts)
    ib = rb[lb]
    start = np.cumsum(counts) - counts
    jp = np.arange(total, dtype=np.int64) - np.repeat(start, counts) + np.repeat(first, counts)
    tiles = np.zeros((total, 4), np.int32)
    tiles[:, 0] = ib
    tiles[:, 1] = jp
    tiles[:, 2] = lb
    return tiles


def pairs_in_tiles(N: int) -> int:
    return N * (N - 1) // 2


def ladder_gate(k, n_active: int, gate: int = 20) -> bool:
    """`k == 1 or 20*k < count_nonzero(mask)`  (rmsd_pruning.py:192)."""
    return k == 1 or gate * k < n_active


def chunk_size(N: int, k) -> int:
    """`int(len(structures) // k)`  (rmsd_pruning.py:136)."""
    return int(N // k)


def run_ladder(N: int, round_fn, gate: int = 20):
    """Drive the ladder: round_fn(k:int, cs:int) -> n_active after the round.
    Returns the list of k actually run (data dependent)."""
    n_active = N
    ran = []
    for k in LADDER:
        if ladder_gate(k, n_active, gate):
            n_active = int(round_fn(int(k), chunk_size(N, k)))
            ran.append(int(k))
    return ran


def global_rows_of(row_blocks: np.ndarray) -> np.ndarray:
    """Global row index of every local sim row (n_rb*32,), for reassembling gathered shards."""
    rb = np.asarray(row_blocks, dtype=np.int64)
    return (rb[:, None] * CB + np.arange(CB, dtype=np.int64)[None, :]).reshape(-1)


def sqrt_threshold_image(thresh: float) -> float:
    """Smallest double t2 with sqrt(t2) >= thresh under correctly-rounded sqrt, so that
    `sqrt(d2) < thresh`  <=>  `d2 < t2` exactly (all_dists + `< thresh`, algebra.py:98-157)."""
    thresh = float(thresh)
    if not (thresh > 0.0):
        return 0.0
    x = thresh * thresh
    while math.sqrt(x) >= thresh:
        x = math.nextafter(x, -math.inf)
    while math.sqrt(x) < thresh:
        x = math.nextafter(x, math.inf)
    return x


def ladder_pairlist_model(pairs, N: int, gate: int = 20):
    """Host model of the fused ladder kernel (eliminate.cu, elim_fused_kernel): the k-ladder of
    rmsd_pruning.py:186-204 driven by an unordered list of similar pairs (i < j) instead of a
    similarity matrix.  Same phases as the kernel; used by the CPU tests to pin the formulation
    against the oracle.  Returns (mask, rounds_run)."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    INF = np.iinfo(np.int64).max
    active = np.ones(N, bool)
    keys_f, keys_s = np.zeros(0, np.int64), np.zeros(0, np.int64)
    rounds = []
    idx = np.arange(N, dtype=np.int64)
    for k in LADDER:
        n_active = int(active.sum())
        if not ladder_gate(k, n_active, gate):
            continue
        k = int(k)
        cs = chunk_size(N, k)
        if cs > 0:
            c = np.minimum(idx // cs, k - 1)
            first = c * cs
            last = np.where(c == k - 1, N, first + cs)
        else:
            first = np.zeros(N, np.int64)
            last = np.full(N, N, np.int64)
        # phase A: cache bitmap of this round's chunking, first similar active partner inside the chunk
        cb = np.zeros(N + 64, bool)
        if keys_f.size:
            if cs > 0:
                ok = (keys_f % cs == 0) & (keys_f // cs < k)
                kl = np.where(keys_f // cs == k - 1, N, keys_f + cs)
            else:
                ok = keys_f == 0
                kl = np.full(keys_f.shape, N)
            ok &= keys_s < kl
            cb[keys_s[ok]] = True
        first_sim = np.full(N, INF)
        if pairs.size:
            i, j = pairs[:, 0], pairs[:, 1]
            ok = active[i] & active[j] & (j < last[i])
            np.minimum.at(first_sim, i[ok], j[ok])
        # phase B
        new_active = active.copy()
        nf, ns = [], []
        for r in np.flatnonzero(active):
            js = first_sim[r]
            limit = js if js != INF else last[r] - 1          # inclusive
            hit = False
            if cb.any() and r + 1 <= limit:
                jj = np.arange(r + 1, limit + 1)
                hit = bool(np.any(active[jj] & cb[first[r] + jj - r]))
            if not hit and js != INF:
                new_active[r] = False
                nf.append(first[r]); ns.append(first[r] + js - r)
        keys_f = np.concatenate([keys_f, np.asarray(nf, np.int64)])
        keys_s = np.concatenate([keys_s, np.asarray(ns, np.int64)])
        active = new_active
        rounds.append(k)
    return active, rounds
