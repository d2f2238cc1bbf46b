"""B200-native drop-in for the rotor-corrected RMSD pruning of tscode/torsion_module.py.

    prune_conformers_rmsd_rot_corr(structures, atomnos, graph, max_rmsd=0.25, verbose=False,
                                   logfunction=None) -> (structures_centred[mask], mask)      # :1013-1161
    rotationally_corrected_rmsd(ref, coord, atomnos, torsions, graph, angles) -> float          # :953-1011

Split of work:
  * chemistry perception (which bonds are "dummy" rotors, their n-fold angle sets, rotation
    masks, sub-graph node lists) is pair-independent host work and stays in Python.  It is
    delegated to the reference's own helpers when `tscode` is importable (the drop-in scenario:
    `perceive_torsions`), or supplied by the caller as a `TorsionInfo` (tests, benchmarks);
  * the per-pair numerics — sum_t n_t local Kabsch RMSDs + rotations + one global Kabsch RMSD —
    run on the GPU for all pairs at once, statelessly (rotcorr.cu);
  * the grouping loop (:1076-1152) is replayed literally on the host on the resulting boolean
    matrix, with Python `set` -> `nx.Graph` -> `connected_components` exactly as the reference
    builds them, because which member of a cluster survives depends on that iteration order;
  * the reference mutates compared structures in place and returns them mutated; the host replay
    tracks every structure's rotor state ((best angle of the pair) + (state of the first
    structure), mod 360 — exact because every angle set is a full n-fold orbit) and the returned
    coordinates are produced from those states (tsc_rotcorr_apply).

Reference guard kept: more than `max_structures` (750, :1056) structures or no dummy rotor ->
centred structures and an all-True mask.  Pass max_structures=None to lift it (beyond what the
reference itself would run; say so when quoting such a number).
"""
from __future__ import annotations

import copy
import os
from dataclasses import dataclass, field

import numpy as np

from . import _host
from ._lib import check, lib, ptr, require_cuda, stream_ptr

MAX_T, MAX_ANG = 20, 6          # rotcorr.cu: 3 bits of best-angle code per rotor in a 64-bit word


class UnsupportedRotors(ValueError):
    """More symmetric rotors (or angles per rotor) than the kernels' 64-bit code holds; install.py's wrapper hands such a
    call back to the reference's own function."""
_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
            "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe").split()


@dataclass
class TorsionInfo:
    """Pair-independent description of the dummy rotors of one molecule."""
    torsions: list                      # T quadruplets (i1, i2, i3, i4), dummy side last (:1049)
    angles: list                        # T tuples of degrees (:112-118)
    rot_masks: np.ndarray               # (T, A) bool, _get_rotation_mask (:301-325)
    node_masks: np.ndarray              # (T, A) bool, heavy atoms of the rotor's sub-graph (:964-977)
    log: list = field(default_factory=list)

    @property
    def T(self):
        return len(self.torsions)


def perceive_torsions(ref_centred, atomnos, graph) -> TorsionInfo:
    """Set-up block of prune_conformers_rmsd_rot_corr (:1026-1049) plus the per-torsion
    quantities rotationally_corrected_rmsd recomputes for every pair (:964-977, :984).  Uses the
    reference's perception helpers: requires the `tscode` package (drop-in scenario)."""
    try:
        import networkx as nx
        from tscode.torsion_module import (_get_hydrogen_bonds, _get_rotation_mask, _get_torsions, _is_nondummy)
        from tscode.utils import get_double_bonds_indices
    except Exception as e:           # pragma: no cover - exercised only without the reference
        raise RuntimeError("torsion perception is delegated to the reference's helpers "
                           "(tscode.torsion_module); install TSCoDe or pass torsion_info=...") from e
    atomnos = np.asarray(atomnos)
    g = copy.deepcopy(graph)
    for hb in _get_hydrogen_bonds(ref_centred, atomnos, g):
        g.add_edge(*hb)
    tors = _get_torsions(g, hydrogen_bonds=_get_hydrogen_bonds(ref_centred, atomnos, g),
                         double_bonds=get_double_bonds_indices(ref_centred, atomnos), keepdummy=True)
    tors = [t for t in tors if not (_is_nondummy(t.i2, t.i3, g) and _is_nondummy(t.i3, t.i2, g))]
    tors = [t for t in tors if 1 not in [atomnos[i] for i in t.torsion]]
    angles = [tuple(t.get_angles()) for t in tors]
    quads = [tuple(t.torsion) if _is_nondummy(t.i2, t.i3, g) else tuple(reversed(t.torsion)) for t in tors]
    A = len(atomnos)
    rot, nodes = np.zeros((len(quads), A), bool), np.zeros((len(quads), A), bool)
    for k, t in enumerate(quads):
        for o in quads:
            if o is not t:
                g.remove_edge(o[1], o[2])
        comp = [s for s in nx.connected_components(g) if t[1] in s][0]
        for o in quads:
            if o is not t:
                g.add_edge(o[1], o[2])
        nodes[k, [i for i in comp if atomnos[i] != 1]] = True
        rot[k] = _get_rotation_mask(g, t)
    return TorsionInfo([tuple(int(x) for x in q) for q in quads], angles, rot, nodes)


class RotCorrPruner:
    """All-pairs rotor-corrected RMSD on the GPU + literal host replay of the grouping loop."""

    def __init__(self, structures_centred, atomnos, info: TorsionInfo, max_rmsd=0.25, *, want_codes=True,
                 want_rmsd=False):
        torch = require_cuda()
        self.torch = torch
        dev = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.dev = dev
        Sc = np.ascontiguousarray(structures_centred, dtype=np.float64)
        self.N, self.A = Sc.shape[0], Sc.shape[1]
        self.info, self.max_rmsd = info, float(max_rmsd)
        T = info.T
        if T > MAX_T or any(len(a) > MAX_ANG for a in info.angles):
            raise UnsupportedRotors(f"at most {MAX_T} rotors with {MAX_ANG} angles each are supported")
        self.Sc = torch.from_numpy(Sc).to(dev)
        atomnos = np.asarray(atomnos)
        self.heavy = torch.from_numpy((atomnos != 1).astype(np.uint8)).to(dev)
        self.i2 = torch.tensor([t[1] for t in info.torsions], dtype=torch.int32, device=dev)
        self.i3 = torch.tensor([t[2] for t in info.torsions], dtype=torch.int32, device=dev)
        self.n_ang = torch.tensor([len(a) for a in info.angles], dtype=torch.int32, device=dev)
        ang = np.zeros((max(T, 1), MAX_ANG))
        for k, a in enumerate(info.angles):
            ang[k, :len(a)] = a
        half = ang * np.pi / 180 / 2                                 # algebra.py:337-341
        self.ang_table = ang
        self.sin_half = torch.from_numpy(np.sin(half)).to(dev)
        self.cos_half = torch.from_numpy(np.cos(half)).to(dev)
        self.rot_mask = torch.from_numpy(np.ascontiguousarray(info.rot_masks, dtype=np.uint8)).to(dev) \
            if T else torch.zeros((1, self.A), dtype=torch.uint8, device=dev)
        self.node_mask = torch.from_numpy(np.ascontiguousarray(info.node_masks, dtype=np.uint8)).to(dev) \
            if T else torch.zeros((1, self.A), dtype=torch.uint8, device=dev)
        N = self.N
        self.Wb = (N + 31) // 32
        self.sim_bits = torch.zeros((max(N, 1), max(self.Wb, 1)), dtype=torch.int32, device=dev)
        self.codes = torch.empty((N, N), dtype=torch.int64, device=dev) if want_codes else None
        self.rmsd = torch.zeros((N, N), dtype=torch.float64, device=dev) if want_rmsd else None
        self.near = torch.zeros(1, dtype=torch.int64, device=dev)

    def similarity(self, row_begin=0, row_end=None):
        row_end = self.N if row_end is None else row_end
        check(lib().tsc_rotcorr_pairs(ptr(self.Sc), self.N, self.A, ptr(self.heavy), self.info.T, ptr(self.i2),
                                      ptr(self.i3), ptr(self.n_ang), ptr(self.sin_half), ptr(self.cos_half),
                                      ptr(self.rot_mask), ptr(self.node_mask), row_begin, row_end, self.max_rmsd,
                                      ptr(self.sim_bits), ptr(self.codes), ptr(self.rmsd), ptr(self.near),
                                      stream_ptr()), "tsc_rotcorr_pairs")

    def scan(self, rank=0, world=1, group=None):
        """Forward scan (tsc_rotcorr_scan): returns (first_hit (N,) int64 numpy, lookup) where
        first_hit[i] is the first j > i similar to i (N if none) and lookup(i, js) gives the best
        rotor angles (len(js), T) of the pairs (i, j <= first_hit[i]) — all the grouping loop needs
        in stateless mode.  The codes are compacted on the device before they cross PCIe.
        Several ranks (rot_corr): rows are dealt round-robin (row i to rank i % world: a row's cost is
        data dependent, neighbours are alike), every rank scans its rows, first hits and compacted codes are
        all-gathered, and every rank holds the complete result."""
        torch, N, T = self.torch, self.N, self.info.T
        dev = self.dev
        if self.codes is None:
            self.codes = torch.empty((N, N), dtype=torch.int64, device=dev)      # only entries the scan writes are read
        first = torch.full((max(N, 1),), N, dtype=torch.int32, device=dev)
        counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.near.zero_()
        check(lib().tsc_rotcorr_scan(ptr(self.Sc), N, self.A, ptr(self.heavy), T, ptr(self.i2), ptr(self.i3),
                                     ptr(self.n_ang), ptr(self.sin_half), ptr(self.cos_half), ptr(self.rot_mask),
                                     ptr(self.node_mask), rank, N, world, self.max_rmsd, ptr(first), ptr(self.codes),
                                     ptr(self.rmsd), ptr(self.near), ptr(counter), stream_ptr()), "tsc_rotcorr_scan")
        rows = torch.arange(rank, N, world, device=dev, dtype=torch.int64)      # the rows this rank scanned
        last = torch.clamp(first[rows].to(torch.int64), max=N - 1)               # last column the loop can visit
        lengths = torch.clamp(last - rows, min=0)
        offsets = torch.cumsum(lengths, 0) - lengths
        total = int(lengths.sum().item())
        if total:
            r = torch.repeat_interleave(torch.arange(rows.numel(), device=dev), lengths)
            c = torch.arange(total, device=dev, dtype=torch.int64) - offsets[r] + rows[r] + 1
            compact_d = self.codes.view(-1)[rows[r] * N + c]
        else:
            compact_d = torch.zeros(0, dtype=torch.int64, device=dev)
        if world > 1:
            import torch.distributed as dist
            from .embeds import gather_varlen
            per = (N + world - 1) // world
            fpad = torch.full((per,), N, dtype=torch.int32, device=dev)
            fpad[:rows.numel()] = first[rows]
            fall = torch.empty(world * per, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(fall, fpad, group=group)
            first_all = torch.full((N,), N, dtype=torch.int32, device=dev)
            for q in range(world):
                nq = (N - q + world - 1) // world
                first_all[q::world] = fall[q * per:q * per + nq]
            compact_all = gather_varlen(compact_d, world, group)                  # rank-major: rank q's rows q, q + world, ...
            near = self.near.clone()
            dist.all_reduce(near, group=group)
            self.near.copy_(near)
            fh = first_all.cpu().numpy().astype(np.int64)
            # offsets of every row into the rank-major compact array
            ln = np.clip(np.minimum(fh, N - 1) - np.arange(N), 0, None)
            off = np.zeros(N, dtype=np.int64)
            base = 0
            for q in range(world):
                lq = ln[q::world]
                off[q::world] = base + np.cumsum(lq) - lq
                base += int(lq.sum())
            compact = compact_all.cpu().numpy().view(np.uint64)
            self.pairs_evaluated = int(ln.sum())
        else:
            fh = first[:N].cpu().numpy().astype(np.int64)
            off = offsets.cpu().numpy()
            compact = compact_d.cpu().numpy().view(np.uint64)
            self.pairs_evaluated = total
        table = self.ang_table

        def lookup(i, js):
            cc = compact[off[i] + (np.asarray(js) - i - 1)]
            return np.stack([table[t][((cc >> np.uint64(3 * t)) & np.uint64(7)).astype(np.int64)] for t in range(T)], axis=-1)
        lookup.T = T
        lookup.compact, lookup.off, lookup.table = compact, np.ascontiguousarray(off, dtype=np.int64), table
        return fh, lookup

    def similar_matrix(self):
        """(N, N) bool, upper triangle, on the host."""
        bits = self.sim_bits.cpu().numpy().view(np.uint32)
        return np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, : self.N].astype(bool)

    def best_angles(self):
        """Per-pair best rotor angles as a lookup `f(i, js) -> (len(js), T) degrees`, decoded lazily
        from the (N, N) uint64 codes (3 bits per rotor) so that large N stays at 8 bytes per pair."""
        codes = self.codes.cpu().numpy().view(np.uint64)
        T, table = self.info.T, self.ang_table

        def lookup(i, js):
            c = codes[i, js]
            return np.stack([table[t][((c >> np.uint64(3 * t)) & np.uint64(7)).astype(np.int64)] for t in range(T)], axis=-1)
        lookup.T = T
        return lookup

    def prune_stateful(self, verbose=False):
        """Exact emulation of the reference loop (:1076-1152) INCLUDING its in-place mutation: rows are
        replayed in the reference's order; each row is one kernel launch over the not-yet-cached
        second structures, evaluated from the current (mutated) coordinates; the copies of the
        structures visited before the reference's `break` are committed.  Returns
        (mutated structures[mask] as numpy, mask, n_near_threshold)."""
        import networkx as nx
        torch, N, A, L = self.torch, self.N, self.A, lib()
        cur = self.Sc.clone()
        staged = torch.empty((N, A, 3), dtype=torch.float64, device=self.dev)
        rmsd_d = torch.empty(N, dtype=torch.float64, device=self.dev)
        codes_d = torch.empty(N, dtype=torch.int64, device=self.dev)
        js_d = torch.empty(N, dtype=torch.int32, device=self.dev)
        js_pin = torch.empty(N, dtype=torch.int32).pin_memory()
        final_mask = np.ones(N, dtype=bool)
        cached = np.zeros((N, N), dtype=bool)
        near = 0
        st = stream_ptr()
        for k in _host.LADDER:
            num_active = int(np.count_nonzero(final_mask))
            if not (k == 1 or 5 * k < num_active):
                continue
            if verbose:
                print(f"Working on subgroups with k={k} ({num_active} candidates left) {' ' * 10}", end="\r")
            d = int(N // k)
            for step in range(int(k)):
                _l = len(range(d * step, num_active)) if step == k - 1 else len(range(d * step, int(d * (step + 1))))
                if _l <= 1:
                    continue
                base = d * step
                matches = set()
                for i_rel in range(_l):
                    i, lo, hi = base + i_rel, base + i_rel + 1, base + _l
                    if lo >= hi:
                        continue
                    js = lo + np.flatnonzero(~cached[i, lo:hi])
                    n = int(js.size)
                    if n == 0:
                        continue
                    js_pin[:n] = torch.from_numpy(js.astype(np.int32))
                    js_d[:n].copy_(js_pin[:n], non_blocking=True)
                    check(L.tsc_rotcorr_row(ptr(cur), N, A, ptr(self.heavy), self.info.T, ptr(self.i2), ptr(self.i3),
                                            ptr(self.n_ang), ptr(self.sin_half), ptr(self.cos_half), ptr(self.rot_mask),
                                            ptr(self.node_mask), i, ptr(js_d), n, ptr(rmsd_d), ptr(codes_d),
                                            ptr(staged), st), "tsc_rotcorr_row")
                    r = rmsd_d[:n].cpu().numpy()
                    hits = np.flatnonzero(r < self.max_rmsd)
                    if hits.size:
                        h = int(hits[0])
                        n_acc = h + 1
                        cached[i, js[:h]] = True
                        matches.add((i_rel, int(js[h] - base)))
                    else:
                        n_acc = n
                        cached[i, js] = True
                    near += int(np.count_nonzero(np.abs(r[:n_acc] - self.max_rmsd) < 1e-6))
                    check(L.tsc_rotcorr_commit(ptr(cur), ptr(staged), ptr(js_d), n_acc, A, st), "tsc_rotcorr_commit")
                for rr in _cluster_rejects(matches):
                    final_mask[rr + base] = 0
        keep = torch.from_numpy(np.flatnonzero(final_mask)).to(self.dev)
        return cur[keep].cpu().numpy(), final_mask, near

    def apply_states(self, idx, state_deg):
        """Centred structures idx with rotor states applied -> numpy (n, A, 3)."""
        torch = self.torch
        idx = np.asarray(idx, dtype=np.int64)
        n, T = idx.size, self.info.T
        if n == 0:
            return np.zeros((0, self.A, 3))
        half = (np.asarray(state_deg, dtype=np.float64).reshape(n, T) if T else np.zeros((n, 1))) * np.pi / 180 / 2
        out = torch.empty((n, self.A, 3), dtype=torch.float64, device=self.dev)
        sh = torch.from_numpy(np.ascontiguousarray(np.sin(half))).to(self.dev)
        ch = torch.from_numpy(np.ascontiguousarray(np.cos(half))).to(self.dev)
        d_idx = torch.from_numpy(idx).to(self.dev)
        check(lib().tsc_rotcorr_apply(ptr(self.Sc), n, self.A, ptr(d_idx), T,
                                      ptr(self.i2), ptr(self.i3), ptr(sh), ptr(ch), ptr(self.rot_mask), ptr(out),
                                      stream_ptr()), "tsc_rotcorr_apply")
        return out.cpu().numpy()


def _cluster_rejects_nx(edges):
    """The reference's survivor choice, literally (torsion_module.py:1136-1152, numba_functions.py:203-220,
    optimization_methods.py:341-355): graph of the matches, connected components, keep `group[0]` of each."""
    import networkx as nx
    g = nx.Graph(edges)
    rejects = []
    for group in [tuple(g.subgraph(c).nodes) for c in nx.connected_components(g)]:
        rejects.extend(set(group) - {group[0]})
    return rejects


def _cluster_rejects_fast(edges):
    """Same result as _cluster_rejects_nx without building networkx objects (3 900 chunks of BASELINE configs[3] spent
    0.45 s in Graph / subgraph-view construction).  Which member of a cluster is `group[0]` depends on iteration
    orders, so this restates networkx 3.x step by step with plain dicts and sets — Python's own, hence the same
    orders: nodes in order of first appearance in the edge iteration (Graph.add_edges_from), components in node
    order, each found by the same breadth-first search into a SET (connected._plain_bfs), and the first node of
    `G.subgraph(c).nodes` is the first element of a new set built from c when 2 |c| < |G|, else the first node of G
    that lies in c (coreviews.FilterAtlas.__iter__).  `_cluster_rejects` checks it against networkx itself once per
    process and falls back to the literal form if the installed networkx behaves differently."""
    adj = {}
    for u, v in edges:
        if u not in adj:
            adj[u] = {}
        if v not in adj:
            adj[v] = {}
        adj[u][v] = None
        adj[v][u] = None
    n = len(adj)
    seen_all = set()
    rejects = []
    for v0 in adj:
        if v0 in seen_all:
            continue
        target = n - len(seen_all)
        seen = {v0}
        nextlevel = [v0]
        full = False
        while nextlevel and not full:
            thislevel = nextlevel
            nextlevel = []
            for x in thislevel:
                for w in adj[x]:
                    if w not in seen:
                        seen.add(w)
                        nextlevel.append(w)
                if len(seen) == target:
                    full = True
                    break
        seen_all.update(seen)
        # show_nodes(self.nbunch_iter(c)): a NEW set filled element by element in c's iteration order (every element
        # of c is a node, so the membership test of nbunch_iter is dropped; set(iter(..)) inserts one by one exactly as
        # the generator form does, whereas set(seen) would copy the table with a different pre-sizing)
        nodes = set(iter(seen))
        if 2 * len(nodes) < n:
            first = next(iter(nodes))
        else:
            first = next(x for x in adj if x in nodes)
        rejects.extend(x for x in nodes if x != first)
    return rejects


_CLUSTER_IMPL = []


def _cluster_rejects(edges):
    if not _CLUSTER_IMPL:
        import random
        rnd = random.Random(12345)
        ok = True
        for trial in range(40):
            m = rnd.choice((3, 8, 30, 200))
            es = set()
            for _ in range(rnd.randrange(1, 3 * m)):
                a, b = rnd.randrange(m), rnd.randrange(m)
                if a != b:
                    es.add((min(a, b), max(a, b)))
            es = es if trial % 2 else list(es)
            if es and sorted(_cluster_rejects_fast(es)) != sorted(_cluster_rejects_nx(es)):
                ok = False
                break
        _CLUSTER_IMPL.append(_cluster_rejects_fast if ok else _cluster_rejects_nx)
    return _CLUSTER_IMPL[0](edges)


def ladder_replay(similar, N, best_angles=None, verbose=False):
    """Literal replay of the grouping loop (torsion_module.py:1076-1152) on similar[i, j]
    (i < j) with rotor-state tracking (see module docstring).  The inner pair loop is vectorised
    per row — same visiting order, same cache contents, same `matches` set insertion order.
    Returns (final_mask, state (N, T) degrees)."""
    import networkx as nx
    final_mask = np.ones(N, dtype=bool)
    cached = np.zeros((N, N), dtype=bool)          # cache_set (:1054) as a dense matrix
    if best_angles is None:
        T, lookup = 0, None
    elif callable(best_angles):
        T, lookup = best_angles.T, best_angles
    else:
        T, lookup = best_angles.shape[2], (lambda i, js: best_angles[i, js])
    state = np.zeros((N, T))
    for k in _host.LADDER:
        num_active = int(np.count_nonzero(final_mask))
        if not (k == 1 or 5 * k < num_active):                                     # :1083
            continue
        if verbose:
            print(f"Working on subgroups with k={k} ({num_active} candidates left) {' ' * 10}", end="\r")
        d = int(N // k)
        for step in range(int(k)):
            if step == k - 1:
                _l = len(range(d * step, num_active))                              # :1093-1094 (quirk kept)
            else:
                _l = len(range(d * step, int(d * (step + 1))))
            if _l <= 1:
                continue
            base = d * step
            matches = set()
            for i_rel in range(_l):
                i = base + i_rel
                lo, hi = i + 1, base + _l
                if lo >= hi:
                    continue
                fresh = ~cached[i, lo:hi]
                hits = np.flatnonzero(fresh & similar[i, lo:hi])
                stop = hits[0] if hits.size else hi - lo                            # first uncached similar pair
                visited = np.flatnonzero(fresh[:stop + 1]) if hits.size else np.flatnonzero(fresh)
                if visited.size == 0:
                    continue
                js = lo + visited
                if T:                                                               # in-place mutation (:1004-1008)
                    state[js] = (lookup(i, js) + state[i]) % 360.0
                if hits.size:
                    cached[i, js[:-1]] = True                                       # :1123-1125
                    matches.add((i_rel, int(js[-1] - base)))                        # :1119-1120
                else:
                    cached[i, js] = True
            for r in _cluster_rejects(matches):                                    # :1136-1152
                final_mask[r + base] = 0
    return final_mask, state


def ladder_replay_scan(first_hit, N, best_angles=None, verbose=False, native=None):
    """The grouping loop (torsion_module.py:1076-1152) driven by first_hit[i] = first later structure
    similar to i (N if none) instead of a similarity matrix.  Equivalent to ladder_replay: the pairs row
    i has cached (visited and found dissimilar) always form a prefix (i, reach[i]] of its columns, so a
    visit of row i in a chunk ending at `hi` touches the new columns (reach[i], min(first_hit, hi) - 1]
    — cached from then on — plus first_hit[i] itself when it lies inside the chunk (similar pairs are
    never cached: the reference re-evaluates and re-mutates them in every round).
    The per-row part of a chunk runs in C on the host (tsc_host_rotcorr_chunk) when `best_angles` comes
    from RotCorrPruner.scan(); the Python form below is the same loop and serves as its check.
    Returns (final_mask, state (N, T) degrees)."""
    import ctypes
    import networkx as nx
    first_hit = np.ascontiguousarray(first_hit, dtype=np.int64)
    final_mask = np.ones(N, dtype=bool)
    reach = np.arange(N, dtype=np.int64)
    T = best_angles.T if best_angles is not None else 0
    state = np.zeros((N, max(T, 0)))
    if native is None:
        native = best_angles is None or hasattr(best_angles, "compact")
    if native:
        L = lib()
        if T:
            compact = np.ascontiguousarray(best_angles.compact, dtype=np.uint64)
            if compact.size == 0:
                compact = np.zeros(1, np.uint64)
            off = best_angles.off
            table = np.ascontiguousarray(best_angles.table, dtype=np.float64)
            assert table.shape[1] == MAX_ANG
        else:                                   # no rotor states (the TFD pruning runs the same loop)
            compact, off, table = np.zeros(1, np.uint64), np.zeros(max(N, 1), np.int64), np.zeros((1, MAX_ANG))
        mi, mj = np.empty(max(N, 1), np.int32), np.empty(max(N, 1), np.int32)
        rej = np.empty(max(N, 1), np.int32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        p_first, p_reach, p_state, p_compact, p_off, p_table, p_mi, p_mj, p_rej = (vp(a) for a in (
            first_hit, reach, state, compact, off, table, mi, mj, rej))      # the arrays never move
    if native and not verbose and N > 0:
        # the whole loop below in native code (capi.cu: tsc_host_ladder_replay; the Python loop spent 25 us per chunk)
        ladder = np.ascontiguousarray(_host.LADDER, dtype=np.int64)
        mask8 = np.ones(N, dtype=np.uint8)
        scratch = np.empty(3 * N, dtype=np.int32)
        rc = int(L.tsc_host_ladder_replay(N, vp(ladder), int(ladder.size), 5, p_first, p_reach, p_state, T, p_compact, p_off,
                                          p_table, vp(mask8), vp(scratch)))
        if rc < 0:
            raise RuntimeError("tsc_host_ladder_replay: bad arguments")
        return mask8.astype(bool), state
    for k in _host.LADDER:
        num_active = int(np.count_nonzero(final_mask))
        if not (k == 1 or 5 * k < num_active):                                     # :1083
            continue
        if verbose:
            print(f"Working on subgroups with k={k} ({num_active} candidates left) {' ' * 10}", end="\r")
        d = int(N // k)
        for step in range(int(k)):
            if step == k - 1:
                _l = len(range(d * step, num_active))                              # :1093-1094 (quirk kept)
            else:
                _l = len(range(d * step, int(d * (step + 1))))
            if _l <= 1:
                continue
            base = d * step
            hi = base + _l
            matches = set()
            if native:
                n = int(L.tsc_host_rotcorr_chunk(base, hi, p_first, p_reach, p_state, T, p_compact, p_off, p_table,
                                                 p_mi, p_mj))
                if n:
                    # survivor choice of the chunk (:1136-1152) in native code as well: CPython's set / dict orders
                    # restated exactly (capi.cu: tsc_host_cluster_rejects; checked against networkx in the tests)
                    nr = int(L.tsc_host_cluster_rejects(p_mi, p_mj, n, _l, p_rej))
                    if nr < 0:
                        raise RuntimeError("tsc_host_cluster_rejects: bad arguments")
                    final_mask[base + rej[:nr]] = False
                continue
            else:
                fh = first_hit[base:hi]
                new_hi = np.minimum(fh - 1, hi - 1)
                todo = np.flatnonzero((new_hi > reach[base:hi]) | (fh < hi))
                for i_rel in todo.tolist():
                    i = base + i_rel
                    p = int(fh[i_rel])
                    nh = int(new_hi[i_rel])
                    lo = int(reach[i]) + 1
                    hit = p < hi
                    if nh >= lo:
                        js = np.arange(lo, nh + 1)
                        if hit:
                            js = np.append(js, p)
                        reach[i] = nh
                    elif hit:
                        js = np.array([p])
                    else:
                        continue
                    if T:                                                           # in-place mutation (:1004-1008)
                        state[js] = (best_angles(i, js) + state[i]) % 360.0
                    if hit:
                        matches.add((i_rel, p - base))                              # :1119-1120
            if not matches:
                continue
            for r in _cluster_rejects(matches):                                    # :1136-1152
                final_mask[r + base] = 0
    return final_mask, state


def centre_structures(structures):
    """`np.array([s - s.mean(axis=0) for s in structures])` (torsion_module.py:1023), bit-identical: for a regular
    (N, A, 3) input the library's host helper (tsc_host_centre: rows added in numpy's order, sums divided by A; ~6x
    faster than numpy on 20 000 x 63 atoms, where this statement was a third of the whole call), otherwise numpy."""
    structures = np.asarray(structures, dtype=np.float64)
    if structures.ndim == 3 and structures.shape[0] and structures.shape[1] and structures.shape[2] == 3:
        src = np.ascontiguousarray(structures)
        out = np.empty_like(src)
        rc = lib().tsc_host_centre(src.ctypes.data, src.shape[0], src.shape[1], out.ctypes.data,
                                   min(os.cpu_count() or 1, 8))
        if rc != 0:
            raise RuntimeError(f"tsc_host_centre failed ({rc})")
        return out
    return np.array([s - s.mean(axis=0) for s in structures])


def prune_conformers_rmsd_rot_corr(structures, atomnos, graph, max_rmsd=0.25, verbose=False, logfunction=None,
                                   *, torsion_info: TorsionInfo | None = None, max_structures=750, mode=None,
                                   group=None, rank=None, world=None):
    """Drop-in for tscode.torsion_module.prune_conformers_rmsd_rot_corr (:1013-1161).

    mode "exact" (default up to 2000 structures): row-by-row replay from the current, mutated
    coordinates — same visiting order, same mutations, same returned structures as the reference.
    mode "stateless" (default above): pairs evaluated from the centred input — a forward scan per row
    up to its first similar partner (mode "allpairs": every pair) — + host replay with rotor-state algebra; masks agree with the reference on every fixture, but for rotors whose
    n-fold images differ only at noise level (e.g. a methyl-capped alkyne: the heavy atoms sit on
    the axis) the hydrogens of the returned structures may end up in another image."""
    structures = centre_structures(structures)                                                       # :1023
    atomnos = np.asarray(atomnos)
    N = structures.shape[0]
    final_mask = np.ones(N, dtype=bool)
    if N == 0:
        return structures, final_mask
    info = torsion_info if torsion_info is not None else perceive_torsions(structures[0], atomnos, graph)
    if info.T == 0 or (max_structures is not None and N > max_structures):                            # :1056-1060
        return structures[final_mask], final_mask
    if logfunction is not None:                                                                       # :1063-1074
        logfunction('\n >> Dihedrals considered for subsymmetry corrections:')
        for i, (torsion, angle) in enumerate(zip(info.torsions, info.angles)):
            sym = ''.join(_SYMBOLS[atomnos[a]] if atomnos[a] < len(_SYMBOLS) else '?' for a in torsion)
            logfunction(' {:2s} - {:21s} : {} : {}-fold'.format(str(i), str(list(torsion)), sym, len(angle)))
        logfunction("\n")
    if mode is None:
        mode = "exact" if N <= 2000 else "stateless"
    if mode == "exact":
        pr = RotCorrPruner(structures, atomnos, info, max_rmsd, want_codes=False)
        out, mask, _ = pr.prune_stateful(verbose=verbose)
        return out, mask
    if mode == "allpairs":             # every pair evaluated (cross-check of the scan; O(N^2) whatever the data)
        pr = RotCorrPruner(structures, atomnos, info, max_rmsd)
        pr.similarity()
        mask, state = ladder_replay(pr.similar_matrix(), N, pr.best_angles(), verbose=verbose)
    else:
        if group is not None or world is not None:
            import torch.distributed as dist
            world = dist.get_world_size(group) if world is None else int(world)
            rank = dist.get_rank(group) if rank is None else int(rank)
        else:
            rank, world = 0, 1
        pr = RotCorrPruner(structures, atomnos, info, max_rmsd, want_codes=False)
        first_hit, lookup = pr.scan(rank, world, group)       # rows dealt to the ranks; the replay runs on every rank
        mask, state = ladder_replay_scan(first_hit, N, lookup, verbose=verbose)
    keep = np.flatnonzero(mask)
    out = pr.apply_states(keep, state[keep])
    return out, mask


def rotationally_corrected_rmsd(ref, coord, atomnos, torsions, graph, angles, *, torsion_info=None):
    """Drop-in for tscode.torsion_module.rotationally_corrected_rmsd (:953-1011), including its
    side effect: `coord` is rotated IN PLACE to the best rotor angles.  `graph` is only needed
    when torsion_info is not given (masks / node lists are then taken from the reference's
    helpers)."""
    ref = np.asarray(ref, dtype=np.float64)
    atomnos = np.asarray(atomnos)
    if torsion_info is None:
        import networkx as nx
        from tscode.torsion_module import _get_rotation_mask
        A = len(atomnos)
        rot, nodes = np.zeros((len(torsions), A), bool), np.zeros((len(torsions), A), bool)
        for k, t in enumerate(torsions):
            for o in torsions:
                if o is not t:
                    graph.remove_edge(o[1], o[2])
            comp = [s for s in nx.connected_components(graph) if t[1] in s][0]
            for o in torsions:
                if o is not t:
                    graph.add_edge(o[1], o[2])
            nodes[k, [i for i in comp if atomnos[i] != 1]] = True
            rot[k] = _get_rotation_mask(graph, t)
        torsion_info = TorsionInfo([tuple(t) for t in torsions], [tuple(a) for a in angles], rot, nodes)
    pr = RotCorrPruner(np.stack([ref, np.asarray(coord, dtype=np.float64)]), atomnos, torsion_info, 1e300,
                       want_codes=True, want_rmsd=True)
    pr.similarity()
    r = float(pr.rmsd[0, 1].item())
    best = pr.best_angles()(0, np.array([1]))
    coord[...] = pr.apply_states([1], best)[0]
    return r
