"""B200-native drop-in for tscode/rmsd_pruning.py.

Same call signatures and return values as the reference:

    prune_conformers_rmsd(structures, atomnos, rmsd_thr=0.5) -> (structures[mask], mask)   # :164-206
    rmsd_and_max_numba(p, q) -> (rmsd, max_deviation)                                      # :6-41
    _rmsd_similarity(ref, structures, rmsd_thr=0.5) -> bool                                # :208-224

plus the device-resident, batched form (`RmsdPruner`) that bench.py and multi-GPU runs use.
All arithmetic runs in hand-written sm_100a kernels behind the C-ABI of
include/tscode_b200.h; torch only owns device memory, streams and the process group.
There is no CPU fallback: without the built extension or a CUDA device every call raises.
"""
from __future__ import annotations

import numpy as np

from . import _host
from ._lib import check, lib, ptr, require_cuda, stream_ptr

import ctypes
import functools

# screen = tcgen05 / TMEM pre-screen on FP16 operands (rmsd_screen.cu; default); dmma / fma = FP64 tensor cores /
# FMA pipe (rmsd_sim.cu: the north star's FP64 variants, and the fallback above tsc_screen_max_atoms heavy atoms)
VARIANTS = {"dmma": 0, "fma": 1, "screen": 5}


import threading

_STAGING = {"lock": threading.Lock(), "buf": None, "event": None}


class _Staging:
    """The process-wide pinned float64 staging buffer for pageable inputs (cudaHostAlloc costs tens of ms: paid once,
    grown on demand), handed out to ONE upload at a time: acquire() takes the lock and waits for the DMA of the
    previous user to drain (its event), release(event) records the new user's last copy.  Two pruners built before
    either runs, or two threads pruning at once, therefore never share live staging memory."""

    @staticmethod
    def acquire(numel):
        import torch
        _STAGING["lock"].acquire()
        try:
            if _STAGING["event"] is not None:
                _STAGING["event"].synchronize()
                _STAGING["event"] = None
            buf = _STAGING["buf"]
            if buf is None or buf.numel() < numel:
                buf = torch.empty(numel, dtype=torch.float64).pin_memory()
                _STAGING["buf"] = buf
            return buf[:numel]
        except BaseException:
            _STAGING["lock"].release()
            raise

    @staticmethod
    def release(event):
        _STAGING["event"] = event
        _STAGING["lock"].release()


@functools.lru_cache(maxsize=8)
def _copy_stream(device_str):
    import torch
    return torch.cuda.Stream(device=torch.device(device_str))


UPLOAD_CHUNKS = 8       # pieces a host ensemble is uploaded in (_upload_pack_screen_pipelined)
HOST_GATHER_SHARE = 0.45     # part of structures[mask] taken from the caller's host array while the rest comes over PCIe


class _PeerLists:
    """Several ranks: the confirmed-pair lists are exchanged WITHOUT a collective.  One symmetric-memory allocation per
    (group, world, list capacity) holds two list arrays (world blocks each) and two rows of flags; every rank maps every
    peer's copy over NVLink (torch.distributed._symmetric_memory) and, after its verify kernels, pushes its own list into
    block `rank` of every rank's array with plain stores, sized on the device from the list's header, then raises its
    flag on every rank (eliminate.cu: tsc_pairs_push); the ladder kernel of every rank waits for all flags
    (tsc_elim_fused_p2p).  The arrays alternate from call to call because a rank can be one call ahead of its peers.
    Creation is collective (rendezvous + an agreement on whether it worked); if symmetric memory is not available on any
    rank, all ranks use the NCCL all-gather instead."""
    _cache = {}
    enabled = True

    @classmethod
    def get(cls, group, rank, world, stride, device):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return None
        g = group if group is not None else dist.group.WORLD
        key = (id(g), world, int(stride), str(device))
        obj = cls._cache.pop(key, None)
        if obj is None:
            obj = cls(g, rank, world, int(stride), device)
        cls._cache[key] = obj                             # most recently used last
        while len(cls._cache) > 4:                        # (the same sequence on every rank: constructors run in step)
            cls._cache.pop(next(iter(cls._cache)))
        return obj if obj.ok else None

    def __init__(self, group, rank, world, stride, device):
        import torch
        import torch.distributed as dist
        self.ok, self.err, self.calls = False, None, 0
        n_list = world * stride * 2                       # int32 words of one list array
        try:
            if not self.enabled:
                raise RuntimeError("disabled")
            import torch.distributed._symmetric_memory as symm
            with torch.cuda.device(device):
                self.buf = symm.empty(2 * n_list + 128, dtype=torch.int32, device=device)
                self.buf.zero_()
                torch.cuda.current_stream().synchronize()
                self.hdl = symm.rendezvous(self.buf, group)
                ptrs = [int(p) for p in self.hdl.buffer_ptrs]
                assert len(ptrs) == world
                self.peer_lists = [torch.tensor([p + 4 * par * n_list for p in ptrs], dtype=torch.int64, device=device)
                                   for par in (0, 1)]
                self.peer_flags = [torch.tensor([p + 4 * (2 * n_list + 64 * par) for p in ptrs], dtype=torch.int64,
                                                device=device) for par in (0, 1)]
                self.lists = [self.buf[par * n_list:(par + 1) * n_list].view(world * stride, 2) for par in (0, 1)]
                self.flags = [self.buf[2 * n_list + 64 * par:2 * n_list + 64 * par + 64] for par in (0, 1)]
            good = 1
        except Exception as e:                            # no symmetric memory here: every rank falls back together
            self.err = repr(e)
            good = 0
        t = torch.tensor([good], dtype=torch.int32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)       # (also orders the zeroing before any peer write)
        self.ok = bool(int(t.item()))


def _upload_bounds(N, n_chunks=UPLOAD_CHUNKS):
    """Panel boundaries of the upload chunks: chunk c = panels [b[c], b[c+1])."""
    n_panels = (N + 127) // 128
    n_chunks = max(1, min(n_chunks, n_panels))
    return [n_panels * c // n_chunks for c in range(n_chunks)] + [n_panels]


@functools.lru_cache(maxsize=8)
def _work_lists(N, rank, world, variant, device_str, n_ctas=0, tile_j=32):
    """Device-resident work lists (owned row blocks, tile / item lists) — they depend only on the
    shape of the problem, so repeated prunes of same-sized ensembles reuse them."""
    import torch
    dev = torch.device(device_str)
    rb = _host.owned_row_blocks(N, rank, world)
    out = {"row_blocks_np": rb, "row_blocks": torch.from_numpy(rb).to(dev)}
    if variant == 5:
        # default screen (rmsd_screen.cu): j tiles of 32 conformers; one contiguous, equally expensive stretch of
        # (panel, j tile) pairs per CTA of the persistent grid, plus the same per upload chunk
        n_ctas = n_ctas if n_ctas > 0 else torch.cuda.get_device_properties(dev).multi_processor_count
        items = np.ascontiguousarray(_host.build_screen_items(N, rb, n_ctas, tile_j=tile_j))
        out["n_items"], out["items"], out["items_np"] = int(items.shape[0]), torch.from_numpy(items).to(dev), items
        bounds = _upload_bounds(N)
        chunks = [np.ascontiguousarray(_host.build_screen_items(N, rb, n_ctas, panel_lo=bounds[c], panel_hi=bounds[c + 1],
                                                                tile_j=tile_j))
                  for c in range(len(bounds) - 1)]
        out["chunk_items"] = [(torch.from_numpy(it).to(dev) if it.shape[0] else None, int(it.shape[0])) for it in chunks]
    else:
        tiles = _host.build_tiles(N, rb)
        out["n_tiles"], out["tiles"] = int(tiles.shape[0]), torch.from_numpy(tiles).to(dev)
    return out


class RmsdPruner:
    """All-pairs Kabsch similarity + k-ladder elimination for one ensemble, resident in HBM.

    structures : (N, A, 3) float64, numpy array or torch tensor (host or device)
    atomnos    : (A,) ints; hydrogens (== 1) are ignored        (rmsd_pruning.py:178-179)
    variant    : "screen" (default) = tcgen05 / TMEM pre-screen on FP16 operands (10-bit mantissa, K = 16 atoms per
                 MMA, FP32 accumulation) with a rigorous operand error bound and FP32 exclusion tests (Samuelson's
                 bound; sign test of the key-matrix quartic), exact FP64 verification of everything it cannot
                 exclude; falls back to "dmma" above tsc_screen_max_atoms heavy atoms.  "dmma" = FP64 tensor cores,
                 "fma" = FP64 FMA pipe.  All give identical final similarity bits and masks.
    screen_mode: form of the default screen (0 = Samuelson-type bound only, on 48-wide tiles; 1 = the bound, then
                 the FP32 quartic test where it left a pair undecided; 2 = quartic for every pair; 3 = mode 0 on
                 64-wide tiles); None = chosen from the first structure and a sample of pairs (_host.screen_plan_native).
    screen_frame: rotate the ensemble into the principal axes of the first structure and weight the column-side
                 operand (_host.screen_frame; rmsd_screen.cu, ScFrame), which makes the bound as sharp for elongated
                 and planar molecules as it is for isotropic ones.  Both are speed choices only: every form is
                 conservative and the final bits are exact.
    rank/world/group : row-block sharding over one process per GPU (block-cyclic);
                 every rank holds the whole packed ensemble, computes the similarity rows it
                 owns, and per elimination round contributes its rows' verdicts to an NCCL
                 all-gather.
    """

    def __init__(self, structures, atomnos, rmsd_thr=0.5, *, variant="screen", device=None,
                 rank=0, world=1, group=None, grid_ctas=0, ladder="fused", pair_cap=None, cand_cap=None,
                 pipeline_upload=True, screen_mode=None, screen_frame=True):
        torch = require_cuda()
        self.torch = torch
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.thr = float(rmsd_thr)
        self.variant = VARIANTS[variant] if isinstance(variant, str) else int(variant)
        self.variant_requested = self.variant
        self.rank, self.world, self.group = int(rank), int(world), group
        self.grid_ctas = int(grid_ctas)
        self.pace = 0                     # tsc_rmsd_screen's MMA spacing (measurement aid: tools/screen_check.py)
        self.screen_mode = screen_mode    # form of the default screen: None = chosen from the molecule's shape (below)
        self.frame = None                 # 12 float64 (Q, t) given to tsc_pack_screen / tsc_rmsd_screen; None = identity
        self._verify_progress = None      # device int32: candidate-list entries already verified (pipelined upload)
        self._info_host = None            # pinned landing buffer of the fused ladder's status words
        self._verify_incremental = False
        if ladder not in ("fused", "bitrows"):
            raise ValueError("ladder must be 'fused' or 'bitrows'")
        self.ladder = ladder
        atomnos = np.asarray(atomnos)
        heavy = np.flatnonzero(atomnos != 1).astype(np.int32)
        self._host = None
        if torch.is_tensor(structures):
            src = structures
        else:
            src = torch.as_tensor(np.ascontiguousarray(structures, dtype=np.float64))
        if src.dim() != 3 or src.shape[2] != 3 or src.shape[1] != atomnos.shape[0]:
            raise ValueError(f"structures must be (N, {atomnos.shape[0]}, 3), got {tuple(src.shape)}")
        self._needs_staging = False
        if (pipeline_upload and not src.is_cuda and src.dtype == torch.float64 and src.is_contiguous()
                and src.shape[0] >= 4096):
            # host input: the copy is issued in chunks by run(), overlapped with pack and screen.  Pageable memory
            # (what the reference's callers hold: np.array(poses)) goes through a cached pinned staging buffer,
            # chunk by chunk, so that the host memcpy of one chunk overlaps the DMA of the previous one.
            self._host = src
            self._needs_staging = not src.is_pinned()
            if self.world > 1:
                # several ranks, replicated host input: every rank uploads ONE slice of rows and the slices are
                # all-gathered over NVLink (_upload_sharded) instead of `world` full uploads from the same host
                per = (int(src.shape[0]) + self.world - 1) // self.world
                self._S_all = torch.empty((self.world * per,) + tuple(src.shape[1:]), dtype=torch.float64, device=self.device)
                S = self._S_all[:src.shape[0]]
            else:
                S = torch.empty(src.shape, dtype=torch.float64, device=self.device)
        else:
            S = src.to(self.device, dtype=torch.float64).contiguous()
        self.S = S
        self.N, self.A, self.M = int(S.shape[0]), int(S.shape[1]), int(heavy.size)
        N, M = self.N, self.M
        self.nb_pad = _host.num_blocks_padded(N)
        self.W = self.nb_pad
        if self.variant == 5 and N and M:
            # principal-axes frame and column weights of the first structure (rmsd_screen.cu, ScFrame), and from its
            # shape the form of the screen (_host.screen_mode_for).  Speed decisions: every choice is conservative.
            si, sj = _host.sample_pair_indices(N)
            h32 = heavy.astype(np.int32)
            if not src.is_cuda and src.dtype == torch.float64 and src.is_contiguous():
                plan = _host.screen_plan_native(src.numpy(), h32, self.thr, 0, si, sj)      # on the caller's memory
            else:
                rows = torch.from_numpy(np.concatenate([[0], si, sj]).astype(np.int64))
                smp = src.index_select(0, rows.to(src.device)).to("cpu", torch.float64).contiguous().numpy()
                k = np.arange(si.size, dtype=np.int64)
                plan = _host.screen_plan_native(smp, h32, self.thr, 0, 1 + k, 1 + si.size + k)
            self.frame, mode, self.sample_undecided = plan
            if not screen_frame:                             # (comparison runs: plain Samuelson in the frame as given)
                ratio = _host.screen_frame((src[0].cpu() if src.is_cuda else src[0]).numpy()[heavy])[1]
                self.frame, mode = None, (0 if ratio <= 1.03 else 2)
            if self.screen_mode is None:
                self.screen_mode = mode
            self.tile_j = {0: 48, 3: 64}.get(self.screen_mode, 32)
            if M > int(lib().tsc_screen_max_atoms(self.tile_j)) and self.screen_mode in (0, 3):
                self.screen_mode, self.tile_j = 1, 32        # too many atoms for the wide tiles: Samuelson, then quartic
        else:
            self.tile_j = 32
        if self.variant == 5 and M > int(lib().tsc_screen_max_atoms(self.tile_j)):
            self.variant = 0             # documented fallback: FP64 tensor cores (include/tscode_b200.h)
        with torch.cuda.device(self.device):
            dev = self.device
            self.heavy_idx = torch.from_numpy(heavy).to(dev)
            wl = _work_lists(N, self.rank, self.world, self.variant, str(dev), max(self.grid_ctas, 0), self.tile_j)
            self.row_blocks_np, self.row_blocks = wl["row_blocks_np"], wl["row_blocks"]
            self.n_rb = int(self.row_blocks_np.size)
            self.n_tiles, self.tiles = wl.get("n_tiles", 0), wl.get("tiles")
            self.n_items, self.items = wl.get("n_items", 0), wl.get("items")
            self.packed = torch.empty(max(_host.packed_doubles(N, max(M, 1)), 1), dtype=torch.float64, device=dev)
            n_g = max(self.nb_pad * _host.CB, _host.screen_rows_padded(N))
            self.G = torch.empty(n_g, dtype=torch.float64, device=dev)
            if self.variant == 5:
                L = lib()
                nbytes = max(int(L.tsc_screen_operand_bytes(N, max(M, 1))), 16)
                self.PA = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                self.PB = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                self.PR = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                self.sG = torch.empty(n_g, dtype=torch.float64, device=dev)
                self.G_side = torch.empty(n_g, dtype=torch.float64, device=dev)
                self.CT = torch.empty(max(L.tsc_screen_ct_floats(N), 1), dtype=torch.float32, device=dev)
            self.sim_bits = torch.empty((max(self.n_rb, 1) * _host.CB, self.W), dtype=torch.int32, device=dev)
            self.stats = torch.zeros(4, dtype=torch.int64, device=dev)
            nw = (N + 31) // 32
            rows_pad = self.nb_pad * _host.CB
            self.active = torch.empty(max(nw, 1), dtype=torch.int32, device=dev)
            self.cachebits = torch.empty(max(nw, 1), dtype=torch.int32, device=dev)
            self.mask_bytes = torch.ones(rows_pad, dtype=torch.uint8, device=dev)
            self.row_state = torch.full((rows_pad,), -1, dtype=torch.int32, device=dev)
            self.hist = torch.zeros(len(_host.LADDER) + 2, dtype=torch.int32, device=dev)   # active count per round
            self.key_first = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
            self.key_second = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
            self.n_keys = torch.zeros(1, dtype=torch.int32, device=dev)
            # confirmed-pair list of this rank (header + capacity pairs) and, with several ranks, the
            # gathered lists of all of them; consumed by the fused ladder (eliminate.cu)
            cap = int(pair_cap) if pair_cap is not None else 32 * N // self.world + 4096
            self.pair_stride = cap + 1
            self.pair_list = torch.zeros((self.pair_stride, 2), dtype=torch.int32, device=dev)
            # peer-written lists (no collective) when symmetric memory is available; else the NCCL all-gather, whose
            # landing buffers are only allocated if it is ever used (_enqueue_fused)
            self._peer = (_PeerLists.get(self.group, self.rank, self.world, self.pair_stride, dev)
                          if (self.world > 1 and self.ladder == "fused") else None)
            self.pair_stride_small = min(self.pair_stride, 8 * N // self.world + 2048 + 1)
            self.pair_all = self.pair_all_small = self.pair_list if self.world == 1 else None
            # candidate list the tcgen05 screens append to (local row, j); verify works from it
            self.cand_stride = (int(cand_cap) if cand_cap is not None else 64 * N // self.world + 8192) + 1
            self.cand_list = torch.zeros((self.cand_stride, 2), dtype=torch.int32, device=dev)
            L = lib()
            self.fused_ws = torch.empty(int(L.tsc_elim_fused_ws_words(N)), dtype=torch.int32, device=dev)
            self.fused_out = torch.zeros(int(L.tsc_elim_fused_out_bytes(N)), dtype=torch.uint8, device=dev)
            self._info_off = (N + 3) // 4 * 4
            if self.world > 1:
                self._init_shards()
        self._cands = []
        self._packed_event = None
        self._fused_enqueued = None
        self._pairs_ready = False
        self._rounds_fused = None
        self.ladder_used = None
        self.packed_ready = False

    # ---- multi-GPU plumbing -------------------------------------------------------------------
    def _init_shards(self):
        torch = self.torch
        n_rb_all = [_host.owned_row_blocks(self.N, r, self.world).size for r in range(self.world)]
        self.n_rb_max = max(max(n_rb_all), 1)
        L = self.n_rb_max * _host.CB
        rows_pad = self.nb_pad * _host.CB
        gidx = np.full((self.world, L), rows_pad, dtype=np.int64)     # rows_pad = scratch slot
        for r in range(self.world):
            g = _host.global_rows_of(_host.owned_row_blocks(self.N, r, self.world))
            gidx[r, :g.size] = g
        dev = self.device
        self.gather_index = torch.from_numpy(gidx.reshape(-1)).to(dev)
        self.my_rows = torch.from_numpy(np.minimum(gidx[self.rank], rows_pad - 1)).to(dev)
        self.loc_state = torch.full((L,), -2, dtype=torch.int32, device=dev)
        self.all_state = torch.empty(self.world * L, dtype=torch.int32, device=dev)
        self.state_scratch = torch.full((rows_pad + 1,), -2, dtype=torch.int32, device=dev)

    def _exchange_round(self):
        """All-gather this round's per-row states (verdict + emitted key in one int32) — NCCL over NVLink."""
        import torch.distributed as dist
        n_mine = self.n_rb * _host.CB
        rows_pad = self.nb_pad * _host.CB
        self.loc_state[:n_mine] = self.row_state[self.my_rows[:n_mine]]
        dist.all_gather_into_tensor(self.all_state, self.loc_state, group=self.group)
        self.state_scratch.index_copy_(0, self.gather_index, self.all_state)
        self.row_state.copy_(self.state_scratch[:rows_pad])

    # ---- phases -------------------------------------------------------------------------------
    def pack(self):
        if self._host is not None:                       # phases driven by hand: plain upload first
            self.S.copy_(self._host)
            self._host = None
        if self.N == 0 or self.M == 0:
            return
        L = lib()
        torch = self.torch
        with torch.cuda.device(self.device):
            if self.variant == 5:
                # the FP64 tiled-SoA image is only read by verify: it is written on a side stream while the main
                # stream goes on to the FP16 images and the screen (verify waits for the event)
                main, side = torch.cuda.current_stream(), _copy_stream(str(self.device))
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    check(L.tsc_pack(ptr(self.S), self.N, self.A, ptr(self.heavy_idx), self.M, ptr(self.packed),
                                     ptr(self.G_side), stream_ptr()), "tsc_pack")      # (G itself comes from tsc_pack_screen)
                    self._packed_event = torch.cuda.Event()
                    self._packed_event.record(side)
            else:
                check(L.tsc_pack(ptr(self.S), self.N, self.A, ptr(self.heavy_idx), self.M, ptr(self.packed),
                                 ptr(self.G), stream_ptr()), "tsc_pack")
            if self.variant == 5:
                check(L.tsc_pack_screen(ptr(self.S), self.N, self.A, ptr(self.heavy_idx), self.M, ptr(self.PA),
                                        ptr(self.PB), ptr(self.PR), ptr(self.G), ptr(self.sG), ptr(self.CT), 0, 0,
                                        self.tile_j, self._frame_ptr(), stream_ptr()), "tsc_pack_screen")
        self.packed_ready = True

    def _frame_ptr(self):
        return None if self.frame is None else self.frame.ctypes.data

    def screen(self):
        """All-pairs contraction + closed-form screen (the hot kernel)."""
        if (self.n_tiles == 0 and self.n_items == 0) or self.M == 0:
            return
        if not self.packed_ready:
            self.pack()
        self._pairs_ready = False
        L = lib()
        with self.torch.cuda.device(self.device):
            self.stats.zero_()
            self.cand_list[0].fill_(0 if self.variant == 5 else -1)      # -1: this screen writes no list
            if self.variant == 5:
                check(L.tsc_rmsd_screen(ptr(self.PA), ptr(self.PB), ptr(self.PR), ptr(self.G), ptr(self.sG),
                                        ptr(self.CT), self.N, self.M, ptr(self.items), self.n_items, self.thr,
                                        ptr(self.sim_bits), ptr(self.cand_list), self.cand_stride, self.grid_ctas,
                                        self.screen_mode, self.pace, self._frame_ptr(), stream_ptr()), "tsc_rmsd_screen")
            else:
                check(L.tsc_rmsd_sim_tiles(ptr(self.packed), ptr(self.G), self.N, self.M, ptr(self.tiles),
                                           self.n_tiles, self.thr, ptr(self.sim_bits), self.variant, self.grid_ctas,
                                           stream_ptr()), "tsc_rmsd_sim_tiles")

    def verify(self):
        """Exact re-evaluation of screened pairs; afterwards sim_bits are final and the confirmed
        pairs of the owned rows are also available as an (i, j) list (pair_list)."""
        L = lib()
        with self.torch.cuda.device(self.device):
            if self._packed_event is not None:           # FP64 image written on the side stream (pack())
                self.torch.cuda.current_stream().wait_event(self._packed_event)
                self._packed_event = None
            if self._verify_incremental:                 # the pipelined upload verified as it went: only the rest
                self._verify_incremental = False
                if self.n_rb and self.M:
                    check(L.tsc_rmsd_verify_incr(ptr(self.packed), self.N, self.M, ptr(self.row_blocks), self.n_rb,
                                                 self.thr, ptr(self.sim_bits), ptr(self.stats), ptr(self.pair_list),
                                                 self.pair_stride, ptr(self.cand_list), self.cand_stride,
                                                 ptr(self._verify_progress), 1, stream_ptr()), "tsc_rmsd_verify_incr")
                self._pairs_ready = True
                return
            self.pair_list[0].zero_()
            if self.n_rb and self.M:
                check(L.tsc_rmsd_verify(ptr(self.packed), self.N, self.M, ptr(self.row_blocks), self.n_rb, self.thr,
                                        ptr(self.sim_bits), ptr(self.stats), ptr(self.pair_list), self.pair_stride,
                                        ptr(self.cand_list), self.cand_stride, stream_ptr()), "tsc_rmsd_verify")
        self._pairs_ready = True

    def similarity(self):
        self.screen()
        self.verify()

    def _gate_ptr(self, r):
        import ctypes
        return ctypes.c_void_p(self.hist.data_ptr() + 4 * r)

    def eliminate(self):
        """The k-ladder (rmsd_pruning.py:186-204); returns the boolean mask as a device tensor.

        Default: ONE persistent cooperative kernel over the confirmed-pair lists verify() emitted
        (all-gathered once when several ranks share the rows).  If a list overflowed its capacity
        (a very redundant ensemble), or the pruner was built with ladder="bitrows", the bit-row
        kernels run instead: three launches and, on several ranks, one all-gather per round.
        Both give the same mask (tests/test_gpu_parity.py)."""
        torch = self.torch
        N = self.N
        if N == 0:
            return torch.zeros(0, dtype=torch.bool, device=self.device)
        if self.M == 0:
            raise ValueError("prune_conformers_rmsd needs at least one non-hydrogen atom")
        if self.ladder == "fused" and self._pairs_ready:
            mask = self._eliminate_fused()
            if mask is not None:
                return mask
        return self._eliminate_bitrows()

    def _enqueue_fused(self, tier="small"):
        """All-gather of the confirmed-pair lists (several ranks) + the fused ladder kernel, nothing read back.
        Two tiers: the blocks have room for 32 N / world pairs, but a typical ensemble has a few per structure, so
        first only a short prefix of every block (8 N / world + 2048 pairs) is gathered; if some rank's count does
        not fit the prefix the kernel reports status 1 and the full blocks are gathered (same decision on every
        rank: all see the same headers)."""
        torch = self.torch
        L = lib()
        stride = self.pair_stride_small if (tier == "small" and self.world > 1) else self.pair_stride
        with torch.cuda.device(self.device):
            if self.world > 1 and self._peer is not None:
                pl = self._peer
                pl.calls += 1
                epoch, par = pl.calls, pl.calls & 1
                check(L.tsc_pairs_push(ptr(self.pair_list), self.pair_stride, ptr(pl.peer_lists[par]),
                                       ptr(pl.peer_flags[par]), self.rank, self.world, epoch, stream_ptr()), "tsc_pairs_push")
                rc = L.tsc_elim_fused_p2p(ptr(pl.lists[par]), self.world, self.pair_stride, self.N, 20, ptr(self.fused_ws),
                                          ptr(self.fused_out), ptr(pl.flags[par]), epoch, stream_ptr())
                if rc == 1:
                    self._fused_enqueued = "unavailable"
                    return
                check(rc, "tsc_elim_fused_p2p")
                self._fused_enqueued = "full"
                return
            if self.world > 1:
                import torch.distributed as dist
                if self.pair_all is None:
                    self.pair_all = torch.zeros((self.world * self.pair_stride, 2), dtype=torch.int32, device=self.device)
                    self.pair_all_small = torch.zeros((self.world * self.pair_stride_small, 2), dtype=torch.int32,
                                                      device=self.device)
                if stride == self.pair_stride:
                    lists = self.pair_all
                    dist.all_gather_into_tensor(lists, self.pair_list, group=self.group)
                else:
                    lists = self.pair_all_small
                    dist.all_gather_into_tensor(lists, self.pair_list[:stride], group=self.group)
            else:
                lists = self.pair_list
            rc = L.tsc_elim_fused(ptr(lists), self.world, stride, self.N, 20, ptr(self.fused_ws),
                                  ptr(self.fused_out), stream_ptr())
            if rc == 1:                   # cudaErrorInvalidValue: the ensemble is too large for the fused kernel's
                self._fused_enqueued = "unavailable"      # shared-memory bitmaps (N > ~690 000): bit-row ladder instead
                return
            check(rc, "tsc_elim_fused")
        self._fused_enqueued = tier if self.world > 1 else "full"

    def _eliminate_fused(self):
        torch = self.torch
        N = self.N
        if not self._fused_enqueued:
            self._enqueue_fused()
        while True:
            tier = self._fused_enqueued
            self._fused_enqueued = None
            if tier == "unavailable":
                return None
            with torch.cuda.device(self.device):
                if self._info_host is None:
                    self._info_host = torch.empty(32, dtype=torch.int32, pin_memory=True)
                self._info_host.copy_(self.fused_out[self._info_off:self._info_off + 128].view(torch.int32), non_blocking=True)
                torch.cuda.current_stream().synchronize()                                      # the one sync of the ladder
                info = self._info_host.tolist()
            if info[0] == 1 and tier == "small":
                self._enqueue_fused("full")               # a list did not fit the short prefix
                continue
            break
        if info[0] != 0:
            # 1: a pair list overflowed; anything else: the kernel gave up at a grid barrier (bounded spin, e.g. under a
            # profiler or pre-emption).  Either way the bit-row ladder decides (same decision on every rank: all see
            # the same headers; an abort on one rank only would desynchronise the ranks' collectives, so it raises)
            if info[0] != 1 and self.world > 1:
                raise RuntimeError(f"tsc_elim_fused did not complete (status {info[0]}"
                                   + (": a peer never published its pair list)" if info[0] == 3 else ")"))
            return None
        self._rounds_fused = [int(k) for k in info[8:8 + info[1]]]
        self.ladder_used = "fused"
        return self.fused_out[:N].to(torch.bool)

    def _eliminate_bitrows(self):
        """Bit-row ladder: every candidate round is enqueued without reading anything back; each kernel
        evaluates the reference's gate on the device (eliminate.cu)."""
        torch = self.torch
        N = self.N
        self._rounds_fused = None
        self.ladder_used = "bitrows"
        L = lib()
        with torch.cuda.device(self.device):
            st = stream_ptr()
            self.row_state.fill_(-1)
            self.n_keys.zero_()
            self.hist.zero_()
            # all structures active: active words + hist[0] = N  (no gate, no keys emitted)
            check(L.tsc_elim_commit(ptr(self.row_state), N, N, 1, ptr(self.active), ptr(self.mask_bytes),
                                    ptr(self.key_first), ptr(self.key_second), ptr(self.n_keys), None,
                                    self._gate_ptr(0), st), "tsc_elim_commit(init)")
            self._cands = [int(k) for k in _host.LADDER if _host.ladder_gate(k, N)]     # n_active <= N
            for r, k in enumerate(self._cands):
                cs = _host.chunk_size(N, k)
                gate, out = self._gate_ptr(r), self._gate_ptr(r + 1)
                check(L.tsc_elim_cachebits(ptr(self.key_first), ptr(self.key_second), ptr(self.n_keys), N, cs, k,
                                           ptr(self.cachebits), gate, st), "tsc_elim_cachebits")
                if self.n_rb:
                    check(L.tsc_elim_round(ptr(self.sim_bits), ptr(self.row_blocks), self.n_rb, ptr(self.active),
                                           ptr(self.cachebits), N, cs, k, ptr(self.row_state), gate, st),
                          "tsc_elim_round")
                if self.world > 1:
                    self._exchange_round()
                check(L.tsc_elim_commit(ptr(self.row_state), N, cs, k, ptr(self.active), ptr(self.mask_bytes),
                                        ptr(self.key_first), ptr(self.key_second), ptr(self.n_keys), gate, out, st),
                      "tsc_elim_commit")
            return self.mask_bytes[:N].to(torch.bool)

    @property
    def rounds(self):
        """k of every round that actually ran (data dependent) — reads the per-round
        active counts back from the device."""
        if self._rounds_fused is not None:
            return list(self._rounds_fused)
        h = self.hist.tolist()
        return [k for r, k in enumerate(self._cands) if _host.ladder_gate(k, h[r])]

    def run_async(self):
        """Enqueue the whole prune (upload if the input was pinned host memory, pack, screen, verify, fused
        ladder) without waiting for anything; finish() returns the mask.  Lets the caller do host work — e.g.
        allocating and first-touching the output array — while the GPU runs."""
        if self.N == 0 or self.M == 0 or self.ladder != "fused":
            return
        if self._host is not None:
            self._upload_pack_screen_pipelined()
        else:
            self.pack()
            self.screen()
        self.verify()
        self._enqueue_fused()

    def finish(self):
        """Mask of a prune started with run_async() (or the whole prune if it was not)."""
        if self._fused_enqueued:
            return self.eliminate()
        return self.run()

    def run(self):
        if self._host is not None:
            self._upload_pack_screen_pipelined()
            self.verify()
            return self.eliminate()
        self.pack()
        self.similarity()
        return self.eliminate()

    def _upload_pack_screen_pipelined(self):
        """Pinned host input: H2D copy, pack and screen overlapped.  A 128-row panel only needs the conformers
        from its own first row to the end (rows i, columns j > i), so the ensemble is uploaded in chunks from the
        LAST to the first on a copy stream, and as soon as a chunk has landed its rows are packed and the work
        items of its panels are screened: when the final (first) chunk arrives only its own share of the pairs
        (~2/n_chunks of them) is still to do."""
        torch = self.torch
        host, N = self._host, self.N
        self._host = None                                # the next run() works from the device copy
        if self.world > 1 and N:
            self._upload_sharded(host)
            self.pack()
            self.screen()
            return
        if self.variant != 5 or N == 0 or self.M == 0:
            self.S.copy_(host)
            self.pack()
            self.screen()
            return
        L = lib()
        n_panels = (N + 127) // 128
        pb = _upload_bounds(N)
        n_chunks = len(pb) - 1
        bounds = [q * 128 for q in pb[:-1]] + [N]
        chunk_items = _work_lists(N, self.rank, self.world, self.variant, str(self.device),
                                  max(self.grid_ctas, 0), self.tile_j)["chunk_items"]
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            copy = _copy_stream(str(self.device))
            copy.wait_stream(main)
            events = []
            stage = _Staging.acquire(host.numel()).view(host.shape) if self._needs_staging else None
            ev = None
            try:
                with torch.cuda.stream(copy):
                    for c in reversed(range(n_chunks)):
                        lo, hi = bounds[c], bounds[c + 1]
                        if stage is not None:
                            stage[lo:hi].copy_(host[lo:hi])              # host threads; the previous DMA is in flight
                            self.S[lo:hi].copy_(stage[lo:hi], non_blocking=True)
                        else:
                            self.S[lo:hi].copy_(host[lo:hi], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy)
                        events.append((c, ev))
            finally:
                if stage is not None:
                    _Staging.release(ev)
            self.stats.zero_()
            self.cand_list[0].fill_(0)
            self.pair_list[0].zero_()
            if self._verify_progress is None:
                self._verify_progress = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._verify_progress.zero_()
            self._pairs_ready = False
            st = stream_ptr()
            for c, ev in events:
                lo, hi = bounds[c], bounds[c + 1]
                main.wait_event(ev)
                hi_pad = 0 if c == n_chunks - 1 else hi                  # 0: to the end of the padding rows
                check(L.tsc_pack_blocks(ptr(self.S), N, self.A, ptr(self.heavy_idx), self.M, ptr(self.packed), ptr(self.G),
                                        lo // 32, self.nb_pad if c == n_chunks - 1 else hi // 32, st), "tsc_pack_blocks")
                check(L.tsc_pack_screen(ptr(self.S), N, self.A, ptr(self.heavy_idx), self.M, ptr(self.PA), ptr(self.PB),
                                        ptr(self.PR), ptr(self.G), ptr(self.sG), ptr(self.CT), lo, hi_pad,
                                        self.tile_j, self._frame_ptr(), st), "tsc_pack_screen")
                it_dev, n_it = chunk_items[c]
                if n_it:
                    check(L.tsc_rmsd_screen(ptr(self.PA), ptr(self.PB), ptr(self.PR), ptr(self.G), ptr(self.sG),
                                            ptr(self.CT), N, self.M, ptr(it_dev), n_it, self.thr, ptr(self.sim_bits),
                                            ptr(self.cand_list), self.cand_stride, self.grid_ctas, self.screen_mode,
                                            self.pace, self._frame_ptr(), st), "tsc_rmsd_screen")
                # the candidates this sub-launch appended are verified while the next chunk is still on the bus, so
                # that after the last chunk only its own candidates are left (verify() makes the final call)
                if self.n_rb and c != 0:
                    check(L.tsc_rmsd_verify_incr(ptr(self.packed), N, self.M, ptr(self.row_blocks), self.n_rb, self.thr,
                                                 ptr(self.sim_bits), ptr(self.stats), ptr(self.pair_list),
                                                 self.pair_stride, ptr(self.cand_list), self.cand_stride,
                                                 ptr(self._verify_progress), 0, st), "tsc_rmsd_verify_incr")
            self._verify_incremental = True
        self.packed_ready = True

    def row_slice(self):
        """[lo, hi): the contiguous rows this rank uploads (and whose survivors it returns) when the host input is
        replicated on several ranks."""
        per = (self.N + self.world - 1) // self.world
        lo = min(self.rank * per, self.N)
        return lo, min(lo + per, self.N)

    def _upload_sharded(self, host):
        """Several ranks, the same host array on each: rank r copies rows [r per, (r + 1) per) to its GPU (1 / world of
        the bytes over its own PCIe link) and one NCCL all_gather_into_tensor over NVLink completes the ensemble on
        every GPU — world full uploads from one host made the end-to-end call SLOWER with every GPU added."""
        import torch.distributed as dist
        torch = self.torch
        lo, hi = self.row_slice()
        per = self._S_all.shape[0] // self.world
        with torch.cuda.device(self.device):
            mine = self._S_all[self.rank * per:(self.rank + 1) * per]
            if hi > lo:
                if self._needs_staging:
                    stage = _Staging.acquire((hi - lo) * host[0].numel()).view((hi - lo,) + tuple(host.shape[1:]))
                    ev = None
                    try:
                        stage.copy_(host[lo:hi])
                        mine[:hi - lo].copy_(stage, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record()
                    finally:
                        _Staging.release(ev)
                else:
                    mine[:hi - lo].copy_(host[lo:hi], non_blocking=True)
            if hi - lo < per:
                mine[hi - lo:].zero_()
            dist.all_gather_into_tensor(self._S_all.view(-1), mine.reshape(-1), group=self.group)

    def set_pairs(self, pairs):
        """Replace this rank's confirmed-pair list by explicit (i, j) rows, i < j (tests)."""
        torch = self.torch
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pairs.shape[0]
        self.pair_list[0, 0] = n
        m = min(n, self.pair_stride - 1)
        if m:
            self.pair_list[1:1 + m].copy_(torch.from_numpy(pairs[:m]).to(self.device))
        self._pairs_ready = True

    # ---- introspection ------------------------------------------------------------------------
    def stats_dict(self):
        s = self.stats.tolist()
        return {"candidates": s[0], "confirmed": s[1], "near_threshold": s[2], "degenerate": s[3]}

    def sim_rows_dense(self):
        """Owned similarity rows as a dense bool matrix (n_rb*32, N) — tests only."""
        torch = self.torch
        bits = self.sim_bits[: self.n_rb * _host.CB].cpu().numpy().view(np.uint32)
        rows = _host.global_rows_of(self.row_blocks_np)
        dense = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, : self.N].astype(bool)
        # words left of the diagonal block are never written (and never read): blank them
        col = np.arange(self.N)[None, :]
        dense &= col > rows[:, None]
        return rows, dense


def prune_conformers_rmsd(structures, atomnos, rmsd_thr=0.5, *, group=None, rank=None, world=None, pinned_output=True):
    """Drop-in for tscode.rmsd_pruning.prune_conformers_rmsd (rmsd_pruning.py:164-206).

    Removes similar structures (rmsd < rmsd_thr and max atomic deviation < 2*rmsd_thr on the
    non-hydrogen atoms, rotation-only Kabsch about the origin) with the reference's k-ladder,
    cache behaviour included.  Returns (structures[mask], mask) as numpy arrays.

    Several GPUs (one process per GPU, torch.distributed NCCL group, the SAME host array on every rank): pass
    `group` (or rank / world).  Every rank uploads 1 / world of the rows, the rest arrives over NVLink, the pair matrix
    is row-sharded, and every rank returns the FULL mask and the survivors of its own contiguous row slice
    (RmsdPruner.row_slice) — concatenated in rank order they are structures[mask].

    Large ensembles come back in page-locked host memory (the survivors are gathered on the GPU and copied out at
    PCIe speed instead of being re-indexed by the host); pinned_output=False returns ordinary pageable memory."""
    structures = np.asarray(structures)
    N = structures.shape[0]
    if N == 0:
        return structures[:0], np.zeros(0, dtype=np.bool_)
    if group is not None or world is not None:
        import torch.distributed as dist
        world = dist.get_world_size(group) if world is None else int(world)
        rank = dist.get_rank(group) if rank is None else int(rank)
    else:
        rank, world = 0, 1
    pr = RmsdPruner(structures, atomnos, rmsd_thr, rank=rank, world=world, group=group)
    pr.run_async()
    import torch
    big = (structures.flags.c_contiguous and structures.dtype == np.float64 and structures.nbytes > (1 << 22))
    lo, hi = pr.row_slice()
    out_buf = None
    if big and pinned_output:
        # while the GPU works: the output buffer (page-locked, from torch's caching host allocator: no cudaHostAlloc
        # after the first call of a size)
        out_buf = torch.empty((hi - lo,) + structures.shape[1:], dtype=torch.float64, pin_memory=True)
    elif big:
        out_buf = torch.empty((hi - lo,) + structures.shape[1:], dtype=torch.float64)
        out_buf.zero_()                   # first touch with all host threads (page faults cost more than the copy)
    mask_dev = pr.finish()
    if out_buf is not None and pinned_output:
        with torch.cuda.device(pr.device):
            idx = torch.nonzero(mask_dev[lo:hi]).squeeze(1)                  # (one sync: the survivor count)
            n = int(idx.numel())
            mask_host = torch.empty(N, dtype=torch.bool, pin_memory=True)
            mask_host.copy_(mask_dev, non_blocking=True)
            # The survivors exist twice — in the caller's host array and on the device — so they are fetched from
            # both at once: the first HOST_GATHER_SHARE of them by a multi-threaded host gather out of `structures`,
            # the rest gathered on the GPU and copied out at PCIe speed (94 MB over PCIe alone: 1.7 ms on C3).
            # (one rank only: with several ranks every rank's D2H is 1 / world already and runs on its own link, while
            # the host gathers of all ranks would share the same cores and memory — measured 6.35 against 4.9 ms at 2)
            n_host = int(n * HOST_GATHER_SHARE) if (n >= 4096 and world == 1) else 0
            idx_host = idx[:n_host].cpu() if n_host else None                # (small D2H; the stream is idle here)
            if n > n_host:
                dev_rows = torch.index_select(pr.S[lo:hi], 0, idx[n_host:])
                out_buf[n_host:n].copy_(dev_rows, non_blocking=True)
            if n_host:
                torch.index_select(torch.from_numpy(structures)[lo:hi], 0, idx_host, out=out_buf[:n_host])
            torch.cuda.current_stream().synchronize()
        return out_buf[:n].numpy(), mask_host.numpy().astype(np.bool_, copy=True)
    mask = mask_dev.cpu().numpy().astype(np.bool_)
    if out_buf is not None:
        idx = torch.from_numpy(np.flatnonzero(mask[lo:hi]))
        out = out_buf[:idx.numel()]
        torch.index_select(torch.from_numpy(structures)[lo:hi], 0, idx, out=out)
        return out.numpy(), mask
    return _take_rows(structures[lo:hi], mask[lo:hi]), mask


def _take_rows(structures, mask):
    """structures[mask] (rmsd_pruning.py:206) — same values, dtype and shape; done with torch's
    multi-threaded host index_select when the array allows it (5x faster than numpy's boolean
    indexing on a 100 MB ensemble)."""
    if isinstance(structures, np.ndarray) and structures.flags.c_contiguous and structures.dtype in (np.float64, np.float32) \
            and structures.nbytes > (1 << 22):
        import torch
        return torch.from_numpy(structures).index_select(0, torch.from_numpy(np.flatnonzero(mask))).numpy()
    return structures[mask]


def rmsd_and_max_batch(P, Q, broadcast_p=False):
    """rmsd_and_max_numba over explicit pairs: P, Q (n, M, 3) -> (rmsd (n,), maxdev (n,)) numpy."""
    torch = require_cuda()
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    Pt = torch.as_tensor(np.ascontiguousarray(P, dtype=np.float64)).to(dev)
    Qt = torch.as_tensor(np.ascontiguousarray(Q, dtype=np.float64)).to(dev)
    n, M = int(Qt.shape[0]), int(Qt.shape[1])
    r = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    d = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    if n and M:
        check(lib().tsc_rmsd_pairs(ptr(Pt), ptr(Qt), n, M, 1 if broadcast_p else 0, ptr(r), ptr(d), stream_ptr()),
              "tsc_rmsd_pairs")
    return r[:n].cpu().numpy(), d[:n].cpu().numpy()


def rmsd_and_max_numba(p, q):
    """Drop-in for tscode.rmsd_pruning.rmsd_and_max_numba (rmsd_pruning.py:6-41)."""
    r, d = rmsd_and_max_batch(np.asarray(p)[None], np.asarray(q)[None])
    return float(r[0]), float(d[0])


def _rmsd_similarity(ref, structures, rmsd_thr=0.5):
    """Drop-in for tscode.rmsd_pruning._rmsd_similarity (rmsd_pruning.py:208-224): is `ref`
    similar to any of `structures` (all atoms, no cache)."""
    structures = list(structures) if not isinstance(structures, np.ndarray) else structures
    if len(structures) == 0:
        return False
    r, d = rmsd_and_max_batch(np.asarray(ref)[None], np.asarray(structures), broadcast_p=True)
    return bool(np.any((r < rmsd_thr) & (d < 2 * rmsd_thr)))
