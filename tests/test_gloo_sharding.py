"""world_size-2 gloo test of the multi-GPU host logic on CPU: block-cyclic row ownership, the
per-round all-gather of verdicts/keys and their reassembly give the same mask as one rank.
The per-row scan itself is the oracle's ladder step here (the CUDA kernel is covered by -m gpu)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _round_rows(sim, mask, cache, N, k, cs, rows):
    """Reference ladder step restricted to `rows`."""
    out_mask = {}
    out_key = {}
    for i in rows:
        if i >= N:
            continue
        if not mask[i]:
            out_mask[i] = 0; out_key[i] = -1
            continue
        c = min(i // cs, k - 1) if cs > 0 else k - 1
        first = c * cs if cs > 0 else 0
        last = N if c == k - 1 else first + cs
        keep, key = 1, -1
        for j in range(i + 1, last):
            if mask[j]:
                if (first, first + j - i) in cache:
                    break
                if sim[i, j]:
                    keep, key = 0, first + j - i
                    break
        out_mask[i] = keep; out_key[i] = key
    return out_mask, out_key


def _worker(rank, world, port, N, seed, q):
    sys.path.insert(0, ROOT)
    from tscode_b200 import _host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    sim = np.triu(rng.random((N, N)) < 0.004, 1)
    rb = _host.owned_row_blocks(N, rank, world)
    my_rows = _host.global_rows_of(rb)
    n_rb_max = max(_host.owned_row_blocks(N, r, world).size for r in range(world))
    L = n_rb_max * _host.CB
    rows_pad = _host.num_blocks_padded(N) * _host.CB
    gidx = np.full((world, L), rows_pad, np.int64)
    for r in range(world):
        g = _host.global_rows_of(_host.owned_row_blocks(N, r, world))
        gidx[r, :g.size] = g
    gidx_t = torch.from_numpy(gidx.reshape(-1))
    mask = np.ones(N, bool)
    cache = set()

    def round_fn(k, cs):
        nonlocal mask
        om, ok = _round_rows(sim, mask, cache, N, k, cs, my_rows)
        loc_m = torch.zeros(L, dtype=torch.uint8); loc_k = torch.full((L,), -1, dtype=torch.int32)
        for t, i in enumerate(my_rows):
            if i in om:
                loc_m[t] = om[i]; loc_k[t] = ok[i]
        all_m = torch.empty(world * L, dtype=torch.uint8); all_k = torch.empty(world * L, dtype=torch.int32)
        dist.all_gather_into_tensor(all_m, loc_m); dist.all_gather_into_tensor(all_k, loc_k)
        full_m = torch.ones(rows_pad + 1, dtype=torch.uint8); full_k = torch.full((rows_pad + 1,), -1, dtype=torch.int32)
        full_m.index_copy_(0, gidx_t, all_m); full_k.index_copy_(0, gidx_t, all_k)
        new = full_m[:N].numpy().astype(bool); keys = full_k[:N].numpy()
        for i in np.flatnonzero(keys >= 0):
            c = min(i // cs, k - 1) if cs > 0 else k - 1
            cache.add(((c * cs) if cs > 0 else 0, int(keys[i])))
        mask = new
        return int(mask.sum())

    ran = _host.run_ladder(N, round_fn)
    if rank == 0:
        q.put((mask, ran))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,seed", [(700, 3), (1037, 4)])
def test_two_rank_ladder_matches_single(N, seed):
    from oracle import oracle_c
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + seed
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, seed, q)) for r in range(2)]
    for p in procs:
        p.start()
    mask, ran = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(seed)
    sim = np.triu(rng.random((N, N)) < 0.004, 1)
    ref, _, rounds = oracle_c.prune_heavy(np.zeros((N, 1, 3)), 0.5, sim_bytes=sim.astype(np.uint8))
    assert ran == [int(k) for k in rounds]
    assert np.array_equal(mask, ref)


def _worker_pairs(rank, world, port, N, seed, q):
    """The fused-ladder exchange: every rank emits the similar pairs of the rows it owns into a
    fixed-capacity block (header = count), ONE all-gather, then the whole ladder runs redundantly
    on the complete list (host model of elim_fused_kernel)."""
    sys.path.insert(0, ROOT)
    from tscode_b200 import _host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    sim = np.triu(rng.random((N, N)) < 0.004, 1)
    my_rows = _host.global_rows_of(_host.owned_row_blocks(N, rank, world))
    my_rows = my_rows[my_rows < N]
    mine = np.argwhere(sim[my_rows])
    mine[:, 0] = my_rows[mine[:, 0]]
    stride = 32 * N // world + 4096 + 1
    block = torch.zeros((stride, 2), dtype=torch.int32)
    block[0, 0] = mine.shape[0]
    block[1:1 + mine.shape[0]] = torch.from_numpy(mine.astype(np.int32))
    gathered = torch.empty((world * stride, 2), dtype=torch.int32)
    dist.all_gather_into_tensor(gathered, block)
    g = gathered.numpy().reshape(world, stride, 2)
    pairs = np.concatenate([g[r, 1:1 + g[r, 0, 0]] for r in range(world)])
    assert pairs.shape[0] == int(sim.sum())
    mask, ran = _host.ladder_pairlist_model(pairs, N)
    if rank == 0:
        q.put((mask, ran))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,seed", [(700, 3), (1037, 4)])
def test_two_rank_pair_list_exchange_matches_single(N, seed):
    from oracle import oracle_c
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + seed
    procs = [ctx.Process(target=_worker_pairs, args=(r, 2, port, N, seed, q)) for r in range(2)]
    for p in procs:
        p.start()
    mask, ran = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(seed)
    sim = np.triu(rng.random((N, N)) < 0.004, 1)
    ref, _, rounds = oracle_c.prune_heavy(np.zeros((N, 1, 3)), 0.5, sim_bytes=sim.astype(np.uint8))
    assert ran == [int(k) for k in rounds]
    assert np.array_equal(mask, ref)


def _worker_varlen(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from tscode_b200.embeds import gather_varlen, pose_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = 1003
    lo, hi = pose_range(P, rank, world)
    rng = np.random.default_rng(7)
    verdict = rng.random(P) < 0.13                      # the clash screen's verdicts (same on every rank here)
    poses = rng.normal(size=(P, 5, 3))
    mine = torch.from_numpy(poses[lo:hi][verdict[lo:hi]])
    got = gather_varlen(mine, world)
    if rank == 0:
        q.put((got.numpy(), poses[verdict]))
    dist.barrier()
    dist.destroy_process_group()


def test_variable_length_survivor_gather_keeps_pose_order():
    """Phase-1 -> phase-2 exchange of screen_and_prune: survivors of contiguous pose ranges, gathered with
    counts first, come back in global pose order on every rank (also with an empty rank)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_varlen, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got, want)
    from tscode_b200.embeds import pose_range
    assert [pose_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert pose_range(2, 3, 4) == (2, 2)
