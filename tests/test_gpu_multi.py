"""Two ranks on two GPUs of one box (skipped with fewer): the row-sharded prune with the confirmed-pair lists exchanged
by peer writes into symmetric memory (eliminate.cu: tsc_pairs_push / tsc_elim_fused_p2p) and, for comparison, through the
NCCL all-gather, gives the live reference's mask on every rank.  The 8-rank runs of tools/mgpu_check.py are in profiles/."""
import json
import os
import socket

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, rows, out):
    import torch.distributed as dist
    from tscode_b200 import rmsd_pruning as rp
    from tscode_b200.synth import gen_ensemble, mask_digest
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    res = []
    for peer in (True, False):
        rp._PeerLists.enabled = peer
        rp._PeerLists._cache.clear()
        for r in rows:
            S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"])
            for rep in range(2):                                     # twice: the peer-written arrays alternate
                pr = rp.RmsdPruner(S, np.full(r["M"], 6), r["thr"], rank=rank, world=world, group=dist.group.WORLD)
                m = pr.run().cpu().numpy()
                res.append((peer, pr._peer is not None, r["N"], mask_digest(m) == r["digest"], pr.ladder_used))
    dist.barrier()
    dist.destroy_process_group()
    out[rank] = res


def test_two_ranks_peer_written_lists_and_nccl_give_the_reference_masks():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    rows = [r for r in json.load(open(os.path.join(GOLDEN, "prune_masks.json")))["rows"]
            if not r.get("mixed_h") and r["N"] in (1000, 5000, 2000)][:4]
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, rows, out), nprocs=2, join=True)
        results = dict(out)
    assert set(results) == {0, 1}
    for rank, res in results.items():
        assert res and all(ok for _, _, _, ok, _ in res), (rank, res)
        assert all(ladder == "fused" for *_, ladder in res)
        assert not any(used for peer, used, *_ in res if not peer)
    # (whether symmetric memory is available is the box's business; when it is, both ranks must have used it)
    used = {rank: [u for peer, u, *_ in res if peer] for rank, res in results.items()}
    assert used[0] == used[1] and len(set(used[0])) == 1
