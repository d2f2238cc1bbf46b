#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-sharded prune, the drop-in with a process
group, the row-sharded rot_corr scan and the group-sharded embed pipeline must give the live-reference digests."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.rmsd_pruning import RmsdPruner  # noqa: E402
from tscode_b200.synth import gen_ensemble, mask_digest  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
gold = json.load(open("tests/golden/prune_masks.json"))["rows"] + json.load(open("tests/golden/prune_masks_big.json"))["rows"]
ok = True
for r in gold:
    if r.get("mixed_h") or r["N"] < 64:
        continue
    S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"])
    pr = RmsdPruner(S, np.full(r["M"], 6), r["thr"], rank=rank, world=world)
    mask = pr.run().cpu().numpy()
    good = mask_digest(mask) == r["digest"]
    ok &= good
    if rank == 0:
        print(f"world={world} N={r['N']} M={r['M']} survivors={int(mask.sum())} digest={mask_digest(mask)} "
              f"{'OK' if good else 'MISMATCH'} rounds={pr.rounds}", flush=True)
# ---- an overflowing pair list (tiny capacity): every rank sees the same headers and takes the bit-row ladder ----
r = [x for x in gold if x["N"] == 2000][0]
S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"])
for rep in range(3):                                    # (repeated: the peer-written arrays alternate between calls)
    pr = RmsdPruner(S, np.full(r["M"], 6), r["thr"], rank=rank, world=world, pair_cap=64)
    mask = pr.run().cpu().numpy()
    good = mask_digest(mask) == r["digest"] and pr.ladder_used == "bitrows"
    ok &= good
    pr2 = RmsdPruner(S, np.full(r["M"], 6), r["thr"], rank=rank, world=world)
    good2 = mask_digest(pr2.run().cpu().numpy()) == r["digest"] and pr2.ladder_used == "fused"
    ok &= good2
    if rank == 0:
        print(f"world={world} overflowing pair list -> {pr.ladder_used}: {'OK' if good else 'MISMATCH'}; then fused again: "
              f"{'OK' if good2 else 'MISMATCH'} (peer lists: {pr2._peer is not None})", flush=True)
# ---- the public drop-in with a process group: sharded upload, full mask, survivors of the rank's row slice ----
from tscode_b200.rmsd_pruning import prune_conformers_rmsd  # noqa: E402
r = [x for x in gold if x["N"] == 10000][0]
S = gen_ensemble(r["seed"], r["N"], r["M"], r["n_clusters"], sigma_noise=r["sigma_noise"])
out, mask = prune_conformers_rmsd(S, np.full(r["M"], 6), r["thr"], group=dist.group.WORLD)
per = (r["N"] + world - 1) // world
lo, hi = rank * per, min((rank + 1) * per, r["N"])
good = mask_digest(mask) == r["digest"] and np.array_equal(out, S[lo:hi][mask[lo:hi]])
ok &= good
if rank == 0:
    print(f"world={world} drop-in with group: digest {'OK' if good else 'MISMATCH'}; rank 0 returned {out.shape[0]} survivors", flush=True)

# ---- rot_corr, rows of the forward scan dealt round-robin ----
sys.path.insert(0, "oracle")
import rotor_molecules as rm  # noqa: E402
from tscode_b200.torsion_module import TorsionInfo, prune_conformers_rmsd_rot_corr  # noqa: E402
gb = json.load(open("tests/golden/rotcorr_big.json"))["fixtures"]
for name, f in gb.items():
    g = np.load(f"tests/golden/rotcorr_{name}.npz")
    info = TorsionInfo([tuple(t) for t in f["torsions"]], [tuple(a) for a in f["angles"]], g["rot_masks"].astype(bool),
                       g["node_lists"].astype(bool))
    S4, at4 = rm.ensemble_tritbu63(f["seed"], f["N"])
    import time
    prune_conformers_rmsd_rot_corr(S4, at4, None, f["thr"], torsion_info=info, max_structures=None, mode="stateless",
                                   group=dist.group.WORLD)
    dist.barrier(); t0 = time.perf_counter()
    o4, m4 = prune_conformers_rmsd_rot_corr(S4, at4, None, f["thr"], torsion_info=info, max_structures=None, mode="stateless",
                                            group=dist.group.WORLD)
    dt = time.perf_counter() - t0
    good = mask_digest(m4) == f["digest"]
    ok &= good
    if rank == 0:
        print(f"world={world} rot_corr {name}: survivors={int(m4.sum())} {'OK' if good else 'MISMATCH'} {dt * 1e3:.0f} ms", flush=True)

# ---- configs[4] pipeline, groups dealt to the ranks ----
from tscode_b200.embeds import cyclical_embed_pipeline  # noqa: E402
from tscode_b200.synth import gen_cyclical_groups  # noqa: E402
pipe = json.load(open("tests/golden/embed_pipeline.json"))["rows"]
for name in ("small", "c5"):
    rr = pipe[name]
    d = gen_cyclical_groups(rr["seed"], rr["n_groups"])
    res = cyclical_embed_pipeline(d, np.full(150, 6), rank=rank, world=world, group=dist.group.WORLD)
    v, k, m = res["verdict"].cpu().numpy(), res["kept"].cpu().numpy(), res["mask"].cpu().numpy()
    good = mask_digest(v) == rr["clash_digest"] and mask_digest(k) == rr["kept_digest"] and mask_digest(m) == rr["prune_digest"]
    ok &= good
    if rank == 0:
        print(f"world={world} embed pipeline {name}: {int(v.sum())} / {int(k.sum())} / {int(m.sum())} {'OK' if good else 'MISMATCH'} {res['ms']}", flush=True)

t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t[0]) == 1 else 1)
