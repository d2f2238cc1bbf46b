#!/usr/bin/env python
"""bench.py — headline benchmark of the conformer-ensemble hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--variant screen|dmma|fma]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the north-star target): prune_conformers_rmsd on a synthetic clustered ensemble of
50 000 conformers x 80 heavy atoms, rmsd_thr 0.5 (gen_ensemble(3, 50000, 80, 5000)).  One step = one
whole prune: pack, all-pairs similarity (N(N-1)/2 = 1.25e9 pairs), exact verification, k-ladder -> survivor mask.  At
N > 1 the rows of the pair matrix are dealt to the ranks (strong scaling: the pair count is fixed), the confirmed pairs
are all-gathered once and the ladder runs on every rank.

Prints ONE JSON line (rank 0).
  value         pairs/s with the ensemble already resident in HBM (CUDA events, L2 flushed between steps, max over ranks)
  e2e           the same through the public drop-in prune_conformers_rmsd(numpy, atomnos, thr[, group]) — the SAME call at
                every N: host array in, H2D inside the timed region (1/N of the rows per rank + NVLink all-gather),
                mask and survivors out (GPU gather + D2H)
  roofline      the dominant kernel (all-pairs screen) against the measured dense 16-bit tensor peak
  cpu_baseline  the CPU side on this box: per-pair rate and time-to-the-same-mask of the lazy reference algorithm
                (oracle C port here; `--impl reference` runs the unmodified numba reference from baseline/_ref)
  clash         the secondary metric at every N (configs[1]: 100k two-fragment poses, pose-sharded), poses/s
  configs       one-liners for configs[3] (rot_corr 20 000 x 63) and configs[4] (1M-pose cyclical embed pipeline)
  variants      N = 1: the north star's FP64 variants of the screen (DMMA / FMA) against the self-measured FP64 peak
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C3 = dict(seed=3, N=50_000, M=80, n_clusters=5000, thr=0.5, digest="478bc29df1e239da", survivors=48867)
C2 = dict(seed=0, P=100_000, n_atoms=(50, 50), thresh=1.5, max_clashes=0, digest="6e7eb19c842b4798", passes=12695)
METRIC = "rmsd_pairs_per_s"
UNIT = "pairs/s"
DTYPE = "f64 (fp16/fp32 conservative pre-screen on tcgen05 + exact f64 verify; masks bit-exact)"
REF_ARM_FILE = os.path.join(ROOT, "baseline", "_ref", ".last_reference_arm.json")


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.times, self.proc, self.idx = [], [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())
            self.times.append(time.perf_counter())

    def stop(self, window=None):
        """window = (t0, t1) in time.perf_counter(): only samples received inside it count (the sampler is started
        before the warm-up steps so that it is already running when the short timed region begins); if none fell
        inside, the samples of the 150 ms before it — the warm-up steps of the same workload — are used and the result
        says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows, where = list(self.rows), "whole run of the sampler"
        if window is not None:
            t0, t1 = window
            pairs = list(zip(self.times, self.rows))
            rows, where = [r for t, r in pairs if t0 <= t <= t1 + 0.02], "timed region"
            if not rows:
                rows, where = [r for t, r in pairs if t0 - 0.15 <= t <= t1 + 0.02], "timed region + the warm-up steps right before it"
        sm, smax, power, reasons = [], [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "sampled": where,
                "reasons": sorted(reasons)}


def fp64_peaks(torch):
    """Self-measured FP64 ceilings (register-resident DFMA / DMMA loops, tools/probes): MEASURED_PEAKS.json has no FP64
    entry.  The probe library is a measurement aid, not part of the product: without it the values of round 1 are quoted."""
    import ctypes as C
    path = os.path.join(ROOT, "tools", "probes", "libtsc_probe.so")
    if not os.path.exists(path):
        return {"dfma": 34.2, "dmma": 37.1, "mixed": 36.7, "source": "round-1 measurement (tools/probes not built)"}
    from tscode_b200._lib import check, ptr, stream_ptr
    P = C.CDLL(path)
    P.tsc_bench_fp64.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 4
    scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
    out = {}
    for kind, name in ((0, "dfma"), (1, "dmma"), (2, "mixed")):
        best = 0.0
        for _ in range(3):
            fl = (C.c_double * 2)()
            ms = C.c_float()
            check(P.tsc_bench_fp64(kind, 4000, 2, 512, ptr(scratch), C.cast(fl, C.c_void_p),
                                   C.cast(C.byref(ms), C.c_void_p), stream_ptr()), "tsc_bench_fp64")
            best = max(best, (fl[0] + fl[1]) / (ms.value * 1e-3) / 1e12)
        out[name] = round(best, 2)
    out["source"] = "self-measured in this run (tools/probes/peaks.cu)"
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def config_dict(cfg, world, extra=None):
    d = {"workload": f"BASELINE configs[2]: prune_conformers_rmsd all-pairs, {cfg['N']} conformers x {cfg['M']} "
                     f"heavy atoms, rmsd_thr {cfg['thr']}, gen_ensemble(seed={cfg['seed']}, n_clusters={cfg['n_clusters']})",
         "pairs_per_step": cfg["N"] * (cfg["N"] - 1) // 2,
         "parallelism": f"128-row panels dealt to {world} rank(s) in snake order",
         "l2": "working set per step (96 MB structures + 96 MB packed + 313 MB similarity bits) exceeds the 126 MB L2; "
               "a 512 MB scratch write additionally flushes L2 between timed steps"}
    if extra:
        d.update(extra)
    return d


# ---------------------------------------------------------------------------------------------
# CPU side
# ---------------------------------------------------------------------------------------------
def sample_pairs(rng, N, n):
    import numpy as np
    i = rng.integers(0, N - 1, size=n); j = rng.integers(0, N, size=n)
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    hi = np.minimum(np.where(lo == hi, hi + 1, hi), N - 1)
    return lo.astype(np.int64), hi.astype(np.int64)


def cpu_baseline_port(S, thr, full_time_to_mask=True):
    """Oracle C port (oracle/oracle.c, OpenMP over all host threads): (a) rate of fully evaluated pairs
    (rmsd_and_max + thresholds, rmsd_pruning.py:6-41,:75) on a bounded random sample; (b) the reference's LAZY
    algorithm to the same mask on the FULL ensemble (it evaluates ~1.5 % of the pairs)."""
    import numpy as np
    from oracle import oracle_c
    oracle_c.set_threads(host_threads())
    N = S.shape[0]
    rng = np.random.default_rng(1)
    ii, jj = sample_pairs(rng, N, 100_000)
    oracle_c.eval_pairs(S, thr, ii[:1000], jj[:1000])
    t0 = time.perf_counter(); oracle_c.eval_pairs(S, thr, ii, jj); dt = time.perf_counter() - t0
    n = int(min(max(100_000 * 4.0 / max(dt, 1e-6), 100_000), 2e8))
    ii, jj = sample_pairs(rng, N, n)
    t0 = time.perf_counter(); hits = oracle_c.eval_pairs(S, thr, ii, jj); dt = time.perf_counter() - t0
    out = {"value": n / dt, "unit": UNIT, "cores": oracle_c.num_threads(), "kind": "port",
           "sample": f"{n} random (i<j) pairs of the same {N}x{S.shape[1]} ensemble, every pair fully evaluated "
                     f"(oracle/oracle.c, OpenMP), {dt:.1f} s; {hits} similar"}
    if full_time_to_mask:
        t0 = time.perf_counter(); m, ne, _ = oracle_c.prune_heavy(S, thr); tl = time.perf_counter() - t0
        from tscode_b200.synth import mask_digest
        out["time_to_mask"] = {"seconds": round(tl, 2), "N": N, "pairs_evaluated": int(ne), "pairs_total": N * (N - 1) // 2,
                               "survivors": int(m.sum()), "digest": mask_digest(m), "kind": "port",
                               "note": "the reference's lazy k-ladder (cache hits end most rows early), C port, all host "
                                       "threads; its last rounds are serial as in the reference"}
    return out


def load_numba_reference():
    """The UNMODIFIED reference (pip-installed into baseline/_ref) if numba is importable."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "tscode")):
        return None, "baseline/_ref/tscode not present"
    try:
        if ref not in sys.path:
            sys.path.insert(0, ref)
        import numba
        from tscode.rmsd_pruning import prune_conformers_rmsd, rmsd_and_max_numba
        return (numba, prune_conformers_rmsd, rmsd_and_max_numba), None
    except Exception as e:                                     # numba missing / import error: the C port stands in
        return None, repr(e)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on this box's host cores, rank 0 only.
    Preferred: the unmodified numba code from baseline/_ref (kind "reference"); else the oracle C port (kind "port").
    Step = the reference's prune_conformers_rmsd on a bounded sample (the same generator at a smaller size,
    calibrated to a few seconds); value = pairs the call decides per second = n(n-1)/2 / time (the reference is lazy:
    it reaches the mask by evaluating a few per cent of the pairs — that IS its algorithm, so this is its whole-job
    rate; the rate of fully evaluated pairs is reported beside it).  Once, outside the timed steps: the FULL 50 000 x 80
    prune (time to the same mask), unless --no-full-reference."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    from tscode_b200.synth import gen_ensemble, mask_digest
    n_thr = host_threads()
    cfg = dict(C3)
    if args.n_conformers:
        cfg.update(N=args.n_conformers, n_clusters=max(args.n_conformers // 10, 1), digest=None)
    S = gen_ensemble(cfg["seed"], cfg["N"], cfg["M"], cfg["n_clusters"])
    atomnos = np.full(cfg["M"], 6)
    thr = cfg["thr"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ref, why = load_numba_reference()
    per_pair = None
    if ref is not None:
        numba, prune, rmsd_and_max = ref
        kind, cores = "reference", int(numba.get_num_threads())
        prune(gen_ensemble(0, 1000, 40, 100), np.full(40, 6), 0.5)               # JIT (both specialisations run below)

        def run(X):
            t0 = time.perf_counter(); _, m = prune(X, atomnos, thr); return time.perf_counter() - t0, m

        @numba.njit(parallel=True)
        def eval_pairs(H, ii, jj, t):
            hits = 0
            for k in numba.prange(ii.shape[0]):
                r, d = rmsd_and_max(H[ii[k]], H[jj[k]])
                if r < t and d < 2 * t:
                    hits += 1
            return hits
        rng = np.random.default_rng(1)
        ii, jj = sample_pairs(rng, cfg["N"], 20000)
        eval_pairs(S, ii, jj, thr)
        ii, jj = sample_pairs(rng, cfg["N"], 400_000 * max(1, cores // 4))
        t0 = time.perf_counter(); eval_pairs(S, ii, jj, thr); dtp = time.perf_counter() - t0
        per_pair = {"value": ii.shape[0] / dtp, "unit": UNIT, "what": "rmsd_and_max_numba (unmodified) on explicit random pairs "
                    "inside a numba prange loop", "pairs": int(ii.shape[0]), "seconds": round(dtp, 2)}
    else:
        from oracle import oracle_c
        oracle_c.set_threads(n_thr)
        kind, cores = "port", oracle_c.num_threads()

        def run(X):
            t0 = time.perf_counter(); m, _, _ = oracle_c.prune_heavy(X, thr); return time.perf_counter() - t0, m
    # The per-step sample is a SCALED-DOWN INSTANCE of the same generator (same conformers-per-cluster ratio), not a
    # prefix of the big ensemble: the reference is lazy only where similar pairs exist, and a prefix of 50 000
    # conformers in 5 000 clusters has almost none.  Size calibrated to ~4 s per step (time grows like n^1.5).
    sample = lambda n: gen_ensemble(cfg["seed"], n, cfg["M"], max(n // 10, 1))
    t_cal, _ = run(sample(min(5000, cfg["N"])))
    n_s = int(min(cfg["N"], max(2000, 5000 * (4.0 / max(t_cal, 1e-3)) ** (1 / 1.5))) // 1000 * 1000) or cfg["N"]
    n_s = min(max(n_s, 2000), cfg["N"], 20000)
    S_s = sample(n_s)
    for _ in range(max(0, min(args.warmup, 2))):
        run(S_s)
    ts = [run(S_s)[0] for _ in range(args.steps)]
    pairs_s = n_s * (n_s - 1) // 2
    value = pairs_s * len(ts) / sum(ts)
    full = None
    if world == 1 and not args.no_full_reference:
        tl, m = run(S)
        full = {"seconds": round(tl, 2), "N": cfg["N"], "pairs_total": cfg["N"] * (cfg["N"] - 1) // 2,
                "pairs_per_s": cfg["N"] * (cfg["N"] - 1) / 2 / tl, "survivors": int(m.sum()), "digest": mask_digest(m),
                "matches_expected_digest": (mask_digest(m) == cfg["digest"]) if cfg.get("digest") else None}
    sample = (f"prune_conformers_rmsd on gen_ensemble(seed={cfg['seed']}, N={n_s}, M={cfg['M']}, n_clusters={n_s // 10}) per step "
              f"— the same generator at a size the CPU finishes in seconds — ({pairs_s} pairs decided, lazily), "
              f"{'unmodified numba reference from baseline/_ref' if kind == 'reference' else 'oracle C port (' + str(why) + ')'}, "
              f"{cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ts) / len(ts) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, world, extra={"reference_sample": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "per_pair_rate": per_pair, "full_size_time_to_mask": full,
            "note": "value = pairs DECIDED per second by the reference's lazy algorithm on the sample (its effective rate "
                    "grows with the ensemble: ~2x higher at the full size); full_size_time_to_mask is the same call on "
                    "the whole 50 000 x 80 ensemble, once, outside the timed steps"}
    try:
        os.makedirs(os.path.dirname(REF_ARM_FILE), exist_ok=True)
        json.dump({"kind": kind, "cores": cores, "value": value, "full": full, "when": time.time()}, open(REF_ARM_FILE, "w"))
    except Exception:
        pass
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from tscode_b200.numba_functions import PoseBatch
    from tscode_b200.rmsd_pruning import RmsdPruner, prune_conformers_rmsd
    from tscode_b200.synth import gen_ensemble, gen_poses, mask_digest

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    cfg = dict(C3)
    if args.n_conformers:
        cfg.update(N=args.n_conformers, n_clusters=max(args.n_conformers // 10, 1), digest=None, survivors=None)
    N, M, thr = cfg["N"], cfg["M"], cfg["thr"]
    S_host = gen_ensemble(cfg["seed"], N, M, cfg["n_clusters"])
    atomnos = np.full(M, 6)
    pinned = torch.empty(S_host.shape, dtype=torch.float64).pin_memory()
    pinned.copy_(torch.from_numpy(S_host))
    S_pinned_np = pinned.numpy()
    S_dev = pinned.to(dev)
    pr = RmsdPruner(S_dev, atomnos, thr, variant=args.variant, rank=rank, world=world, group=group, device=dev,
                    ladder=args.ladder)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor(x, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    phase_ms = {"pack": [], "screen": [], "verify": [], "eliminate": []}

    def step(record=False):
        e = [ev() for _ in range(5)]
        e[0].record(); pr.pack()
        e[1].record(); pr.screen()
        e[2].record(); pr.verify()
        e[3].record(); mask = pr.eliminate()
        e[4].record()
        if record:
            torch.cuda.synchronize()
            for k, name in enumerate(phase_ms):
                phase_ms[name].append(e[k].elapsed_time(e[k + 1]))
        return mask

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        mask = step()
    barrier()
    # nvidia-smi needs ~0.1 s to print its first line: keep the GPU on the same workload (untimed) until the sampler is
    # alive (several ranks: the step holds collectives, so every rank runs the same fixed number of extra steps)
    if world == 1:
        t_up = time.perf_counter() + 0.5
        while not sampler.rows and time.perf_counter() < t_up:
            step()
            torch.cuda.synchronize()
    else:
        for _ in range(60):
            step()
        torch.cuda.synchronize()
    barrier()
    t_events = []
    w0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)                      # L2 flush between timed iterations
        e0, e1 = ev(), ev()
        e0.record()
        mask = step(record=True)
        e1.record()
        torch.cuda.synchronize()
        t_events.append(e0.elapsed_time(e1))
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    clocks = sampler.stop(window=(w0, time.perf_counter())) if rank == 0 else None
    verify_stats = pr.stats_dict()          # counters of the last timed step
    if rank == 0 and not args.skip_extras:
        # The timed region lasts a few tens of ms: nvidia-smi (100 ms period) sees it once at best.  Untimed addition:
        # the same step back to back for ~0.6 s under a second sampler (no collective inside: world == 1 only).
        if world == 1:
            s2 = ClockSampler(local)
            s2.start()
            t_end = time.perf_counter() + 0.6
            while time.perf_counter() < t_end:
                for _ in range(20):
                    step()
                torch.cuda.synchronize()
            clocks["sustained"] = s2.stop()
    step_ms, screen_ms = reduce_max([sum(t_events) / len(t_events), statistics.mean(phase_ms["screen"])])
    phase_max = dict(zip(phase_ms, reduce_max([statistics.mean(v) for v in phase_ms.values()])))
    pairs = N * (N - 1) // 2
    value = pairs / (step_ms * 1e-3)
    mask_np = mask.cpu().numpy()
    digest = mask_digest(mask_np)
    digests_equal = True
    if world > 1:
        allm = [None] * world
        dist.all_gather_object(allm, digest)
        digests_equal = len(set(allm)) == 1

    # ---- e2e: the public drop-in, host buffers, the same call at every N -------------------------------------
    e2e = None
    if not args.skip_extras:
        kw = dict(group=group) if world > 1 else {}
        for _ in range(2):
            out, m2 = prune_conformers_rmsd(S_pinned_np, atomnos, thr, **kw)
        ts = []
        for _ in range(max(2, min(args.steps, 5))):
            flush.fill_(1)
            barrier()
            t0 = time.perf_counter()
            out, m2 = prune_conformers_rmsd(S_pinned_np, atomnos, thr, **kw)
            ts.append(time.perf_counter() - t0)
        t_e2e = reduce_max([statistics.mean(ts)])[0]
        n_out = reduce_max([float(out.shape[0])])[0] if world == 1 else None
        if world > 1:
            cnt = torch.tensor([out.shape[0]], dtype=torch.int64, device=dev)
            dist.all_reduce(cnt)
            n_out = float(cnt.item())
        ok = bool(np.array_equal(m2, mask_np)) and int(n_out) == int(mask_np.sum())
        lo, hi = pr.row_slice()
        ok = ok and bool(np.array_equal(out, S_host[lo:hi][mask_np[lo:hi]]))
        from tscode_b200 import rmsd_pruning as _rp
        host_share = float(_rp.HOST_GATHER_SHARE) if world == 1 else 0.0
        e2e = {"value": pairs / t_e2e, "unit": UNIT, "ms_per_call": t_e2e * 1e3,
               "h2d_bytes_per_step": int(S_host.nbytes),                      # summed over ranks: each uploads 1 / world
               # survivors that cross the bus (the API takes HOST_GATHER_SHARE of them from the caller's own host array
               # while the rest is in flight) + the mask
               "d2h_bytes_per_step": int(round(int(mask_np.sum()) * (1.0 - host_share))) * M * 24 + N * world,
               "host_gather_share": host_share,
               "result_equals_device_path": ok,
               "api": "tscode_b200.rmsd_pruning.prune_conformers_rmsd(structures: numpy (pinned), atomnos, rmsd_thr"
                      + (", group=WORLD) on every rank -> (survivors of the rank's row slice, full mask)" if world > 1
                         else ") -> (structures[mask], mask)")
                      + "; H2D of the structures, GPU gather + D2H of the survivors (part of them copied from the caller's "
                        "host array by host threads meanwhile) and of the mask inside the timed region"}
        pcie = {"h2d_gbps_measured": 54.0, "d2h_gbps_measured": 57.0, "source": "tools/pcie_probe.py on this pool (profiles/)"}
        floor_ms = (S_host.nbytes / world / 54e9 + int(mask_np.sum()) * (1.0 - host_share) * M * 24 / world / 57e9) * 1e3
        e2e["pcie_roofline"] = {**pcie, "copy_floor_ms": floor_ms, "frac": floor_ms / (t_e2e * 1e3)}

    # ---- secondary metric at every N: clash-checked poses/s (configs[1]), pose-sharded -------------------------
    clash = None
    if not args.skip_extras:
        from tscode_b200.embeds import pose_range
        frags, conf, R, tt = gen_poses(C2["seed"], C2["P"], C2["n_atoms"])
        plo, phi = pose_range(C2["P"], rank, world)
        pb = PoseBatch(frags, conf[plo:phi], R[plo:phi], tt[plo:phi])
        for _ in range(3):
            v = pb.clash(C2["thresh"], C2["max_clashes"])
        barrier()
        ks = []
        for _ in range(10):
            flush.fill_(1)
            e0, e1 = ev(), ev()
            e0.record(); v = pb.clash(C2["thresh"], C2["max_clashes"]); e1.record()
            torch.cuda.synchronize()
            ks.append(e0.elapsed_time(e1))
        kms = reduce_max([statistics.median(ks)])[0]
        vloc = v.cpu().numpy()
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, vloc)
            vn = np.concatenate(parts)
        else:
            vn = vloc
        Rp = torch.from_numpy(R[plo:phi]).pin_memory(); tp = torch.from_numpy(tt[plo:phi]).pin_memory()
        cp = torch.from_numpy(conf[plo:phi].astype(np.int32)).pin_memory()
        es = []
        for _ in range(5):
            barrier(); t0 = time.perf_counter()
            pb2 = PoseBatch(frags, cp, Rp, tp)
            v2 = pb2.clash(C2["thresh"], C2["max_clashes"]).cpu()
            es.append(time.perf_counter() - t0)
        te = reduce_max([statistics.median(es)])[0]
        flop = 8 * 50 * 50 + 18 * 100
        clash = {"metric": "clash_checked_poses_per_s", "value": C2["P"] / (kms * 1e-3), "unit": "poses/s", "kernel_ms": kms,
                 "workload": "BASELINE configs[1]: 100k two-fragment poses (2 x 50 atoms), fused rotation + clash test, "
                             f"thresh 1.5, max_clashes 0; contiguous pose ranges on {world} rank(s), no collective",
                 "parity": {"passes": int(vn.sum()), "digest": mask_digest(vn), "matches_reference": mask_digest(vn) == C2["digest"]},
                 "algorithmic_flop_per_pose": flop, "achieved_tflops": flop * C2["P"] / (kms * 1e-3) / 1e12,
                 "note": "87 % of these poses clash and leave at the first 8-atom check, so algorithmic flops / time "
                         "overstates the FP64-pipe use; an all-pass pose set is profiled in profiles/",
                 "e2e": {"value": C2["P"] / te, "unit": "poses/s", "h2d_bytes_per_step": int(R.nbytes + tt.nbytes + conf.size * 4),
                         "d2h_bytes_per_step": C2["P"]}}

    # ---- weak scaling beside the strong-scaling headline (N > 1): pairs proportional to the number of GPUs ---------
    weak = None
    if world > 1 and not args.skip_extras and not args.n_conformers:
        Nw = int(50000 * math.sqrt(world) / 128) * 128
        Sw = torch.from_numpy(gen_ensemble(3, Nw, 80, Nw // 10)).to(dev)
        pw = RmsdPruner(Sw, atomnos, thr, variant=args.variant, rank=rank, world=world, group=group, device=dev)
        for _ in range(2):
            mw = pw.run()
        barrier()
        tsw = []
        for _ in range(3):
            flush.fill_(1)
            e0, e1 = ev(), ev()
            e0.record(); mw = pw.run(); e1.record(); torch.cuda.synchronize()
            tsw.append(e0.elapsed_time(e1))
        tw = reduce_max([statistics.mean(tsw)])[0]
        gold = {}
        try:
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "prune_masks_weak.json"))).get(str(world), {})
        except Exception:
            pass
        dw = mask_digest(mw.cpu().numpy())
        weak = {"n_conformers": Nw, "pairs_per_step": Nw * (Nw - 1) // 2, "ms_per_step": tw,
                "value": Nw * (Nw - 1) / 2 / (tw * 1e-3), "unit": UNIT, "survivors": int(mw.sum().item()), "digest": dw,
                "matches_c_oracle_digest": (dw == gold["digest"]) if gold.get("digest") else None,
                "note": "pairs per GPU as at N = 1 (N_conformers = 50000 sqrt(world)); compare value with world x the N = 1 value"}
        del pw, Sw

    # ---- configs[4]: 1M-pose trimolecular cyclical embed -> clash -> de-dup -> prune, at every N ----------------------
    c5 = None
    if not args.skip_extras and not args.n_conformers:
        try:
            from tscode_b200.embeds import cyclical_embed_pipeline
            from tscode_b200.synth import gen_cyclical_groups
            g5 = json.load(open(os.path.join(ROOT, "tests", "golden", "embed_pipeline.json")))["rows"]["c5"]
            d5 = gen_cyclical_groups(g5["seed"], g5["n_groups"])
            a5 = np.full(int(sum(f.shape[1] for f in d5["frags"])), 6)
            kw5 = dict(rank=rank, world=world, group=group)
            for _ in range(2):                       # the first calls pay cudaMalloc of the 720 MB bit matrix and friends
                cyclical_embed_pipeline(d5, a5, **kw5)
            t5 = None
            for _ in range(2):
                barrier(); t0 = time.perf_counter()
                r5 = cyclical_embed_pipeline(d5, a5, **kw5)
                dt5 = reduce_max([time.perf_counter() - t0])[0]
                t5 = dt5 if t5 is None else min(t5, dt5)
            v5, k5, m5 = r5["verdict"].cpu().numpy(), r5["kept"].cpu().numpy(), r5["mask"].cpu().numpy()
            c5 = {"workload": f"BASELINE configs[4]: trimolecular cyclical embed, {r5['n_poses']} poses (3 x 50 atoms): pose "
                              "parameters on the device -> fused transform + clash -> group-local de-dup -> RMSD prune; groups "
                              f"dealt to {world} rank(s)",
                  "ms_total": t5 * 1e3, "poses_per_s": r5["n_poses"] / t5, "phase_ms": r5["ms"],
                  "clash_pass": int(v5.sum()), "kept_after_dedup": int(k5.sum()), "survivors": int(m5.sum()),
                  "matches_live_reference": bool(mask_digest(v5) == g5["clash_digest"] and mask_digest(k5) == g5["kept_digest"]
                                                 and mask_digest(m5) == g5["prune_digest"]),
                  "reference_wall_s": {"generate_clash_dedup": g5["wall_s_generate_clash_dedup"], "prune": g5["wall_s_prune"],
                                       "where": "build container, 8 threads (oracle/gen_golden_c5.py)"}}
            del r5
        except Exception as exc:
            c5 = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    peaks = fp64_peaks(torch) if not args.skip_extras else {"dfma": 34.2, "dmma": 37.1, "mixed": 36.7, "source": "round 1"}
    mp, mp_src = measured_peaks()
    flops = 18.0 * M * pairs / world
    achieved = flops / (screen_ms * 1e-3) / 1e12
    if args.variant == "screen":
        peak = float(mp["bf16_tflops"])
        mode = pr.screen_mode
        kname = f"rmsd_screen_kernel<{mode},{pr.tile_j}>"
        psrc = (f"dense 16-bit tensor peak = the {mp_src} cuBLAS bf16 burst figure of MEASURED_PEAKS.json ({mp['bf16_tflops']} "
                "TFLOP/s); the kernel issues kind::f16 MMAs (FP16 operands, FP32 accumulation in TMEM)")
    else:
        peak = max(peaks["dmma"], peaks["dfma"])
        kname = f"rmsd_sim_kernel<{'ConsumerDMMA' if args.variant == 'dmma' else 'ConsumerFMA'}>"
        psrc = (f"FP64 ceiling {peaks['source']}: DMMA.8x8x4 {peaks['dmma']} / DFMA {peaks['dfma']} / both {peaks['mixed']} "
                f"TFLOP/s; MEASURED_PEAKS.json ({mp_src}) has no FP64 entry")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and world == 1 and not args.n_conformers:
        traffic = json.load(open(tpath)).get(kname)
    roofline = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": psrc, "algorithmic_flop_per_pair": 18 * M,
                "kernel_ms": screen_ms, "kernel_share_of_step": screen_ms / step_ms,
                "note": "18 M flop per pair is the contraction only (SURVEY 8d); the screen's tensor pipe is far from "
                        "saturated — its epilogue (FP32 exclusion tests on CUDA cores) sets the pace, DESIGN.md 4.1"}

    # ---- CPU baseline (rank 0, N = 1 only, bounded) -----------------------------------------------------------
    cpu = None
    tmask = None
    if world == 1 and not args.no_cpu and not args.skip_extras:
        cpu = cpu_baseline_port(S_host, thr, full_time_to_mask=not args.n_conformers)
        tm = cpu.get("time_to_mask")
        if tm and e2e:
            tmask = {"vs": "oracle C port, lazy algorithm, full 50 000 x 80 ensemble, same mask", "cpu_seconds": tm["seconds"],
                     "b200_e2e_seconds": e2e["ms_per_call"] * 1e-3, "speedup": tm["seconds"] / (e2e["ms_per_call"] * 1e-3),
                     "cpu_mask_equals_gpu_mask": tm["digest"] == digest}
            try:
                ra = json.load(open(REF_ARM_FILE))
                if ra.get("full") and time.time() - ra.get("when", 0) < 6 * 3600:
                    tmask["vs_numba_reference"] = {"cpu_seconds": ra["full"]["seconds"], "cores": ra["cores"], "kind": ra["kind"],
                                                   "speedup": ra["full"]["seconds"] / (e2e["ms_per_call"] * 1e-3),
                                                   "source": "this box, `bench.py --impl reference` run just before"}
            except Exception:
                pass

    # ---- N = 1: the north star's FP64 variants, configs[3] ---------------------------------------------------------
    variants = None
    c4 = None
    shaped = None
    if world == 1 and not args.skip_extras and not args.n_conformers:
        variants = {}
        for vname in ("dmma", "fma"):
            pv = RmsdPruner(S_dev, atomnos, thr, variant=vname, device=dev)
            pv.pack(); pv.screen(); torch.cuda.synchronize()
            tv = []
            for _ in range(2):
                e0, e1 = ev(), ev()
                e0.record(); pv.screen(); e1.record(); torch.cuda.synchronize()
                tv.append(e0.elapsed_time(e1))
            pv.verify(); mv = pv.eliminate()
            tfl = 18.0 * M * pairs / (min(tv) * 1e-3) / 1e12
            variants[vname] = {"screen_ms": min(tv), "fp64_tflops": tfl, "frac_of_fp64_peak": tfl / max(peaks["dmma"], peaks["dfma"]),
                               "mask_equal": bool(torch.equal(mv, mask))}
            del pv
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import rotor_molecules as rm
            from tscode_b200.torsion_module import TorsionInfo, prune_conformers_rmsd_rot_corr
            gb = json.load(open(os.path.join(ROOT, "tests", "golden", "rotcorr_big.json")))["fixtures"]
            name = [k for k in gb if k.startswith("tritbu63_s13_")][0]
            f4 = gb[name]
            g4 = np.load(os.path.join(ROOT, "tests", "golden", f"rotcorr_{name}.npz"))
            info = TorsionInfo([tuple(t) for t in f4["torsions"]], [tuple(a) for a in f4["angles"]],
                               g4["rot_masks"].astype(bool), g4["node_lists"].astype(bool))
            S4, at4 = rm.ensemble_tritbu63(f4["seed"], f4["N"])
            prune_conformers_rmsd_rot_corr(S4, at4, None, f4["thr"], torsion_info=info, max_structures=None)
            t0 = time.perf_counter()
            o4, m4 = prune_conformers_rmsd_rot_corr(S4, at4, None, f4["thr"], torsion_info=info, max_structures=None)
            t4 = time.perf_counter() - t0
            c4 = {"workload": f"BASELINE configs[3]: prune_conformers_rmsd_rot_corr, {f4['N']} structures x {S4.shape[1]} atoms, "
                              f"{len(f4['torsions'])} symmetric rotors, size guard lifted (the reference refuses > 750 structures)",
                  "ms_per_call": t4 * 1e3, "survivors": int(m4.sum()), "matches_guard_lifted_reference": mask_digest(m4) == f4["digest"],
                  "reference_wall_s": f4["wall_s"], "reference": f4["reference"]}
        except Exception as exc:
            c4 = {"error": repr(exc)}
        # the same prune on an ELONGATED molecule (reported beside the headline): the general form of the screen runs
        try:
            S_el = gen_ensemble(cfg["seed"], N, M, cfg["n_clusters"], scale=np.array([6.0, 2.0, 1.0]))
            pe = RmsdPruner(torch.from_numpy(S_el).to(dev), atomnos, thr, variant=args.variant, device=dev, ladder=args.ladder)
            el_ms = {"screen": [], "step": []}
            for it in range(5):
                flush.fill_(1)
                e = [ev() for _ in range(4)]
                e[0].record(); pe.pack(); e[1].record(); pe.screen(); e[2].record(); pe.verify(); m_el = pe.eliminate(); e[3].record()
                torch.cuda.synchronize()
                if it >= 2:
                    el_ms["screen"].append(e[1].elapsed_time(e[2])); el_ms["step"].append(e[0].elapsed_time(e[3]))
            shaped = {"workload": "as configs[2] with the base molecule scaled (6, 2, 1) along x, y, z (elongated: what real "
                                  "molecules look like; Samuelson's bound excludes nothing, the quartic test decides every pair)",
                      "screen_mode": pe.screen_mode, "ms_per_step": statistics.mean(el_ms["step"]),
                      "screen_ms": statistics.mean(el_ms["screen"]), "value": pairs / (statistics.mean(el_ms["step"]) * 1e-3),
                      "unit": UNIT, "survivors": int(m_el.sum()), "digest": mask_digest(m_el.cpu().numpy()),
                      "matches_reference": mask_digest(m_el.cpu().numpy()) == "478bc29df1e239da", **pe.stats_dict()}
            del pe, S_el
        except Exception as exc:
            shaped = {"error": repr(exc)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": DTYPE if args.variant == "screen" else "f64", "data": "synthetic",
            "config": config_dict(cfg, world, extra={"variant": args.variant, "screen_mode": pr.screen_mode,
                                                     "ladder": pr.ladder_used, "ladder_rounds": pr.rounds}),
            "wall_ms_total": wall_ms, "phase_ms": phase_max,
            "parity": {"survivors": int(mask_np.sum()), "digest": digest,
                       "matches_reference": (digest == cfg["digest"]) if cfg["digest"] else None,
                       "same_mask_on_every_rank": digests_equal, **verify_stats},
            "roofline": roofline, "cpu_baseline": cpu, "time_to_mask_speedup": tmask, "e2e": e2e, "clash": clash,
            "weak_scaling": weak, "variants": variants,
            "configs": {"C4_rot_corr": c4, "C5_embed_pipeline": c5}, "elongated_molecule": shaped,
            "gpu_launches": args.steps * ((2 + 1 + 2 if args.variant == "screen" else 1 + 1 + 2) +
                                          (1 if pr.ladder_used == "fused" else 3 * len(pr.rounds))),
            "clocks": clocks, "fp64_peaks_tflops": peaks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="screen", choices=["screen", "dmma", "fma"])
    ap.add_argument("--ladder", default="fused", choices=["fused", "bitrows"])
    ap.add_argument("--n-conformers", type=int, default=0, help="override N (testing only; invalid as a bench value)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-full-reference", action="store_true",
                    help="reference arm: skip the single full-size (50 000 x 80) time-to-mask run")
    ap.add_argument("--skip-extras", action="store_true",
                    help="profiling aid: only the timed prune steps (no e2e / clash / peak legs); not a bench value")
    args = ap.parse_args()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1: the CPU arm must use the host's cores whatever launched it
        n = str(host_threads())
        os.environ["OMP_NUM_THREADS"] = n
        os.environ["NUMBA_NUM_THREADS"] = n
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
