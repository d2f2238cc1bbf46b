#!/usr/bin/env python
"""bench.py — headline benchmark of the conformer-ensemble hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--variant dmma|fma]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the north-star target): prune_conformers_rmsd on a synthetic
clustered ensemble of 50 000 conformers x 80 heavy atoms, rmsd_thr 0.5 (gen_ensemble(3, 50000,
80, 5000)).  One step = one whole prune: repack, all-pairs similarity
(N(N-1)/2 = 1.25e9 pairs), exact verification, k-ladder elimination -> survivor mask.  At N > 1
the rows of the pair matrix are sharded block-cyclically over the ranks (strong scaling: the
total pair count is fixed) with an NCCL all-gather per elimination round.

Prints ONE JSON line (rank 0).  `value` = pairs/s with the ensemble already resident in HBM;
`e2e` = the same through the public drop-in `prune_conformers_rmsd(numpy, atomnos, thr)` with
host buffers (H2D of the structures and D2H of the mask inside the timed region);
`roofline` = the dominant kernel (all-pairs screen) against the self-measured FP64 ceiling;
`cpu_baseline` = the oracle C port of the reference's per-pair evaluation on the host cores;
`clash` = the secondary metric (BASELINE configs[1]: 100k two-fragment poses, fused transform
+ clash screen), poses/s.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tscode_b200.synth import gen_ensemble, gen_poses, mask_digest  # noqa: E402

C3 = dict(seed=3, N=50_000, M=80, n_clusters=5000, thr=0.5, digest="478bc29df1e239da", survivors=48867)
C2 = dict(seed=0, P=100_000, n_atoms=(50, 50), thresh=1.5, max_clashes=0, digest="6e7eb19c842b4798", passes=12695)
METRIC = "rmsd_pairs_per_s"
UNIT = "pairs/s"


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
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
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def fp64_peaks(torch):
    """Self-measured FP64 ceilings (tsc_bench_fp64): MEASURED_PEAKS.json has no FP64 entry."""
    import ctypes as C
    from tscode_b200._lib import check, lib, ptr, stream_ptr
    scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
    out = {}
    for kind, name in ((0, "dfma"), (1, "dmma"), (2, "mixed")):
        best = 0.0
        for _ in range(3):
            fl = (C.c_double * 2)()
            ms = C.c_float()
            check(lib().tsc_bench_fp64(kind, 4000, 2, 512, ptr(scratch), C.cast(fl, C.c_void_p),
                                       C.cast(C.byref(ms), C.c_void_p), stream_ptr()), "tsc_bench_fp64")
            best = max(best, (fl[0] + fl[1]) / (ms.value * 1e-3) / 1e12)
        out[name] = round(best, 2)
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ---------------------------------------------------------------------------------------------
def cpu_baseline_pairs(S, thr, target_s=12.0):
    """Oracle C port of rmsd_and_max + thresholds (rmsd_pruning.py:6-41,:75) on all host threads,
    over a bounded random sample of (i<j) pairs of the same ensemble."""
    from oracle import oracle_c
    N = S.shape[0]
    rng = np.random.default_rng(1)
    n_cal = 100_000

    def sample(n):
        i = rng.integers(0, N - 1, size=n); j = rng.integers(0, N, size=n)
        lo, hi = np.minimum(i, j), np.maximum(i, j)
        hi = np.where(lo == hi, hi + 1, hi)
        return lo.astype(np.int64), np.minimum(hi, N - 1).astype(np.int64)
    ii, jj = sample(n_cal)
    oracle_c.eval_pairs(S, thr, ii[:1000], jj[:1000])
    t0 = time.perf_counter(); oracle_c.eval_pairs(S, thr, ii, jj); dt = time.perf_counter() - t0
    n = int(min(max(n_cal * target_s / max(dt, 1e-6), n_cal), 5e8))
    ii, jj = sample(n)
    t0 = time.perf_counter(); hits = oracle_c.eval_pairs(S, thr, ii, jj); dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": oracle_c.num_threads(), "kind": "port",
            "sample": f"{n} random (i<j) pairs of the same {N}x{S.shape[1]} ensemble, rmsd_and_max + thresholds "
                      f"(oracle/oracle.c, OpenMP), {dt:.1f} s; {hits} similar"}, n, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle C port: the reference is Python +
    numba and cannot travel) on the host cores, same config / metric; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle_c
    cfg = dict(C3)
    if args.n_conformers:
        cfg.update(N=args.n_conformers, n_clusters=max(args.n_conformers // 10, 1))
    S = gen_ensemble(cfg["seed"], cfg["N"], cfg["M"], cfg["n_clusters"])
    base, n, dt = cpu_baseline_pairs(S, cfg["thr"], target_s=4.0)     # calibrates the per-step sample
    per_step = n
    rng = np.random.default_rng(2)
    N = cfg["N"]

    def step():
        i = rng.integers(0, N - 1, size=per_step); j = rng.integers(0, N, size=per_step)
        lo, hi = np.minimum(i, j), np.maximum(i, j)
        hi = np.minimum(np.where(lo == hi, hi + 1, hi), N - 1)
        t0 = time.perf_counter()
        oracle_c.eval_pairs(S, cfg["thr"], lo.astype(np.int64), hi.astype(np.int64))
        return time.perf_counter() - t0
    for _ in range(args.warmup):
        step()
    ts = [step() for _ in range(args.steps)]
    total = sum(ts)
    value = per_step * args.steps / total
    # time-to-the-same-mask of the lazy reference algorithm (it evaluates ~1.5 % of the pairs)
    lazy = None
    if args.lazy_n:
        Sl = gen_ensemble(cfg["seed"], args.lazy_n, cfg["M"], max(args.lazy_n // 10, 1))
        t0 = time.perf_counter(); m, ne, _ = oracle_c.prune_heavy(Sl, cfg["thr"]); tl = time.perf_counter() - t0
        lazy = {"N": args.lazy_n, "seconds": round(tl, 3), "pairs_evaluated": int(ne),
                "pairs_total": args.lazy_n * (args.lazy_n - 1) // 2, "survivors": int(m.sum())}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(cfg, args, extra={
                "reference_sample": f"{per_step} random (i<j) pairs per step, every pair fully evaluated "
                                    "(rmsd_and_max + thresholds), OpenMP over all host threads"}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle_c.num_threads(), "kind": "port",
                             "sample": f"{per_step} pairs/step x {args.steps} steps"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "lazy_reference_time_to_mask": lazy}
    print(json.dumps(line), flush=True)


def config_dict(cfg, args, extra=None):
    d = {"workload": f"BASELINE configs[2]: prune_conformers_rmsd all-pairs, {cfg['N']} conformers x {cfg['M']} "
                     f"heavy atoms, rmsd_thr {cfg['thr']}, gen_ensemble(seed={cfg['seed']}, n_clusters={cfg['n_clusters']})",
         "pairs_per_step": cfg["N"] * (cfg["N"] - 1) // 2,
         "parallelism": f"row-block-cyclic x{args.gpus}",
         "l2": "working set per step (96 MB structures + 96 MB packed + 313 MB similarity bits written) exceeds the "
               "126 MB L2; a 512 MB scratch write additionally flushes L2 between timed steps"}
    if extra:
        d.update(extra)
    return d


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from tscode_b200 import _host
    from tscode_b200.numba_functions import PoseBatch
    from tscode_b200.rmsd_pruning import RmsdPruner, prune_conformers_rmsd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
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
    pr = RmsdPruner(S_dev, atomnos, thr, variant=args.variant, rank=rank, world=world, device=dev, ladder=args.ladder)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    for _ in range(max(args.warmup, 3)):
        mask = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
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
    clocks = sampler.stop() if rank == 0 else None
    verify_stats = pr.stats_dict()          # counters of the last timed step (the extras below run more steps)
    if rank == 0 and world == 1 and not args.skip_extras:
        # The timed region above lasts a few tens of ms: nvidia-smi (100 ms period) sees it once at best, usually
        # between kernels.  Two untimed additions: (a) the same step back to back for ~0.6 s under a second sampler;
        # (b) the SM clock the screen kernel itself ran at, from clock64 / globaltimer stamps of CTA 0's first work
        # item (tsc_set_trace_buffer).  Neither enters any reported rate.
        from tscode_b200._lib import lib as _lib, ptr as _ptr
        s2 = ClockSampler(local)
        s2.start()
        t_end = time.perf_counter() + 0.6
        while time.perf_counter() < t_end:
            for _ in range(20):
                step()
            torch.cuda.synchronize()
        clocks["sustained"] = s2.stop()
        if args.variant in ("f16", "tf32"):
            trace = torch.zeros(8 * 96, dtype=torch.int64, device=dev)
            _lib().tsc_set_trace_buffer(_ptr(trace))
            pr.screen(); torch.cuda.synchronize()
            _lib().tsc_set_trace_buffer(None)
            tr = trace.cpu().numpy().reshape(96, 8)
            if tr[95, 7] > tr[8, 7] > 0:
                clocks["sm_mhz_in_screen_kernel"] = round(1e3 * float(tr[95, 3] - tr[8, 3]) / float(tr[95, 7] - tr[8, 7]), 1)
    step_ms = sum(t_events) / len(t_events)
    t = torch.tensor([step_ms, statistics.mean(phase_ms["screen"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, screen_ms = float(t[0]), float(t[1])
    pairs = N * (N - 1) // 2
    value = pairs / (step_ms * 1e-3)
    mask_np = mask.cpu().numpy()

    # ---- e2e through the public drop-in API, host buffers (N = 1: the API is single-GPU) -------
    e2e = None
    if args.skip_extras:
        pass
    elif world == 1:
        for _ in range(2):
            prune_conformers_rmsd(S_pinned_np, atomnos, thr)
        ts = []
        for _ in range(max(2, min(args.steps, 5))):
            flush.fill_(1); torch.cuda.synchronize()
            t0 = time.perf_counter()
            out, m2 = prune_conformers_rmsd(S_pinned_np, atomnos, thr)
            ts.append(time.perf_counter() - t0)
        assert np.array_equal(m2, mask_np)
        e2e = {"value": pairs / statistics.mean(ts), "unit": UNIT, "h2d_bytes_per_step": int(S_host.nbytes),
               "d2h_bytes_per_step": int(N), "ms_per_call": statistics.mean(ts) * 1e3,
               "api": "tscode_b200.rmsd_pruning.prune_conformers_rmsd(structures: numpy (pinned), atomnos, rmsd_thr) "
                      "-> (structures[mask], mask); includes the host-side structures[mask] gather"}
    else:
        # sharded API: every rank uploads the ensemble from its pinned host copy and reads the mask back
        ts = []
        for it in range(2 + max(2, min(args.steps, 5))):
            barrier()
            t0 = time.perf_counter()
            p2 = RmsdPruner(pinned, atomnos, thr, variant=args.variant, rank=rank, world=world, device=dev,
                            ladder=args.ladder)
            m2 = p2.run().cpu().numpy()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if it >= 2:
                ts.append(float(dt[0]))
            del p2
        e2e = {"value": pairs / statistics.mean(ts), "unit": UNIT, "h2d_bytes_per_step": int(S_host.nbytes) * world,
               "d2h_bytes_per_step": int(N) * world, "ms_per_call": statistics.mean(ts) * 1e3,
               "api": "RmsdPruner(pinned host structures, rank, world).run() on every rank (includes allocation)"}

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    peaks = fp64_peaks(torch) if not args.skip_extras else {"dfma": 34.2, "dmma": 37.1, "mixed": 36.7}
    mp, mp_src = measured_peaks()
    flops = 18.0 * M * pairs / world
    achieved = flops / (screen_ms * 1e-3) / 1e12
    if args.variant == "f16":
        peak = float(mp["bf16_tflops"])
        kname = "rmsd_ts_kernel<1,4,true>"
        psrc = (f"16-bit dense tensor peak = the {mp_src} cuBLAS bf16 burst figure in MEASURED_PEAKS.json "
                f"({mp['bf16_tflops']} TFLOP/s); the kernel issues kind::f16 MMAs of shape 128x48x16 (math time 24 cycles), "
                "which the tensor pipe executes at a fixed ~41-56 cycles each (tools/umma_probe.py; ncu: tensor pipe "
                "38.5 % active), see DESIGN.md 4.1b and 7")
    elif args.variant == "tf32":
        peak = mp["bf16_tflops"] / 2.0
        kname = "rmsd_ts_kernel<1,4,false>"
        psrc = (f"TF32 dense = half of the {mp_src} cuBLAS bf16 burst figure in MEASURED_PEAKS.json "
                f"({mp['bf16_tflops']} TFLOP/s)")
    else:
        peak = max(peaks["dmma"], peaks["dfma"])
        kname = f"rmsd_sim_kernel<{'ConsumerDMMA' if args.variant == 'dmma' else 'ConsumerFMA'}>"
        psrc = ("self-measured FP64 ceiling on this GPU in this run (tsc_bench_fp64: register-resident "
                f"DMMA.8x8x4 {peaks['dmma']} / DFMA {peaks['dfma']} / both {peaks['mixed']} TFLOP/s); "
                f"MEASURED_PEAKS.json ({mp_src}) has no FP64 entry")
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp) and world == 1 and not args.n_conformers:
        traffic = json.load(open(tp)).get(kname)
    roofline = {"bound": "tensor", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": psrc,
                "algorithmic_flop_per_pair": 18 * M, "kernel_ms": screen_ms,
                "kernel_share_of_step": screen_ms / step_ms}

    # ---- CPU baseline (rank 0, bounded sample) --------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu and not args.skip_extras:
        cpu, _, _ = cpu_baseline_pairs(S_host, thr, target_s=12.0)

    # ---- secondary metric: clash-checked poses/s (BASELINE configs[1]) ------------------------------------
    clash = None
    if world == 1 and not args.skip_extras:
        frags, conf, R, tt = gen_poses(C2["seed"], C2["P"], C2["n_atoms"])
        pb = PoseBatch(frags, conf, R, tt)
        for _ in range(3):
            v = pb.clash(C2["thresh"], C2["max_clashes"])
        torch.cuda.synchronize()
        ks = []
        for _ in range(10):
            flush.fill_(1)
            e0, e1 = ev(), ev()
            e0.record(); v = pb.clash(C2["thresh"], C2["max_clashes"]); e1.record()
            torch.cuda.synchronize()
            ks.append(e0.elapsed_time(e1))
        vn = v.cpu().numpy()
        Rp = torch.from_numpy(R).pin_memory(); tp = torch.from_numpy(tt).pin_memory()
        cp = torch.from_numpy(conf.astype(np.int32)).pin_memory()
        es = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            pb2 = PoseBatch(frags, cp, Rp, tp)
            v2 = pb2.clash(C2["thresh"], C2["max_clashes"]).cpu()
            es.append(time.perf_counter() - t0)
        kms = statistics.median(ks)
        clash = {"metric": "clash_checked_poses_per_s", "value": C2["P"] / (kms * 1e-3), "unit": "poses/s",
                 "kernel_ms": kms, "workload": "BASELINE configs[1]: 100k two-fragment poses (2 x 50 atoms), fused "
                 "rotation + clash test, thresh 1.5, max_clashes 0",
                 "parity": {"passes": int(vn.sum()), "digest": mask_digest(vn), "matches_reference": mask_digest(vn) == C2["digest"]},
                 "roofline": {"bound": "fp64", "kernel": "embed_clash_kernel<false,2>", "unit": "TFLOP/s",
                              "algorithmic_flop_per_pose": 8 * 50 * 50 + 18 * 100,
                              "achieved": (8 * 50 * 50 + 18 * 100) * C2["P"] / (kms * 1e-3) / 1e12, "peak": peaks["dfma"],
                              "frac": (8 * 50 * 50 + 18 * 100) * C2["P"] / (kms * 1e-3) / 1e12 / peaks["dfma"],
                              "note": "algorithmic work counts every inter-fragment pair (SURVEY 8d); the kernel leaves a "
                                      "pose at the first 8-atom check that exceeds max_clashes, so clashing poses do less"},
                 "e2e": {"value": C2["P"] / statistics.median(es), "unit": "poses/s",
                         "h2d_bytes_per_step": int(R.nbytes + tt.nbytes + conf.size * 4), "d2h_bytes_per_step": C2["P"]}}

    # ---- same prune on an ELONGATED molecule (reported beside the headline, never part of it) -----------------
    # C3's generator draws an isotropic Gaussian blob, the best case of the pre-screen's first stage
    # (Samuelson's bound); for elongated molecules the second, FP32-quartic stage decides every pair and the epilogue
    # rather than the tensor pipe sets the pace.
    shaped = None
    if world == 1 and not args.skip_extras and not args.n_conformers and args.variant in ("f16", "tf32"):
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
            shaped = {"workload": "as C3 with the base molecule scaled (6, 2, 1) along x, y, z (elongated)",
                      "ms_per_step": statistics.mean(el_ms["step"]), "screen_ms": statistics.mean(el_ms["screen"]),
                      "value": pairs / (statistics.mean(el_ms["step"]) * 1e-3), "unit": UNIT,
                      "survivors": int(m_el.sum()), "digest": mask_digest(m_el.cpu().numpy()),
                      # the live reference's mask of this ensemble (48 867 survivors, 76.9 s;
                      # tests/golden/prune_masks_aniso_big.json), also reproduced by the C oracle
                      "matches_reference": mask_digest(m_el.cpu().numpy()) == "478bc29df1e239da", **pe.stats_dict()}
            del pe, S_el
        except Exception as exc:                                  # an extra: never take the headline down with it
            shaped = {"error": repr(exc)}

    rounds = len(pr.rounds)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(cfg, args, extra={
                "variant": args.variant, "ladder": pr.ladder_used, "ladder_rounds": pr.rounds}),
            "wall_ms_total": wall_ms,
            "phase_ms": {k: statistics.mean(v) for k, v in phase_ms.items()},
            "parity": {"survivors": int(mask_np.sum()), "digest": mask_digest(mask_np),
                       "matches_reference": (mask_digest(mask_np) == cfg["digest"]) if cfg["digest"] else None,
                       **verify_stats},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clash": clash, "elongated_molecule": shaped,
            "gpu_launches": args.steps * ((5 if args.variant in ("tf32", "f16") else 4) +
                                          (1 if pr.ladder_used == "fused" else 3 * rounds)), "clocks": clocks,
            "fp64_peaks_tflops": peaks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="f16", choices=["dmma", "fma", "tf32", "f16"])
    ap.add_argument("--ladder", default="fused", choices=["fused", "bitrows"])
    ap.add_argument("--n-conformers", type=int, default=0, help="override N (testing only; invalid as a bench value)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--lazy-n", type=int, default=10000)
    ap.add_argument("--skip-extras", action="store_true",
                    help="profiling aid: only the timed prune steps (no e2e / clash / peak legs); not a bench value")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
