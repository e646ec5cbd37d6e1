#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric:

    MLL+grad evals/s, 4096 x (n=512, SM-4) light curves (config C2), fp64, N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch: K(x,x') build + Cholesky + solve +
log-det + full raw-parameter gradient for every light curve of the batch (one "eval" per
light curve).  Under torchrun the 4096 light curves of BASELINE config 2 are SPLIT over the
ranks (``batch.shard_range``: 512 per GPU at N = 8; strong scaling - SURVEY.md section 8e);
light curves are independent, so there is no collective inside the evaluation and the only
exchange is the all-gather of the [B/G, 1+P] results, which is inside both timed regions.
``--scaling weak`` gives every rank its own 4096 instead.  Rank 0 prints ONE JSON line, which
also carries the C5 survey batch (16384 sources x (4 bands x 256 epochs), sharded the same way).

`--impl reference` times the CPU restatement of the reference path (oracle/, torch CPU with
all host threads: GPyTorch itself is not installable in this image - SURVEY.md F3) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 512
Q_MIX = 4
LC_PER_GPU = 4096
F_EVAL = N_POINTS ** 3 + 4 * N_POINTS ** 2          # algorithmic flops / eval
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused kernel on the C2 batch
# (4096 light curves), from the `ncu --set full` capture summarised in
# profiles/r02w_ncu_summary.txt (45.84 GB + 10.40 GB); scales with light curves per GPU
TRAFFIC_BYTES_PER_LC = 56.245e9 / 4096      # ncu dram__bytes_read + write of one C2 launch (r02w)
TRAFFIC_SOURCE = "profiles/r02w_ncu_summary.txt"
KERNEL_NAME = "pgm::sm_mll_grad_kernel<0,4,1>"
PREWARM_STEPS = 30
_OUT = sys.stdout      # replaced in main() by a private handle to the real stdout
METRIC = "MLL+grad evals/s, 4096x n=512 SM-4 lightcurves"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lightcurves", type=int, default=LC_PER_GPU,
                    help="light curves: the global batch (strong scaling) or per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the batch is split over the ranks; weak: per rank")
    ap.add_argument("--c5-sources", type=int, default=16384, help="global sources of the C5 line")
    ap.add_argument("--cpu-sample", type=int, default=128,
                    help="light curves in the CPU sample (128 x n=512: about 1.5 s per step and mode; "
                         "r01's 32 gave +-8 %% run-to-run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short C1 / C3 / C4 / C5 timings appended to the N=1 line")
    return ap.parse_args()


def workload_config(args, world):
    strong = getattr(args, "scaling", "strong") == "strong"
    glob = args.lightcurves if strong else args.lightcurves * world
    return {"workload": "C2: batch of 4096 synthetic 1-D light curves n=512, SM-4 kernel, "
                        "FixedNoise likelihood, batched exact MLL+grad",
            "global_lightcurves": glob,
            "lightcurves_per_gpu": (glob + world - 1) // world,
            "n": N_POINTS, "num_mixtures": Q_MIX, "params_per_lightcurve": 13,
            "parallelism": (f"the {glob} independent light curves split contiguously over "
                            f"{world} GPU(s) (batch.shard_range)" if strong else
                            f"{args.lightcurves} independent light curves per GPU x {world}")
                           + "; no collective inside the evaluation, one all-gather of the "
                             "[B/G, 1+P] results per step inside the timed region",
            "data_note": "512 distinct seeded light curves, global index g uses seed 1000 + g % 512 "
                         "(the work per light curve does not depend on the data)",
            "l2": "256 MiB L2 flush between timed steps (outside the per-step CUDA events)",
            "prewarm_steps": PREWARM_STEPS}


# ---------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# CPU reference arm / baseline (oracle = torch-CPU restatement of the reference path)
# ---------------------------------------------------------------------------------------
def cpu_reference(sample, steps, warmup):
    """evals/s of the reference's CPU path on `sample` C2 light curves per step.
    Runs (i) the per-light-curve loop pgmuvi would run and (ii) torch batch mode, and
    reports the faster (all host threads)."""
    import torch
    from oracle.sm_gp import ModelSpec, batched_mll_and_grad, mll_and_grad_autograd
    from pgmuvi_b200 import synthetic as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bt = S.make_batch_1d(sample, N_POINTS, Q=Q_MIX)
    spec = ModelSpec(d=1, Q=Q_MIX)
    t = lambda a: torch.tensor(a, dtype=torch.float64)
    x, y, nz, raw, lb, ub = (t(bt[k]) for k in ("x", "y", "noise", "raw", "lb", "ub"))
    kinds = torch.tensor(bt["kinds"])

    def step_batched():
        return batched_mll_and_grad(x, y, nz, raw, kinds, lb, ub, spec)

    def step_loop():
        for b in range(sample):
            mll_and_grad_autograd(x[b], y[b], nz[b], raw[b], kinds, lb[b], ub[b], spec)

    results = {}
    for name, fn in (("batched", step_batched), ("per_lightcurve_loop", step_loop)):
        for _ in range(max(1, min(warmup, 2))):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = (time.perf_counter() - t0) / steps
        results[name] = sample / dt
    best = max(results, key=results.get)
    return results[best], best, results, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    val, mode, both, cores = cpu_reference(args.cpu_sample, steps, args.warmup)
    sample = (f"{args.cpu_sample} of the C2 light curves per step x {steps} steps, fp64, torch CPU "
              f"{cores} threads, best of {both} (oracle restatement of GPyTorch's Cholesky path; "
              "linear in the number of light curves)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT,
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * args.cpu_sample / val, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample, "mode": mode},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), file=_OUT, flush=True)


# ---------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------
def other_configs(dev, dmma_peak):
    """Short CUDA-event timings of the other BASELINE.json configs on one GPU (not part of the
    headline metric; DESIGN.md section 4 quotes them): C1 = 300-iteration Adam fit of the bundled
    AlfOri light curve, C3 = one MLL+grad of the n = 8000 2-D GP, C4 = n = 32768 SM-8 (C5 has its
    own sharded line, :func:`c5_line`)."""
    import warnings
    import torch
    from pgmuvi_b200 import ops, synthetic as S
    T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)
    out = {}
    try:
        tf32_peak = ops.peak_probe(4, 4096)
    except Exception:
        tf32_peak = None

    def ev_ms(fn, reps):
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    def single(bt, kind, Q, reps):
        x, y, nz, raw = (T(bt[k][0]) for k in ("x", "y", "noise", "raw"))
        kk, lo, hi = T(bt["kinds"], torch.int32), T(bt["lb"][0]), T(bt["ub"][0])
        n = x.shape[0]
        ms = ev_ms(lambda: ops.sm_mll_grad_large(x, y, nz, raw, kk, lo, hi, kind, Q, False, True),
                   reps)
        tf = (n ** 3 + 4 * n ** 2) / (ms * 1e-3) / 1e12
        rec = {"n": n, "ms_per_eval": ms, "tflops": tf, "frac_of_dmma_peak": tf / dmma_peak}
        # float32 model (the reference's default dtype) through pgm_sm_mll_grad_tf32x3_f32: the
        # trailing updates (n > 12800) and K~^-1 = X^T X on tcgen05 (3xTF32); same flop count
        f = lambda t: None if t is None else t.float().unsqueeze(0)
        x32 = f(x if x.dim() == 2 else x.unsqueeze(-1))
        ms32 = ev_ms(lambda: ops.sm_mll_grad_staged(x32, f(y), f(nz), f(raw), kk, lo.float(),
                                                    hi.float(), None, kind, Q, False, True,
                                                    tf32x3=True), reps)
        rec["f32_tf32x3"] = {"ms_per_eval": ms32, "tflops": tf * ms / ms32,
                             "speedup_vs_f64": ms / ms32,
                             "tf32_tcgen05_peak_tflops": tf32_peak,
                             "note": "P trailing updates (panel schedule) + G products on tcgen05 "
                                     "3xTF32; T phase, panels and solves FP64 DMMA"}
        return rec

    try:
        import tempfile
        from pgmuvi_b200.lightcurve import Lightcurve
        csv = S.alfori_csv(os.path.join(tempfile.gettempdir(), f"alfori_vband_{os.getpid()}.csv"))
        best = 1e30
        for _ in range(2):
            torch.manual_seed(0)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                lc = Lightcurve.from_csv(csv, xtransform="minmax", subsample_seed=0)
            lc.set_model("1D", num_mixtures=4)
            lc.set_default_constraints()
            lc.set_hypers({"covar_module.mixture_means":
                           torch.tensor([1 / 2100.0, 1 / 400.0, 1 / 1000.0, 1 / 200.0]),
                           "covar_module.mixture_scales":
                           torch.tensor([1.0e-4, 5.0e-4, 2.0e-4, 1.0e-3])})
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = lc.fit(optim="Adam", training_iter=300, lr=0.1)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out["C1_alfori_fit"] = {"n": 1000, "iterations": 300, "seconds": best,
                                "final_loss": float(res["loss"][-1])}
        # C2 for float32 models through the tensor-core path (staged engine; no trailing updates at
        # n = 512, so only the K~^-1 products of the gradient run on tcgen05)
        bt2 = S.make_batch_1d(64, N_POINTS, Q=Q_MIX, seed0=1000)
        rep = lambda a: torch.tensor(np.concatenate([np.asarray(a)] * 64, 0), dtype=torch.float32,
                                     device=dev)
        x2, y2, n2, r2, l2, u2 = (rep(bt2[k]) for k in ("x", "y", "noise", "raw", "lb", "ub"))
        k2 = T(bt2["kinds"], torch.int32)
        ms2 = ev_ms(lambda: ops.sm_mll_grad_staged(x2, y2, n2, r2, k2, l2, u2, None, 0, Q_MIX, False,
                                                   True, tf32x3=True), 3)
        tf2 = 4096 * (N_POINTS ** 3 + 4 * N_POINTS ** 2) / (ms2 * 1e-3) / 1e12
        out["C2_f32_tf32x3_staged"] = {
            "lightcurves": 4096, "n": N_POINTS, "ms_per_eval_batch": ms2,
            "evals_per_s": 4096 / ms2 * 1e3, "tflops": tf2,
            "tf32_tcgen05_peak_tflops": tf32_peak,
            "frac_of_3xtf32_peak": None if not tf32_peak else tf2 / (tf32_peak / 3.0),
            "note": "pgm_sm_mll_grad_tf32x3_f32: float buffers; Cholesky / inverse FP64 DMMA (dataflow "
                    "schedule), K~^-1 = X^T X on tcgen05 3xTF32 with the FP64 contraction epilogue; at "
                    "n = 512 the epilogue outweighs the 16 chunks of MMAs per tile, the fused FP64 "
                    "kernel (headline) is faster"}
        del x2, y2, n2, r2, l2, u2
        torch.cuda.empty_cache()
        # separable 2-D models (gps.py:1274-1342) through the fused kernel, C2-sized batch of
        # 4 bands x 128 epochs: SM-4(time) x ScaleKernel(RBF)(wavelength), and the reference's default
        # '2DSeparable' composition ScaleKernel(Matern-1.5)(time) x ScaleKernel(RBF)(wavelength)
        sep = {}
        for name, kind, mk in (("sm4_x_rbf", 3, lambda: S.make_batch_sep(64, 4, 128, Q=4, kind=3)),
                               ("matern15_x_rbf", 14,
                                lambda: S.make_batch_stat(64, 14, n_bands=4, n_per_band=128))):
            bs = mk()
            rep64 = lambda a: torch.tensor(np.concatenate([np.asarray(a)] * 64, 0),
                                           dtype=torch.float64, device=dev)
            xs, ys, ns, rs, ls, us = (rep64(bs[k]) for k in ("x", "y", "noise", "raw", "lb", "ub"))
            ks = T(bs["kinds"], torch.int32)
            Qs = bs.get("Q", 0)
            mss = ev_ms(lambda: ops.sm_mll_grad(xs, ys, ns, rs, ks, ls, us, None, kind, Qs, False,
                                                True), 3)
            tfs = 4096 * (N_POINTS ** 3 + 4 * N_POINTS ** 2) / (mss * 1e-3) / 1e12
            sep[name] = {"kernel_kind": kind, "lightcurves": 4096, "n": 512,
                         "ms_per_eval_batch": mss, "evals_per_s": 4096 / mss * 1e3, "tflops": tfs,
                         "frac_of_dmma_peak": tfs / dmma_peak}
            del xs, ys, ns, rs, ls, us
        out["separable_2d_4096x512"] = sep
        torch.cuda.empty_cache()
        out["C3_2d_n8000"] = single(S.make_batch_2d(1, 8, 1000, Q=4), 1, 4, 3)
        out["C4_1d_n32768_sm8"] = single(S.make_batch_1d(1, 32768, Q=8), 0, 8, 2)
    except Exception as exc:   # never let the side measurements break the headline line
        out["error"] = repr(exc)
    return out


def c5_line(args, dev, rank, world, dmma_peak):
    """BASELINE config 5: 16384 sources x (4 bands x 256 epochs = 1024 rows), 2-D SM-4, split
    over the ranks (2048 per GPU at N = 8) through BatchEngine with host buffers; the all-gather
    of the results is inside both timed regions.  32 distinct seeded sources are replicated."""
    import torch
    import torch.distributed as dist
    from pgmuvi_b200 import synthetic as S
    from pgmuvi_b200.batch import BatchEngine, HostBatch, gather_results, shard_range
    G5 = args.c5_sources
    a, z = shard_range(G5, rank, world)
    B5 = z - a
    bt = S.make_batch_2d(32, 4, 256, Q=4)
    idx = np.arange(a, z) % 32
    for k in ("x", "y", "noise", "raw", "lb", "ub"):
        bt[k] = bt[k][idx]
    hb = HostBatch.from_numpy(bt, pin=True)
    eng = BatchEngine(kind=1, Q=4, learn_noise=False, device=dev)
    counts = [shard_range(G5, r, world)[1] - shard_range(G5, r, world)[0] for r in range(world)]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    d = eng.upload(hb)
    reps = 2
    for _ in range(1):
        gather_results(torch.cat([t.unsqueeze(1) if t.dim() == 1 else t
                                  for t in eng.evaluate_device(d, True)[:2]], 1), counts)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        mll, grad, info = eng.evaluate_device(d, True)
        gather_results(torch.cat([mll.unsqueeze(1), grad], 1), counts)
    e1.record()
    sync()
    dev_ms = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        d = eng.upload(hb)
        mll, grad, info = eng.evaluate_device(d, True)
        allr = gather_results(torch.cat([mll.unsqueeze(1), grad], 1), counts)
        host = allr.cpu()
    sync()
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    tm = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = (float(v) for v in tm.tolist())
    tf = G5 * (1024 ** 3 + 4 * 1024 ** 2) / (dev_ms * 1e-3) / 1e12
    return {"workload": f"C5: {G5} sources x (4 bands x 256 epochs), 2-D SM-4, fp64, "
                        f"{B5} per GPU", "global_sources": G5, "ms_per_eval_batch": dev_ms,
            "evals_per_s": G5 / dev_ms * 1e3, "e2e_evals_per_s": G5 / e2e_ms * 1e3,
            "tflops_all_gpus": tf, "frac_of_dmma_peak": tf / (dmma_peak * world),
            "h2d_bytes_per_step": hb.h2d_bytes(), "d2h_bytes_per_step": host.numel() * 8,
            "info_nonzero": int((info != 0).sum().item())}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pgmuvi_b200 import ops, synthetic as S
    from pgmuvi_b200.batch import BatchEngine, HostBatch, gather_results, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    strong = args.scaling == "strong"
    G = args.lightcurves if strong else args.lightcurves * world      # global batch
    a, z = shard_range(G, rank, world) if strong else (rank * args.lightcurves,
                                                       (rank + 1) * args.lightcurves)
    B = z - a                                                          # this rank's shard
    counts = [shard_range(G, r, world)[1] - shard_range(G, r, world)[0] for r in range(world)] \
        if strong else [args.lightcurves] * world
    # 512 distinct seeded light curves of the C2 shape; global index g -> seed 1000 + g % 512
    distinct = min(G, 512)
    bt = S.make_batch_1d(distinct, N_POINTS, Q=Q_MIX, seed0=1000)
    idx = np.arange(a, z) % distinct
    for k in ("x", "y", "noise", "raw", "lb", "ub"):
        bt[k] = bt[k][idx]
    hb = HostBatch.from_numpy(bt, pin=True)
    eng = BatchEngine(kind=ops.KIND_SM1D, Q=Q_MIX, learn_noise=False, device=dev)
    d = eng.upload(hb)
    tail_k = eng.tail_split(B, 1)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        mll, grad, info = eng.evaluate_device(d, want_grad=True)
        return gather_results(torch.cat([mll.unsqueeze(1), grad], 1), counts), info

    # ---- device-resident throughput ("value") -------------------------------------------
    # a fresh box starts with cold clocks / power state: about one second of untimed pre-warm
    # steps before the W warm-up steps (config.prewarm_steps), so that the K timed steps are not
    # the ones that absorb the ramp (one r01f run measured 38 ms instead of 31 ms per step
    # right after start-up)
    for _ in range(PREWARM_STEPS):
        step_device()
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                                   # L2 flush, outside the timed events
        ev[i][0].record()
        kev[i][0].record()
        mll, grad, info = eng.evaluate_device(d, want_grad=True)
        kev[i][1].record()
        gather_results(torch.cat([mll.unsqueeze(1), grad], 1), counts)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a_.elapsed_time(b_) for a_, b_ in ev]
    kern_ms = [a_.elapsed_time(b_) for a_, b_ in kev]
    total_ms = float(sum(step_ms))
    bad = int((info != 0).sum().item())

    # ---- end to end through the public host API ("e2e") -----------------------------------
    # pinned host inputs -> H2D -> kernel -> all-gather of the results -> D2H of all of them
    def step_e2e():
        dd = eng.upload(hb)
        m_, g_, i_ = eng.evaluate_device(dd, want_grad=True)
        allr = gather_results(torch.cat([m_.unsqueeze(1), g_], 1), counts)
        out = (eng._host_out("all", allr), eng._host_out("info", i_))
        torch.cuda.current_stream().synchronize()
        return out

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_e2e()
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    d2h = sum(t.numel() * t.element_size() for t in out)

    tm = torch.tensor([total_ms, e2e_s * 1e3, float(np.mean(kern_ms))], dtype=torch.float64,
                      device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms_avg = (float(v) for v in tm.tolist())

    dmma = ops.peak_probe(0, 8192)
    dfma = ops.peak_probe(1, 8192)
    del d, flush
    torch.cuda.empty_cache()
    c5 = None
    if not args.no_other_configs:
        try:
            c5 = c5_line(args, dev, rank, world, dmma)
        except Exception as exc:       # never let a side measurement break the headline line
            c5 = {"error": repr(exc)}

    if rank == 0:
        evals = G * args.steps
        value = evals / (total_ms * 1e-3)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        achieved = B * F_EVAL / (kern_ms_avg * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": KERNEL_NAME,
                    "achieved": achieved, "peak": dmma, "unit": "TFLOP/s",
                    "frac": achieved / dmma,
                    "peak_source": "self-measured FP64 DMMA (mma.sync.m8n8k4.f64) probe in this "
                                   "run; MEASURED_PEAKS.json holds no fp64 figure "
                                   f"(its bf16 {peaks.get('bf16_tflops')} TF / HBM "
                                   f"{peaks.get('hbm_gbs')} GB/s do not bound an fp64 kernel); "
                                   "nominal B200 FP64 37 TFLOP/s",
                    "dfma_probe_tflops": dfma,
                    "lightcurves_per_launch": B,
                    "algorithmic_flops_per_launch": B * F_EVAL,
                    "kernel_ms_avg": kern_ms_avg,
                    "traffic": TRAFFIC_BYTES_PER_LC * B,
                    "traffic_unit": "bytes per launch (ncu dram read+write, " + TRAFFIC_SOURCE + ")",
                    "algorithmic_bytes_per_launch": B * (3 * N_POINTS + 2 * 13 + 1) * 8}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
                "e2e": {"value": evals / (e2e_ms * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": hb.h2d_bytes(), "d2h_bytes_per_step": d2h,
                        "includes": "H2D of the shard, kernel, all-gather of the results over the "
                                    "ranks, D2H of all gathered results"},
                # one fused-kernel launch per step, plus the 7 staged-engine launches of the
                # tail-balancing split (batch.BatchEngine.tail_split) when the shard size needs it
                "gpu_launches": args.steps * (1 + (7 if tail_k else 0)),
                "tail_split_lightcurves": tail_k, "roofline": roofline, "clocks": clocks,
                "cholesky_info_nonzero": bad, "wall_s_timed_loop": t_wall,
                "step_ms": [round(v, 3) for v in step_ms], "c5": c5}
        if world == 1 and not args.no_cpu_baseline:
            val, mode, both, cores = cpu_reference(args.cpu_sample, 2, 1)
            line["cpu_baseline"] = {
                "value": val, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{args.cpu_sample} of the C2 light curves x 2 steps, fp64, torch CPU "
                          f"{cores} threads, best of {both}", "mode": mode}
        if world == 1 and not args.no_other_configs:
            line["other_configs"] = other_configs(dev, dmma)
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # the contract is ONE JSON line on stdout: keep a private handle to the real stdout for it and
    # send everything libraries print to fd 1 (NCCL's version banner, warnings) to stderr
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
