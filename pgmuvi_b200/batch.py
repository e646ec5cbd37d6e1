"""Batched engine: many independent light curves on one GPU (or sharded over ranks).

pgmuvi has no batch API (one ``Lightcurve`` = one source); surveys loop over
``Lightcurve.fit``.  ``BatchEngine`` is the B200-native replacement of that loop for the
hot path: ``evaluate`` = one ``model(train_x) -> -mll -> backward`` pass for every light
curve (pgmuvi/trainers.py:179-181), ``fit`` = the whole ``trainers.train`` loop
(trainers.py:177-207) on device.  Host buffers in, host buffers out; device staging buffers
are owned by the engine and reused.

Multi-GPU: light curves are independent, so ranks take contiguous shards and nothing is
exchanged inside the loop; ``gather_results`` does the single end-of-job all-gather.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import ops
from ._lib import (KIND_SM1D, OPT_ADAM, OPT_ADAMW, OPT_SGD)

OPTIM_KINDS = {"SGD": OPT_SGD, "Adam": OPT_ADAM, "AdamW": OPT_ADAMW}


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced split of ``total`` light curves: returns (start, stop)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def optimizer_defaults(optim: str, eps: float = 1e-8):
    """torch.optim defaults as trainers.py:141-147 constructs them."""
    if optim not in OPTIM_KINDS:
        raise ValueError("optim must be either 'SGD', 'Adam', 'AdamW'")
    return dict(optim_kind=OPTIM_KINDS[optim], beta1=0.9, beta2=0.999, eps=eps,
                weight_decay=0.01 if optim == "AdamW" else 0.0)


@dataclass
class HostBatch:
    """Host-side (preferably pinned) inputs of a batch, float64 / int32, C-contiguous."""
    x: torch.Tensor                      # [B, n, d]
    y: torch.Tensor                      # [B, n]
    noise: Optional[torch.Tensor]        # [B, n] variance or None
    raw: torch.Tensor                    # [B, P]
    kinds: torch.Tensor                  # [P] int32
    lb: torch.Tensor                     # [B, P] or [P]
    ub: torch.Tensor
    n_valid: Optional[torch.Tensor] = None  # [B] int32

    @staticmethod
    def from_numpy(d, pin=True):
        def t(a, dt=torch.float64):
            if a is None:
                return None
            out = torch.as_tensor(np.ascontiguousarray(a), dtype=dt)
            return out.pin_memory() if pin and torch.cuda.is_available() else out
        return HostBatch(t(d["x"]), t(d["y"]), t(d.get("noise")), t(d["raw"]),
                         t(d["kinds"], torch.int32), t(d["lb"]), t(d["ub"]),
                         t(d.get("n_valid"), torch.int32))

    def slice(self, start, stop):
        s = lambda a: None if a is None else a[start:stop]
        per_lc = self.lb.dim() == 2
        return HostBatch(s(self.x), s(self.y), s(self.noise), s(self.raw), self.kinds,
                         s(self.lb) if per_lc else self.lb, s(self.ub) if per_lc else self.ub,
                         s(self.n_valid))

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in
                   (self.x, self.y, self.noise, self.raw, self.kinds, self.lb, self.ub,
                    self.n_valid) if t is not None)


class BatchEngine:
    """Owns the device buffers for batches of one model family on one GPU."""

    def __init__(self, kind=KIND_SM1D, Q=4, learn_noise=False, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("pgmuvi_b200.BatchEngine needs a CUDA device (no CPU fallback)")
        self.kind, self.Q, self.learn_noise = kind, Q, learn_noise
        self.device = torch.device(device if device is not None
                                   else f"cuda:{torch.cuda.current_device()}")
        self._dev = {}
        self._out = {}

    # -- staging -------------------------------------------------------------------------
    def _stage(self, name, host):
        if host is None:
            return None
        buf = self._dev.get(name)
        if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
            buf = torch.empty(host.shape, dtype=host.dtype, device=self.device)
            self._dev[name] = buf
        buf.copy_(host, non_blocking=True)
        return buf

    def upload(self, hb: HostBatch):
        """Host -> device copy of one batch (async on the current stream)."""
        return dict(x=self._stage("x", hb.x), y=self._stage("y", hb.y),
                    noise=self._stage("noise", hb.noise), raw=self._stage("raw", hb.raw),
                    kinds=self._stage("kinds", hb.kinds), lb=self._stage("lb", hb.lb),
                    ub=self._stage("ub", hb.ub), n_valid=self._stage("n_valid", hb.n_valid))

    def _host_out(self, name, dev):
        buf = self._out.get(name)
        if buf is None or buf.shape != dev.shape or buf.dtype != dev.dtype:
            buf = torch.empty(dev.shape, dtype=dev.dtype).pin_memory()
            self._out[name] = buf
        buf.copy_(dev, non_blocking=True)
        return buf

    # -- device-resident calls -------------------------------------------------------------
    def evaluate_device(self, d, want_grad=True):
        return ops.sm_mll_grad(d["x"], d["y"], d["noise"], d["raw"], d["kinds"], d["lb"],
                               d["ub"], d["n_valid"], self.kind, self.Q, self.learn_noise,
                               want_grad)

    def fit_device(self, d, maxiter=300, miniter=None, stop=1e-5, lr=0.1, optim="AdamW",
                   eps=1e-8, stopavg=30, keep_history=True):
        """trainers.train for every light curve; ``d['raw']`` is updated in place.
        Defaults are ``Lightcurve.fit``'s (lightcurve.py:5223-5229: AdamW, lr 0.1, 300
        iterations, stop 1e-5, stopavg 30, miniter=None -> training_iter)."""
        od = optimizer_defaults(optim, eps)
        miniter = maxiter if miniter is None else miniter
        return ops.sm_fit(d["x"], d["y"], d["noise"], d["raw"], d["kinds"], d["lb"], d["ub"],
                          d["n_valid"], self.kind, self.Q, self.learn_noise, od["optim_kind"],
                          float(lr), od["beta1"], od["beta2"], od["eps"], od["weight_decay"],
                          int(maxiter), int(miniter), float(stop or 0.0), int(stopavg),
                          bool(keep_history))

    # -- host in / host out (the call a user makes) ---------------------------------------
    def evaluate(self, hb: HostBatch, want_grad=True):
        """MLL (+ gradient w.r.t. raw parameters) of every light curve in the batch.
        Returns pinned host tensors (mll [B], grad [B,P], info [B]); synchronises."""
        d = self.upload(hb)
        mll, grad, info = self.evaluate_device(d, want_grad)
        out = (self._host_out("mll", mll), self._host_out("grad", grad),
               self._host_out("info", info))
        torch.cuda.current_stream().synchronize()
        return out

    def fit(self, hb: HostBatch, **kw):
        """Fit every light curve; returns dict(loss [iters,B], raw [B,P], raw_hist, n_iter,
        info) on the host."""
        d = self.upload(hb)
        loss, raw_hist, n_iter, info = self.fit_device(d, **kw)
        out = dict(loss=loss.cpu(), raw=d["raw"].cpu(), raw_hist=raw_hist.cpu(),
                   n_iter=n_iter.cpu(), info=info.cpu())
        return out


def gather_results(local: torch.Tensor, counts=None):
    """All-gather per-light-curve results [B_local, ...] over the ranks of the default
    process group (NCCL on GPUs, gloo on CPU); shards may be uneven (``counts`` = list of
    per-rank sizes, needed only then).  No-op without torch.distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None or len(set(counts)) == 1:
        out = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(out, local.contiguous())
        return torch.cat(out, 0)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], 0)
