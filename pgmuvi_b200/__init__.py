"""pgmuvi_b200 - the exact-GP training hot path of ICSM/pgmuvi on B200 (sm_100a).

Public surface:

* ``train``                          - drop-in for ``pgmuvi.trainers.train`` (seam #1)
* ``B200ExactMarginalLogLikelihood`` - drop-in for ``gpytorch.mlls.ExactMarginalLogLikelihood``
  on the models of the path (seam #2)
* ``Lightcurve``                     - host-side mirror of the slice of ``pgmuvi.lightcurve.Lightcurve``
  that calls the path (seam #0; with the real pgmuvi installed one keeps its class)
* ``fit_batch`` / ``BatchEngine``    - the batch surface pgmuvi lacks (one launch per GPU)

The CUDA engine lives in ``libpgmuvi_b200.so`` (C ABI in ``include/pgmuvi_b200.h``, bound with
ctypes in ``_lib``); it is loaded on first use and there is no CPU fallback.
"""
from .batch import BatchEngine, HostBatch, fit_batch, gather_results, shard_range  # noqa: F401
from .lightcurve import Lightcurve  # noqa: F401
from .mll import B200ExactMarginalLogLikelihood, UnsupportedModelError, pack_model  # noqa: F401
from .trainers import train  # noqa: F401

__all__ = ["train", "fit_batch", "BatchEngine", "HostBatch", "B200ExactMarginalLogLikelihood",
           "Lightcurve", "pack_model", "UnsupportedModelError", "gather_results", "shard_range"]
