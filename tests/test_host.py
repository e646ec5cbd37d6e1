"""Host-side logic on CPU: the `train` seam (schema of pgmuvi/trainers.py:167-209), the
Lightcurve mirror (likelihood / constraint / hyper set-up that fixes the parameterisation), the
GPyTorch-shaped packing, and the multi-rank sharding (gloo, world_size 2).  The CUDA engine is
replaced by the oracle here ONLY to drive the host code; the GPU twins of these tests are in
test_gpu_parity.py / test_gpu_train.py."""
import os

import numpy as np
import pytest
import torch

from oracle import ModelSpec, train_loop
from pgmuvi_b200 import gp, trainers
from pgmuvi_b200.batch import gather_results, shard_range
from pgmuvi_b200.lightcurve import Lightcurve, UnsupportedModel
from pgmuvi_b200.mll import UnsupportedModelError, pack_model

OPT_NAMES = {0: "SGD", 1: "Adam", 2: "AdamW"}


def fake_sm_fit(x, y, fixed_noise, raw, kinds, lb, ub, n_valid, kind, Q, learn_noise, optim_kind,
                lr, b1, b2, eps, wd, maxiter, miniter, stop, stopavg, keep_history):
    """Oracle stand-in with pgm_sm_fit_f64's contract (include/pgmuvi_b200.h)."""
    B, P = raw.shape
    d = x.shape[-1]
    spec = ModelSpec(d=d, Q=Q, kind=kind, learn_noise=learn_noise)
    loss = torch.full((maxiter, B), float("nan"), dtype=torch.float64)
    hist = torch.zeros(maxiter + 1, B, P, dtype=torch.float64)
    n_iter = torch.zeros(B, dtype=torch.int32)
    for b in range(B):
        n = x.shape[1] if n_valid is None else int(n_valid[b])
        res = train_loop(x[b, :n], y[b, :n], None if fixed_noise is None else fixed_noise[b, :n],
                         raw[b],
                         kinds, lb if lb.dim() == 1 else lb[b], ub if ub.dim() == 1 else ub[b],
                         spec, maxiter=maxiter, miniter=miniter, stop=stop or None, lr=lr,
                         optim=OPT_NAMES[optim_kind], eps=eps, stopavg=stopavg)
        k = len(res["loss"])
        loss[:k, b] = torch.tensor(np.array(res["loss"], dtype=float))
        hist[:k + 1, b] = torch.tensor(np.stack(res["raw"]))
        n_iter[b] = k
        raw[b] = hist[k, b]
    return loss, hist, n_iter, torch.zeros(B, dtype=torch.int32)


@pytest.fixture
def cpu_engine(monkeypatch):
    monkeypatch.setattr(trainers.ops, "sm_fit", fake_sm_fit)
    monkeypatch.setattr(trainers, "engine_device", lambda t=None: torch.device("cpu"))


def _lc_1d(n=60, yerr=True, seed=0, **kw):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(2450000.0, 2450400.0, n))
    y = np.sin(2 * np.pi * t / 57.0) + 0.1 * rng.standard_normal(n)
    kw.setdefault("xtransform", "minmax")
    return Lightcurve(t, y, yerr=np.full(n, 0.1) if yerr else None, **kw), t, y


def test_train_returns_the_reference_schema(cpu_engine):
    lc, t, y = _lc_1d()
    lc.set_model("1D", num_mixtures=2)
    lc.set_default_constraints()
    lc.set_hypers({"covar_module.mixture_means": torch.tensor([1 / 57.0, 1 / 110.0])})
    before = {k: v.clone() for k, v in lc.get_parameters().items()}
    res = trainers.train(lc, maxiter=5, miniter=5, stop=1e-5, lr=0.1, optim="AdamW", stopavg=3)
    assert set(res) == {"loss", "delta_loss"} | set(before)              # trainers.py:167-171
    assert len(res["loss"]) == 5 and len(res["delta_loss"]) == 4
    for k, v0 in before.items():
        assert len(res[k]) == 6                                          # initial value first
        assert isinstance(res[k][0], np.ndarray) and res[k][0].shape == tuple(v0.shape)
        assert np.allclose(res[k][0], v0.numpy(), rtol=1e-6)
    after = lc.get_parameters()                                          # params were updated
    for k in before:
        assert np.allclose(res[k][-1], after[k].numpy(), rtol=1e-6)
    assert set(before) == {"mean_module.constant", "covar_module.mixture_weights",
                           "covar_module.mixture_means", "covar_module.mixture_scales"}
    assert res["delta_loss"][0] == pytest.approx(float(res["loss"][1] - res["loss"][0]))
    # reported frequencies are in raw (day^-1) units: 1/xtransform.inverse(1/v, shift=False)
    assert res["covar_module.mixture_means"][0].reshape(-1)[0] == pytest.approx(1 / 57.0, rel=1e-4)


def test_train_early_stop_and_argument_errors(cpu_engine, capsys):
    lc, *_ = _lc_1d(n=40)
    lc.set_model("1D", num_mixtures=1)
    lc.set_default_constraints()
    res = trainers.train(lc, maxiter=40, miniter=2, stop=10.0, lr=1e-3, optim="Adam", stopavg=3)
    assert len(res["loss"]) == 4            # stop and i > miniter  ->  breaks at i = 3
    assert "so we will end training here" in capsys.readouterr().out
    with pytest.raises(ValueError):
        trainers.train(model=lc.model)                              # trainers.py:98-103
    with pytest.raises(NotImplementedError):
        trainers.train(lc, lossfn="elbo")
    with pytest.raises(ValueError):
        trainers.train(lc, optim="LBFGS")
    with pytest.raises(NotImplementedError):
        trainers.train(lc, optim="NUTS")


def test_train_without_lightcurve_records_raw_parameters(cpu_engine):
    lc, *_ = _lc_1d(n=30, yerr=False)
    lc.set_model("1D", num_mixtures=1)
    res = trainers.train(model=lc.model, likelihood=lc.likelihood,
                         train_x=lc._xdata_transformed, train_y=lc._ydata_transformed,
                         maxiter=3, optim="SGD", lr=1e-3)
    names = [n for n, _ in lc.model.named_parameters()]
    assert names[0] == "likelihood.noise_covar.raw_noise"             # SURVEY A.9 order
    assert set(res) == {"loss", "delta_loss"} | set(names)
    assert all(len(res[n]) == 4 for n in names)


def test_likelihood_and_default_constraints_follow_the_reference():
    lc, t, y = _lc_1d()
    lc.set_likelihood()
    assert isinstance(lc.likelihood, gp.FixedNoiseGaussianLikelihood)
    assert torch.allclose(lc.likelihood.noise, torch.full((60,), 0.1, dtype=torch.float32) ** 2)   # tests/tests.py:144-154
    lc.set_likelihood(variance=True)
    assert torch.allclose(lc.likelihood.noise, torch.full((60,), 0.1, dtype=torch.float32))        # tests/tests.py:156-167
    lc.set_model("1D", likelihood="learn", num_mixtures=3)
    lc.set_default_constraints()
    cov = lc.model.covar_module
    assert type(cov.raw_mixture_means_constraint).__name__ == "GreaterThan"    # 1/span, :3920-3932
    assert float(cov.raw_mixture_means_constraint.lower_bound) == pytest.approx(1.0)
    nc = lc.likelihood.second_noise_covar.raw_noise_constraint
    assert type(nc).__name__ == "Interval"
    assert float(nc.lower_bound) == pytest.approx(1e-4)                      # min(1e-4, 0.1/10)
    assert float(nc.upper_bound) == pytest.approx(float(lc._ydata_transformed.std()))
    mc = lc.model.mean_module.raw_constant_constraint
    assert float(mc.lower_bound) == pytest.approx(float(y.min()), rel=1e-6)
    # re-registering a constraint keeps the RAW value (tutorial cells 32-37)
    assert float(lc.model.mean_module.raw_constant) == 0.0
    # no yerr -> GaussianLikelihood with noise >= 1e-4 * std(y)
    lc2, *_ = _lc_1d(yerr=False)
    lc2.set_model("1D", num_mixtures=2)
    lc2.set_default_constraints()
    assert isinstance(lc2.likelihood, gp.GaussianLikelihood)
    nb = lc2.likelihood.noise_covar.raw_noise_constraint
    assert float(nb.lower_bound) == pytest.approx(1e-4 * float(lc2._ydata_transformed.std()))


def test_2d_constraints_and_dimension_checks():
    rng = np.random.default_rng(1)
    x = np.concatenate([np.stack([np.sort(rng.uniform(0, 300, 30)), np.full(30, wl)], 1)
                        for wl in (0.8, 1.2, 2.2)])
    y = rng.standard_normal(90)
    lc = Lightcurve(x, y, yerr=np.full(90, 0.05), xtransform="minmax")
    with pytest.raises(ValueError):                           # tests/test_2d_integration.py:167-186
        lc.set_model("1D", num_mixtures=2)
    lc.set_model("2D", num_mixtures=3)
    lc.set_default_constraints()
    con = lc.model.covar_module.raw_mixture_means_constraint
    assert type(con).__name__ == "Interval"                   # tests/test_2d_constraints.py:67-78
    ts = np.sort(lc._xdata_transformed[:, 0].numpy())
    dt = np.diff(ts)
    assert float(con.upper_bound) == pytest.approx(1 / (2 * dt[dt > 0].min()), rel=1e-5)
    assert lc.model.covar_module.raw_mixture_means.shape == (3, 1, 2)
    lc.set_hypers({"covar_module.mixture_means": torch.tensor([[0.01, 1.0], [0.02, 1.0],
                                                               [0.03, 1.0]])})
    assert lc.model.covar_module.mixture_means.shape == (3, 1, 2)          # keeps [Q,1,2]
    pk = pack_model(lc.model)
    assert (pk.kind, pk.Q, pk.d, pk.P) == (1, 3, 2, 1 + 3 + 12)
    with pytest.raises(UnsupportedModel):
        lc.set_model("1DSKI")                    # approximate (KISS-GP) models: outside the path
    lc.set_model("2DLinear", num_mixtures=3)     # non-constant means stay on the host
    pk = pack_model(lc.model)
    assert pk.external_mean and pk.P == 1 + 3 + 12 and int(pk.kinds[0]) == 0
    assert float(pk.raw()[0]) == 0.0 and pk.names[1] == "covar_module.raw_mixture_weights"


def test_separable_models_pack_onto_the_separable_kinds():
    """pgmuvi/gps.py:1274-1423 (SeparableGPModel / AchromaticGPModel / WavelengthDependentGPModel
    with the spectral-mixture time kernel): parameter names as GPyTorch registers them under
    ``covar_module.kernels.{0,1}``, packed as [mean | w | mu | sigma | (noise) | lam]."""
    rng = np.random.default_rng(2)
    x = np.concatenate([np.stack([np.sort(rng.uniform(0, 300, 25)), np.full(25, wl)], 1)
                        for wl in (0.8, 1.2, 2.2)])
    y = rng.standard_normal(75)
    expect = {"rbf": (3, 2), "matern": (4, 2), "rq": (5, 3)}
    for wk, (kind, nl) in expect.items():
        lc = Lightcurve(x, y, yerr=np.full(75, 0.05), xtransform="minmax")
        lc.set_model("2DWavelengthDependent", num_mixtures=3, wavelength_kernel_type=wk,
                     time_kernel_type="sm", mean_module="constant")
        lc.set_default_constraints()
        pk = pack_model(lc.model)
        assert (pk.kind, pk.Q, pk.d, pk.P) == (kind, 3, 2, 1 + 9 + nl)
        assert pk.names[:4] == ["mean_module.raw_constant",
                                "covar_module.kernels.0.raw_mixture_weights",
                                "covar_module.kernels.0.raw_mixture_means",
                                "covar_module.kernels.0.raw_mixture_scales"]
        assert pk.names[4:6] == ["covar_module.kernels.1.raw_outputscale",
                                 "covar_module.kernels.1.base_kernel.raw_lengthscale"]
        # the time-kernel means get the 2-D Interval constraint, the wavelength kernel keeps
        # GPyTorch's Positive defaults; the initial lengthscale is max(span/2, 1) (gps.py:1576)
        assert type(lc.model.covar_module.kernels[0].raw_mixture_means_constraint).__name__ == "Interval"
        assert float(lc.model.covar_module.kernels[1].base_kernel.lengthscale) == pytest.approx(1.0)
        assert pk.kinds.tolist()[-nl:] == [1] * nl
        keys = set(lc.get_parameters())
        assert "covar_module.kernels.1.base_kernel.lengthscale" in keys
        if wk == "rq":       # the reference's str.lstrip("raw_") quirk
            assert "covar_module.kernels.1.base_kernel.lpha" in keys
    lc = Lightcurve(x, y, xtransform="minmax")              # Gaussian likelihood: learned noise
    lc.set_model("2DWavelengthDependent")                   # the reference's defaults, gps.py:1566-1609
    pk = pack_model(lc.model)
    assert (pk.kind, pk.Q, pk.external_mean) == (8 + 5 * 1 + 1, 0, True)   # Matern x RBF, quad mean
    assert type(lc.model.mean_module).__name__ == "CustomQuadConstantMean"
    lc.set_model("2DAchromatic", num_mixtures=2, time_kernel_type="sm")
    pk = pack_model(lc.model)
    assert (pk.kind, pk.P, pk.learn_noise) == (6, 1 + 6 + 1 + 1, True)
    assert pk.names[-2:] == ["likelihood.noise_covar.raw_noise",
                             "covar_module.kernels.1.raw_constant"]
    with pytest.raises(ValueError):
        lc.set_model("2DAchromatic", time_kernel_type="nonsense")
    lc.set_model("2DAchromatic", time_kernel_type="quasi_periodic")   # gps.py:915-935
    pk = pack_model(lc.model)
    assert (pk.kind, pk.Q, pk.P) == (8 + 5 * 2 + 4, 0, 1 + 1 + 4 + 1)
    assert pk.names[2:6] == ["covar_module.kernels.0.raw_outputscale",
                             "covar_module.kernels.0.base_kernel.kernels.0.raw_lengthscale",
                             "covar_module.kernels.0.base_kernel.kernels.0.raw_period_length",
                             "covar_module.kernels.0.base_kernel.kernels.1.raw_lengthscale"]
    # stationary time kernels (N3): the reference's default Matern time kernel packs onto the
    # kinds 8 + 5 TK + WK with no mixtures
    lc.set_model("2DAchromatic", time_kernel_type="matern")
    pk = pack_model(lc.model)
    assert (pk.kind, pk.Q, pk.d, pk.P) == (8 + 5 * 1 + 4, 0, 2, 1 + 1 + 2 + 1)
    lc.set_model("2DSeparable")                              # gps.py:1316-1319: Matern x RBF
    pk = pack_model(lc.model)
    assert (pk.kind, pk.P) == (8 + 5 * 1 + 1, 1 + 1 + 4)
    assert pk.names[2:] == ["covar_module.kernels.0.raw_outputscale",
                            "covar_module.kernels.0.base_kernel.raw_lengthscale",
                            "covar_module.kernels.1.raw_outputscale",
                            "covar_module.kernels.1.base_kernel.raw_lengthscale"]
    with pytest.raises(ValueError):
        Lightcurve(x[:, 0], y).set_model("2DSeparable")


def test_pack_model_rejects_models_outside_the_path():
    lc, *_ = _lc_1d(n=20)
    lc.set_model("1D", num_mixtures=2)
    pk = pack_model(lc.model)
    assert pk.names == ["mean_module.raw_constant", "covar_module.raw_mixture_weights",
                        "covar_module.raw_mixture_means", "covar_module.raw_mixture_scales"]
    assert pk.fixed_noise is not None and not pk.learn_noise

    lc.model.mean_module = None
    with pytest.raises(UnsupportedModelError):
        pack_model(lc.model)


def test_shard_range_covers_everything_once():
    for total, world in ((4096, 8), (10, 3), (5, 8), (0, 2)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 11                                                # uneven shards: 6 + 5
    a, b = shard_range(total, rank, world)
    local = torch.arange(a, b, dtype=torch.float64).unsqueeze(1) * torch.ones(1, 3)
    counts = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0]
              for r in range(world)]
    full = gather_results(local, counts)
    even = gather_results(torch.full((4, 2), float(rank)))
    q.put((rank, full[:, 0].tolist(), even[:, 0].tolist()))
    dist.destroy_process_group()


def test_gather_results_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, full, even in outs:
        assert full == [float(i) for i in range(11)]
        assert even == [0.0] * 4 + [1.0] * 4


def _ragged_lcs(ns=(50, 64, 37), seed=4, **kw):
    rng = np.random.default_rng(seed)
    out = []
    for n in ns:
        per = rng.uniform(40, 90)
        t = np.sort(rng.uniform(2450000.0, 2450000.0 + 6 * per, n))
        y = np.sin(2 * np.pi * t / per) + 0.1 * rng.standard_normal(n)
        kw.setdefault("xtransform", "minmax")
        out.append((Lightcurve(t, y, yerr=np.full(n, 0.1), **kw).double(), per))
    return out


def test_fit_batch_equals_a_loop_over_lightcurve_fit(cpu_engine):
    """fit_batch packs a ragged batch with the SAME host set-up as Lightcurve.fit: per light
    curve the loss history and fitted parameters equal the single-source call."""
    from pgmuvi_b200.batch import fit_batch
    single = []
    torch.manual_seed(7)                     # initialize_from_data draws from the torch RNG
    for lc, per in _ragged_lcs():
        res = lc.fit(model="1D", num_mixtures=2, periods=[per, 2.1 * per], training_iter=4,
                     optim="AdamW", lr=0.05)
        single.append((np.array(res["loss"], dtype=float), pack_model(lc.model).raw().detach()))
    lcs = _ragged_lcs()
    torch.manual_seed(7)
    out = fit_batch([lc for lc, _ in lcs], model="1D", num_mixtures=2,
                    periods=[[per, 2.1 * per] for _, per in lcs], training_iter=4, optim="AdamW",
                    lr=0.05, device="cpu")
    assert out["loss"].shape == (4, 3) and out["raw"].shape[0] == 3
    assert out["n_iter"].tolist() == [4, 4, 4] and out["info"].tolist() == [0, 0, 0]
    for b, (lc, per) in enumerate(lcs):
        # the single-source history is in the model's dtype (fp32 here), the batch one is fp64
        assert np.allclose(out["loss"][:, b].numpy(), single[b][0], rtol=1e-6, atol=1e-7)
        assert np.allclose(out["raw"][b].numpy(), single[b][1].numpy(), rtol=1e-6, atol=1e-7)
        assert np.allclose(np.array(lc.results["loss"], dtype=float), single[b][0], rtol=1e-10)
        assert len(lc.results["covar_module.mixture_means"]) == 5
        assert abs(out["periods"][b][0] - per) < 0.2 * per
    with pytest.raises(ValueError):
        fit_batch([], model="1D")


def _gloo_fit_worker(rank, world, port, q):
    import torch.distributed as dist
    from pgmuvi_b200.batch import fit_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    trainers.ops.sm_fit = fake_sm_fit
    lcs = _ragged_lcs()
    torch.manual_seed(7)
    out = fit_batch([lc for lc, _ in lcs], model="1D", num_mixtures=2,
                    periods=[[per, 2.1 * per] for _, per in lcs], training_iter=3, optim="Adam",
                    lr=0.05, device="cpu")
    has_results = [hasattr(lc, "results") for lc, _ in lcs]
    q.put((rank, out["raw"].numpy(), out["loss"].numpy(), has_results))
    dist.destroy_process_group()


def test_fit_batch_shards_over_ranks_and_gathers_gloo():
    """world_size 2: rank 0 fits light curves 0-1, rank 1 fits light curve 2; both end with all
    fitted parameters (one all-gather at the end), equal to the single-process result."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_fit_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted((q.get(timeout=300) for _ in procs), key=lambda o: o[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    assert outs[0][3] == [True, True, False] and outs[1][3] == [False, False, True]
    assert outs[0][1].shape[0] == 3 and np.isfinite(outs[0][2]).all()


def test_packed_noise_index_for_both_layouts():
    """the learned-noise Parameter sits at params[4] only in the spectral-mixture
    layout [mean, w, mu, sigma, noise, ...]; the stationary kinds pack [mean, noise, ...]."""
    lc, *_ = _lc_1d(n=30, yerr=False)                 # GaussianLikelihood: learned noise
    lc.set_model("1D", num_mixtures=2)
    pk = pack_model(lc.model)
    assert pk.learn_noise and pk.noise_index == 4
    assert pk.params[pk.noise_index] is lc.likelihood.noise_covar.raw_noise
    for name in ("1DMatern", "1DQuasiPeriodic"):
        lc.set_model(name)
        pk = pack_model(lc.model)
        assert pk.learn_noise and pk.noise_index == 1
        assert pk.params[pk.noise_index] is lc.likelihood.noise_covar.raw_noise
    lc2, *_ = _lc_1d(n=30, yerr=True)                  # plain FixedNoise: nothing learned
    lc2.set_model("1DMatern")
    assert pack_model(lc2.model).noise_index is None


def test_fit_batch_rejects_non_constant_means():
    """mean-function parameters live on the host; the one-launch batch loop would
    silently train a free constant instead, so fit_batch refuses such models."""
    from pgmuvi_b200.batch import fit_batch
    lcs = [_lc_1d(n=24, seed=s)[0] for s in (1, 2)]
    with pytest.raises(UnsupportedModelError):
        fit_batch(lcs, model="1DLinear", num_mixtures=2, training_iter=2)


def test_torch_optimizer_history_keys_follow_the_reference(cpu_engine, monkeypatch):
    """parameters without 'raw' in their name keep it (lightcurve.py:9031-9077):
    'mean_module.weights', not 'mean_module.eights'."""
    keys = []
    lc, *_ = _lc_1d(n=24)
    lc.set_model("1DLinear", num_mixtures=2)
    names = [n for n, _ in lc.model.named_parameters() if n.startswith("mean_module")]
    assert names
    snap = {}
    for n in names:
        key = trainers._strip_raw(n) if "raw" in n else n
        snap[n] = key
    assert all(not k.startswith("mean_module.eights") for k in snap.values())
    assert set(snap.values()) <= set(lc.get_parameters())


def test_flicker_term_packs_as_an_extra_spectral_mixture_component():
    """N3: WavelengthDependentGPModel(time_kernel_type='sm', add_flicker=True) has the time kernel
    SMK(Q) + ScaleKernel(RBFKernel) (gps.py:992-1002).  The RBF term IS an SM component (weight =
    outputscale, mean frequency 0, frequency scale 1 / (2 pi lengthscale)): pack_model emits Q + 1
    mixtures with a frozen zero and a PGM_CON_RSOFTPLUS slot.  Check the packed model's dense
    covariance (oracle) against the explicit sum written with the modules' own parameters."""
    import math
    import warnings
    from oracle import ModelSpec, constrain, kernel_dense
    from pgmuvi_b200._lib import CON_INTERVAL, CON_RSOFTPLUS, KIND_SEP_RBF
    torch.manual_seed(3)
    n = 40
    x = torch.stack([torch.sort(torch.rand(n) * 50).values, torch.randint(0, 3, (n,)) + 0.5], 1).float()
    lik = gp.FixedNoiseGaussianLikelihood(noise=torch.full((n,), 0.01))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = gp.WavelengthDependentGPModel(x, torch.randn(n), lik, time_kernel_type="sm", num_mixtures=3,
                                          add_flicker=True, mean_module="constant")
    assert any("work-in-progress" in str(v.message) for v in w)
    add = m.covar_module.kernels[0]
    smk, fl, wl = add.kernels[0], add.kernels[1], m.covar_module.kernels[1]
    with torch.no_grad():
        smk.raw_mixture_weights.copy_(torch.tensor([0.3, -0.2, 0.1]))
        smk.raw_mixture_means.copy_(torch.tensor([-2.0, -1.0, -3.0]).reshape(3, 1, 1))
        smk.raw_mixture_scales.copy_(torch.tensor([-3.0, -2.5, -4.0]).reshape(3, 1, 1))
        fl.raw_outputscale.fill_(0.4)
        fl.base_kernel.raw_lengthscale.fill_(1.7)
        wl.raw_outputscale.fill_(0.2)
    pk = pack_model(m, lik)
    assert (pk.kind, pk.Q, pk.d, pk.P) == (KIND_SEP_RBF, 4, 2, 15)
    assert pk.kinds.tolist().count(CON_RSOFTPLUS) == 1 and pk.kinds[8] == CON_INTERVAL
    assert pk.names[4] == "?" and float(pk.lb[8]) == float(pk.ub[8]) == 0.0
    raw = pk.raw().detach().double()
    theta = constrain(raw, pk.kinds, pk.lb, pk.ub)
    xd = x.double()
    K = kernel_dense(xd, xd, theta, ModelSpec(d=2, Q=4, kind=KIND_SEP_RBF, learn_noise=False))
    sp = torch.nn.functional.softplus
    tau = xd[:, :1] - xd[:, :1].T
    wq, mu, sg = (sp(p.detach().double().reshape(-1)) for p in
                  (smk.raw_mixture_weights, smk.raw_mixture_means, smk.raw_mixture_scales))
    Kt = sum(wq[q] * torch.exp(-2 * math.pi ** 2 * sg[q] ** 2 * tau ** 2) * torch.cos(2 * math.pi * mu[q] * tau)
             for q in range(3))
    osf, ell = sp(fl.raw_outputscale.detach().double()), sp(fl.base_kernel.raw_lengthscale.detach().double())
    Kt = Kt + osf * torch.exp(-0.5 * tau ** 2 / ell ** 2)
    dl = xd[:, 1:] - xd[:, 1:].T
    Kl = sp(wl.raw_outputscale.detach().double()) * torch.exp(
        -0.5 * dl ** 2 / sp(wl.base_kernel.raw_lengthscale.detach().double()) ** 2)
    assert torch.allclose(K, Kt * Kl, rtol=1e-12, atol=1e-14)
    # scatter leaves the frozen slot alone and round-trips the real parameters
    pk.scatter_raw_(raw + 0.25)
    assert float(fl.base_kernel.raw_lengthscale) == pytest.approx(1.95)
