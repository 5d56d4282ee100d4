"""``GaussianProcessSurrogate`` with the reference's public surface
(``/root/reference/baybe/surrogates/base.py:49-78,213-247,308-465`` and
``gaussian_process/core.py:125-341``): ``fit(searchspace, objective, measurements)`` is cached on
its context, ``posterior`` / ``posterior_stats`` take candidates in experimental representation,
and the model owns input Normalize / output Standardize (core.py:130-141 "Scaling Workaround").

The fitted model lives on the GPU as a ``DeviceGP`` (caches built by ``bb_model_build``); every
posterior evaluation runs the tcgen05 kernel.  Hyper-parameter fitting (*before*
the hot path) evaluates the exact marginal likelihood and its gradient on the GPU (``bb_fit_eval``, float64)
under scipy's L-BFGS-B on the BayBE preset's MAP objective
(``presets/baybe.py:57-144``: Matern-5/2 ARD, Gamma(3, rate(d)) lengthscale prior with lower bound
2.5e-2, Gamma(2, e^5) noise prior with floor 1e-4, constant mean, no output scale).
"""

from __future__ import annotations

import math
from typing import ClassVar, Sequence

import numpy as np
import pandas as pd
import torch
from attrs import define, field

from baybe_b200.engine import DeviceGP
from baybe_b200.searchspace import objective_affine

__all__ = ["GaussianProcessSurrogate", "ModelNotTrainedError", "fit_map", "fit_map_hyperparameters",
           "fit_map_hyperparameters_device", "DeviceMLL", "HostMLL"]

MIN_INFERRED_NOISE_LEVEL = 1e-4
MIN_LENGTHSCALE = 2.5e-2


class ModelNotTrainedError(Exception):
    """Same name/meaning as baybe.exceptions.ModelNotTrainedError (surrogates/base.py:240-243)."""


def _torch_kernel(family: str, d2: torch.Tensor) -> torch.Tensor:
    if family == "rbf":
        return torch.exp(-0.5 * d2)
    r = d2.clamp_min(1e-30).sqrt()
    if family == "matern12":
        return torch.exp(-r)
    if family == "matern32":
        s = math.sqrt(3.0) * r
        return (1.0 + s) * torch.exp(-s)
    s = math.sqrt(5.0) * r
    return (1.0 + s + (5.0 / 3.0) * d2.clamp_min(0.0)) * torch.exp(-s)


class HostMLL:
    """Float64 torch-autograd twin of ``DeviceMLL`` (same call signature); the independent cross-check of
    the device objective and the ``fit_backend="host"`` path."""

    def __init__(self, Xa: np.ndarray, y_std: np.ndarray, task_ids=None, n_tasks: int = 1,
                 family: str = "matern52", device=None):
        self.X = torch.as_tensor(np.ascontiguousarray(Xa), dtype=torch.float64)
        self.y = torch.as_tensor(np.ascontiguousarray(y_std), dtype=torch.float64)
        self.n, self.d = self.X.shape
        self.T = int(n_tasks)
        self.family = family
        self.tid = (torch.zeros(self.n, dtype=torch.long) if task_ids is None
                    else torch.as_tensor(np.asarray(task_ids), dtype=torch.long))
        self.np = self.d + 2 + self.T * self.T

    def __call__(self, theta: np.ndarray) -> tuple[float, np.ndarray, bool]:
        t = torch.tensor(np.asarray(theta, dtype=np.float64), requires_grad=True)
        d, n, T = self.d, self.n, self.T
        ls, nz, c, B = t[:d], t[d], t[d + 1], t[d + 2:].reshape(T, T)
        diff = (self.X[:, None, :] - self.X[None, :, :]) / ls  # direct differences, like the device kernels
        d2 = (diff * diff).sum(-1)
        eye = torch.eye(n, dtype=torch.float64)
        K = _torch_kernel(self.family, d2) * (1.0 - eye) + eye  # exact unit diagonal (x1 is x2)
        K = K * B[self.tid][:, self.tid] + nz * eye
        L, info = torch.linalg.cholesky_ex(K)
        if int(info) != 0:
            return float("nan"), np.zeros(self.np), False
        r = (self.y - c).unsqueeze(-1)
        alpha = torch.cholesky_solve(r, L)
        mll = -0.5 * (r * alpha).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * math.log(2 * math.pi)
        mll.backward()
        return float(mll.detach()), t.grad.numpy().copy(), True


class DeviceMLL:
    """Exact marginal log likelihood and its gradient on the GPU (``bb_fit_setup`` / ``bb_fit_eval``,
    float64): theta = [lengthscale[d] | noise | mean constant | B[T*T]]."""

    def __init__(self, Xa: np.ndarray, y_std: np.ndarray, task_ids=None, n_tasks: int = 1,
                 family: str = "matern52", device=None):
        import ctypes as C

        from baybe_b200 import _lib
        from baybe_b200.engine import _require_cuda, _stream_ptr

        self._C, self._lib_mod, self._stream_ptr = C, _lib, _stream_ptr
        self.device = _require_cuda(device)
        self.lib = _lib.load()
        self.n, self.d = Xa.shape
        self.T = int(n_tasks)
        self.family = _lib.KERNEL_FAMILY[family]
        self.np = self.d + 2 + self.T * self.T
        xa = np.ascontiguousarray(Xa, dtype=np.float64)
        yy = np.ascontiguousarray(y_std, dtype=np.float64)
        tt = None if task_ids is None else np.ascontiguousarray(task_ids, dtype=np.int32)
        nbytes = self.lib.bb_fit_workspace_bytes(self.n, self.d, self.T)
        with torch.cuda.device(self.device):
            self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            self._base = (self._ws.data_ptr() + 255) // 256 * 256
            dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
            _lib.check(self.lib.bb_fit_setup(
                C.c_void_p(self._base), nbytes, self.n, self.d, self.T, dp(xa), dp(yy),
                None if tt is None else tt.ctypes.data_as(C.POINTER(C.c_int32)), _stream_ptr()), "bb_fit_setup")

    def __call__(self, theta: np.ndarray) -> tuple[float, np.ndarray, bool]:
        """(mll, d mll / d theta, positive_definite)."""
        C = self._C
        th = np.ascontiguousarray(theta, dtype=np.float64)
        if th.shape != (self.np,):
            raise ValueError(f"theta must have {self.np} entries")
        val = C.c_double(0.0)
        grad = np.zeros(self.np)
        bad = C.c_int32(0)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        with torch.cuda.device(self.device):
            self._lib_mod.check(self.lib.bb_fit_eval(
                C.c_void_p(self._base), self.n, self.d, self.T, self.family, dp(th), C.byref(val), dp(grad),
                C.byref(bad), self._stream_ptr()), "bb_fit_eval")
        return float(val.value), grad, bad.value == 0


def fit_map(Xn: np.ndarray, y_std: np.ndarray, active: Sequence[int], task_ids=None, n_tasks: int = 1,
            max_iter: int = 200, config=None, backend: str = "device", device=None) -> dict:
    """MAP fit of (lengthscales, noise, constant mean[, output scale][, task covariance]) on normalised inputs
    and standardised targets: maximises (log marginal likelihood + log priors) / n like
    ``botorch.fit.fit_gpytorch_mll`` on an ``ExactMarginalLogLikelihood`` (core.py:340-341), starting from the
    preset's initial values (prior modes for the BayBE preset, presets/baybe.py:100-107,134-144).

    The marginal likelihood and its gradient come from ``DeviceMLL`` (GPU, default) or ``HostMLL`` (torch
    autograd); priors, bounds and scipy's L-BFGS-B step are shared host code."""
    from scipy.optimize import minimize

    from baybe_b200.kernels import gp_preset

    Xa = np.ascontiguousarray(np.asarray(Xn, dtype=np.float64)[:, list(active)])
    n, da = Xa.shape
    cfg = gp_preset("BAYBE", da) if config is None else config
    has_tasks = task_ids is not None
    T = n_tasks if has_tasks else 1
    mll = (DeviceMLL if backend == "device" else HostMLL)(Xa, y_std, task_ids, T, cfg.family, device)
    fit_os = cfg.outputscale and cfg.outputscale_trainable
    i_os = da + 2 if fit_os else None
    i_w = da + 2 + (1 if fit_os else 0)
    n_task_par = T * T + T if has_tasks else 0
    x0 = np.concatenate([
        np.full(da, max(cfg.lengthscale_initial_value, cfg.lengthscale_lower)),
        [max(cfg.noise_initial_value, cfg.noise_lower)], [0.0],
        [cfg.outputscale_initial_value] if fit_os else [],
        np.concatenate([np.eye(T).reshape(-1) * 0.8 + 0.2, np.full(T, 0.1)]) if n_task_par else []])
    bounds = [(cfg.lengthscale_lower, None)] * da + [(cfg.noise_lower, None), (None, None)] + \
             ([(1e-6, None)] if fit_os else []) + [(1e-6, None)] * n_task_par
    os_fixed = cfg.outputscale_initial_value if (cfg.outputscale and not fit_os) else 1.0

    def parts(x):
        osv = x[i_os] if fit_os else os_fixed
        if not n_task_par:
            return osv, np.ones((1, 1)), None
        W = x[i_w: i_w + T * T].reshape(T, T)
        v = x[i_w + T * T:]
        return osv, W @ W.T + np.diag(v), W

    def objective(x):
        osv, Bi, W = parts(x)
        theta = np.concatenate([x[: da + 2], (osv * Bi).reshape(-1)])
        val, g, ok = mll(theta)
        if not ok or not np.isfinite(val):
            return 1e10, np.zeros_like(x)
        ls, nz = x[:da], x[da]
        grad = np.zeros_like(x)
        grad[: da + 2] = g[: da + 2]
        lp = 0.0
        if cfg.lengthscale_prior is not None:
            lp += float(np.sum(cfg.lengthscale_prior.log_prob(ls)))
            grad[:da] += cfg.lengthscale_prior.grad(ls)
        if cfg.noise_prior is not None:
            lp += float(cfg.noise_prior.log_prob(nz))
            grad[da] += float(cfg.noise_prior.grad(nz))
        gB = g[da + 2:].reshape(T, T)
        if fit_os:
            grad[i_os] = float(np.sum(gB * Bi))
            if cfg.outputscale_prior is not None:
                lp += float(cfg.outputscale_prior.log_prob(osv))
                grad[i_os] += float(cfg.outputscale_prior.grad(osv))
        if n_task_par:
            grad[i_w: i_w + T * T] = (osv * (gB + gB.T) @ W).reshape(-1)
            grad[i_w + T * T:] = osv * np.diag(gB)
        return -(val + lp) / n, -grad / n

    res = minimize(objective, x0, jac=True, method="L-BFGS-B", bounds=bounds,
                   options={"maxiter": max_iter, "ftol": 1e-10, "gtol": 1e-7})
    osv, Bi, _ = parts(res.x)
    return {"lengthscale": res.x[:da].copy(), "noise": float(res.x[da]), "mean_const": float(res.x[da + 1]),
            "outputscale": float(osv) if cfg.outputscale else None,
            "task_covar": Bi if n_task_par else None, "family": cfg.family, "objective": float(res.fun),
            "n_iter": int(res.nit), "n_eval": int(res.nfev), "backend": backend}


def fit_map_hyperparameters(Xn, y_std, active, task_ids=None, n_tasks: int = 1, max_iter: int = 200,
                            config=None) -> dict:
    """``fit_map`` with the objective evaluated by float64 torch autograd on the host."""
    return fit_map(Xn, y_std, active, task_ids, n_tasks, max_iter, config, backend="host")


def fit_map_hyperparameters_device(Xn, y_std, active, task_ids=None, n_tasks: int = 1, max_iter: int = 200,
                                   device=None, config=None) -> dict:
    """``fit_map`` with the objective evaluated on the GPU (``bb_fit_eval``)."""
    return fit_map(Xn, y_std, active, task_ids, n_tasks, max_iter, config, backend="device", device=device)


class _Posterior:
    """Marginal (t-batch) posterior with the attribute names BayBE reads from BoTorch posteriors
    (``mean``, ``variance``, ``quantile``; surrogates/base.py:352-375)."""

    def __init__(self, mean: torch.Tensor, variance: torch.Tensor):
        self.mean = mean.reshape(-1, 1, 1)
        self.variance = variance.reshape(-1, 1, 1)

    def quantile(self, value: torch.Tensor) -> torch.Tensor:
        p = torch.as_tensor(value, dtype=torch.float64, device=self.mean.device)
        z = math.sqrt(2.0) * torch.erfinv(2.0 * p - 1.0)
        return self.mean + self.variance.sqrt() * z.to(self.mean.dtype)


@define
class GaussianProcessSurrogate:
    """GP surrogate whose posterior runs on the B200 engine."""

    supports_transfer_learning: ClassVar[bool] = True
    supports_multi_output: ClassVar[bool] = False

    hyperparameters: dict | None = field(default=None)
    """Optional fixed hyper-parameters {lengthscale (per active column), noise, mean_const,
    [outputscale], [task_covar], [family]}; when omitted they are MAP-fitted (BayBE preset)."""

    device: str | None = field(default=None)
    max_fit_iter: int = field(default=200)
    kernel_or_factory: object = field(default=None)
    """None (BayBE preset), a preset name ("BAYBE", "CHEN", "EDBO") or a ``baybe_b200.kernels`` kernel object
    (``GaussianProcessSurrogate(kernel_or_factory=...)``, surrogates/gaussian_process/core.py:147-186)."""

    fit_backend: str = field(default="device")
    """"device": marginal likelihood + gradient on the GPU (bb_fit_eval); "host": float64 torch autograd
    (kept as the independent cross-check of the device objective)."""

    device_gp: DeviceGP | None = field(init=False, default=None, eq=False, repr=False)
    fitted_hyperparameters: dict | None = field(init=False, default=None, eq=False, repr=False)
    _searchspace = field(init=False, default=None, eq=False, repr=False)
    _objective = field(init=False, default=None, eq=False, repr=False)
    _measurements_hash: int | None = field(init=False, default=None, eq=False, repr=False)
    _target_name: str | None = field(init=False, default=None, eq=False, repr=False)

    # ---- SurrogateProtocol -------------------------------------------------------------
    def fit(self, searchspace, objective, measurements: pd.DataFrame) -> None:
        """Train on the given context; repeated calls with an unchanged context are no-ops
        (surrogates/base.py:419-424)."""
        h = int(pd.util.hash_pandas_object(measurements, index=True).sum())
        if (self.device_gp is not None and searchspace is self._searchspace
                and objective == self._objective and h == self._measurements_hash):
            return
        if getattr(objective, "is_multi_output", False):
            raise NotImplementedError("multi-output objectives are outside the B200 engine's scope")
        _, _, target = objective_affine(objective)
        if measurements[target].isna().any():
            raise ValueError("partial measurements are not supported (handle_missing_values)")
        comp = searchspace.transform(measurements, allow_extra=True)
        train_x = comp.to_numpy(dtype=np.float64)
        train_y = measurements[target].to_numpy(dtype=np.float64)
        bounds = np.asarray(searchspace.scaling_bounds.to_numpy(copy=True), dtype=np.float64)
        d = train_x.shape[1]
        task_col = searchspace.task_idx
        n_tasks = searchspace.n_tasks if task_col is not None else 1
        active = [j for j in range(d) if j != task_col]
        hp = dict(self.hyperparameters) if self.hyperparameters is not None else None
        if hp is None:
            rng = np.where(np.abs(bounds[1] - bounds[0]) < 1e-12, 1.0, bounds[1] - bounds[0])
            Xn = (train_x - bounds[0]) / rng
            ys = train_y.std(ddof=1) if len(train_y) > 1 else 1.0
            ys = ys if ys >= 1e-8 else 1.0
            tids = None if task_col is None else np.rint(train_x[:, task_col]).astype(int)
            from baybe_b200.kernels import Kernel, gp_preset, resolve_kernel

            kf = self.kernel_or_factory
            if kf is None:
                config = gp_preset("BAYBE", len(active))
            elif isinstance(kf, str):
                config = gp_preset(kf, len(active))
            elif isinstance(kf, Kernel):
                config = resolve_kernel(kf)
            else:
                raise TypeError("kernel_or_factory must be None, a preset name or a baybe_b200.kernels.Kernel")
            hp = fit_map(Xn, (train_y - train_y.mean()) / ys, active, tids, n_tasks, self.max_fit_iter, config,
                         backend=self.fit_backend, device=self.device)
        ls_full = np.full(d, -1.0)
        ls_full[active] = np.broadcast_to(np.asarray(hp["lengthscale"], dtype=np.float64), (len(active),))
        task_covar = hp.get("task_covar")
        if task_col is not None and task_covar is None:
            task_covar = np.eye(n_tasks)
        if self.device_gp is not None:
            self.device_gp.close()
        self.device_gp = DeviceGP(
            train_x, train_y, np.stack([bounds[0], bounds[1]]), hp.get("family", "matern52"), ls_full,
            hp["noise"], hp.get("mean_const", 0.0), hp.get("outputscale"), task_col, task_covar,
            device=self.device,
        )
        self.fitted_hyperparameters = hp
        self._searchspace, self._objective = searchspace, objective
        self._measurements_hash, self._target_name = h, target

    def to_botorch(self):
        raise ImportError(
            "GaussianProcessSurrogate of baybe_b200 is not backed by a botorch.models.Model; use "
            "baybe_b200.recommenders.B200Recommender, which scores through the CUDA engine")

    # ---- posterior ---------------------------------------------------------------------
    def _require_fit(self):
        if self.device_gp is None or self._searchspace is None:
            raise ModelNotTrainedError("The surrogate must be trained before a posterior can be computed.")

    def posterior(self, candidates: pd.DataFrame, *, joint: bool = False) -> _Posterior:
        """Posterior at candidates given in experimental representation (surrogates/base.py:213-247).
        Only the marginal (``joint=False``, t-batch) form is on the fast path."""
        self._require_fit()
        if joint:
            raise NotImplementedError("joint q-batch posteriors are outside the B200 fast path")
        comp = self._searchspace.transform(candidates, allow_extra=True)
        return self._posterior_comp(torch.from_numpy(comp.to_numpy(dtype=np.float64, copy=True)))

    def _posterior_comp(self, candidates_comp: torch.Tensor) -> _Posterior:
        """Posterior for un-scaled comp-rep rows (surrogates/base.py:249-272)."""
        self._require_fit()
        mu, var = self.device_gp.posterior(candidates_comp)
        return _Posterior(mu, var)

    def posterior_stats(self, candidates: pd.DataFrame, stats: Sequence = ("mean", "std")) -> pd.DataFrame:
        """Posterior statistics per candidate, columns ``{target}_{stat}`` (base.py:308-384)."""
        self._require_fit()
        for st in (x for x in stats if isinstance(x, float)):
            if not 0.0 < st < 1.0:
                raise ValueError(
                    f"Posterior quantile statistics can only be computed for quantiles between 0 and 1 "
                    f"(non-inclusive). Provided value: '{st}' as part of '{stats=}'.")
        post = self.posterior(candidates, joint=False)
        out = pd.DataFrame(index=candidates.index)
        for st in stats:
            if isinstance(st, float):
                name, vals = f"Q_{st}", post.quantile(torch.tensor(st))
            elif st == "mean":
                name, vals = st, post.mean
            elif st in ("std", "var"):
                name, vals = st, post.variance
                if st == "std":
                    vals = torch.sqrt(vals)
            else:
                raise TypeError(f"The utilized posterior does not support the statistic '{st}'.")
            out[f"{self._target_name}_{name}"] = vals.reshape(-1).double().cpu().numpy()
        return out
