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

__all__ = ["GaussianProcessSurrogate", "ModelNotTrainedError", "fit_map_hyperparameters",
           "fit_map_hyperparameters_device", "DeviceMLL"]

MIN_INFERRED_NOISE_LEVEL = 1e-4
MIN_LENGTHSCALE = 2.5e-2


class ModelNotTrainedError(Exception):
    """Same name/meaning as baybe.exceptions.ModelNotTrainedError (surrogates/base.py:240-243)."""


def _matern52(Xa: torch.Tensor, Xb: torch.Tensor, ls: torch.Tensor) -> torch.Tensor:
    a, b = Xa / ls, Xb / ls
    d2 = (a * a).sum(-1, keepdim=True) + (b * b).sum(-1, keepdim=True).T - 2.0 * a @ b.T
    r = d2.clamp_min(1e-30).sqrt()
    s = math.sqrt(5.0) * r
    return (1.0 + s + (5.0 / 3.0) * d2.clamp_min(0.0)) * torch.exp(-s)


def fit_map_hyperparameters(Xn: np.ndarray, y_std: np.ndarray, active: Sequence[int], task_ids=None,
                            n_tasks: int = 1, max_iter: int = 200) -> dict:
    """MAP fit of (lengthscales, noise, constant mean[, task covariance]) on normalised inputs and
    standardised targets: maximises (log marginal likelihood + log priors) / n like
    ``botorch.fit.fit_gpytorch_mll`` on an ``ExactMarginalLogLikelihood`` (core.py:340-341),
    starting from the prior modes (``initial_value=prior.mode``, presets/baybe.py:100-107,134-144).
    """
    from scipy.optimize import minimize

    X = torch.as_tensor(Xn[:, list(active)], dtype=torch.float64)
    y = torch.as_tensor(y_std, dtype=torch.float64)
    n, da = X.shape
    conc_l, rate_l = 3.0, 2.0 / math.exp(math.sqrt(2.0) - 3.0) / math.sqrt(da)
    conc_n, rate_n = 2.0, 1.0 / math.exp(-5.0)
    ls0 = (conc_l - 1.0) / rate_l
    nz0 = (conc_n - 1.0) / rate_n
    T = n_tasks
    tid = None if task_ids is None else torch.as_tensor(task_ids, dtype=torch.long)
    n_task_par = 0 if tid is None else T * T + T
    x0 = np.concatenate([np.full(da, ls0), [max(nz0, MIN_INFERRED_NOISE_LEVEL)], [0.0],
                         np.concatenate([np.eye(T).reshape(-1) * 0.8 + 0.2, np.full(T, 0.1)]) if n_task_par else []])
    bounds = [(MIN_LENGTHSCALE, None)] * da + [(MIN_INFERRED_NOISE_LEVEL, None), (None, None)] + \
             [(1e-6, None)] * n_task_par

    def unpack(t):
        ls, nz, c = t[:da], t[da], t[da + 1]
        B = None
        if n_task_par:
            W = t[da + 2: da + 2 + T * T].reshape(T, T)
            v = t[da + 2 + T * T:]
            B = W @ W.T + torch.diag(v)
        return ls, nz, c, B

    def objective(theta_np):
        t = torch.tensor(theta_np, dtype=torch.float64, requires_grad=True)
        ls, nz, c, B = unpack(t)
        K = _matern52(X, X, ls)
        if B is not None:
            K = K * B[tid][:, tid]
        K = K + nz * torch.eye(n, dtype=torch.float64)
        L, info = torch.linalg.cholesky_ex(K)
        if int(info) != 0:
            return 1e10, np.zeros_like(theta_np)
        r = (y - c).unsqueeze(-1)
        a = torch.cholesky_solve(r, L)
        mll = -0.5 * (r * a).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * math.log(2 * math.pi)
        lp = ((conc_l - 1.0) * torch.log(ls) - rate_l * ls).sum() + (conc_n - 1.0) * torch.log(nz) - rate_n * nz
        loss = -(mll + lp) / n
        loss.backward()
        return float(loss.detach()), t.grad.numpy().copy()

    res = minimize(objective, x0, jac=True, method="L-BFGS-B", bounds=bounds,
                   options={"maxiter": max_iter, "ftol": 1e-10, "gtol": 1e-7})
    t = torch.tensor(res.x, dtype=torch.float64)
    ls, nz, c, B = unpack(t)
    return {"lengthscale": ls.numpy(), "noise": float(nz), "mean_const": float(c),
            "task_covar": None if B is None else B.numpy(), "objective": float(res.fun),
            "n_iter": int(res.nit)}


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


def fit_map_hyperparameters_device(Xn: np.ndarray, y_std: np.ndarray, active: Sequence[int], task_ids=None,
                                   n_tasks: int = 1, max_iter: int = 200, device=None) -> dict:
    """Same MAP objective, start point, bounds and optimiser as ``fit_map_hyperparameters``; the marginal
    likelihood and its gradient are evaluated on the GPU (Cholesky, K^-1 and the n^2 d gradient contraction
    in float64 kernels), the Gamma priors and scipy's L-BFGS-B step on the host."""
    from scipy.optimize import minimize

    Xa = np.ascontiguousarray(np.asarray(Xn, dtype=np.float64)[:, list(active)])
    n, da = Xa.shape
    T = n_tasks if task_ids is not None else 1
    mll = DeviceMLL(Xa, y_std, task_ids, T, "matern52", device)
    conc_l, rate_l = 3.0, 2.0 / math.exp(math.sqrt(2.0) - 3.0) / math.sqrt(da)
    conc_n, rate_n = 2.0, 1.0 / math.exp(-5.0)
    ls0 = (conc_l - 1.0) / rate_l
    nz0 = (conc_n - 1.0) / rate_n
    n_task_par = 0 if task_ids is None else T * T + T
    x0 = np.concatenate([np.full(da, ls0), [max(nz0, MIN_INFERRED_NOISE_LEVEL)], [0.0],
                         np.concatenate([np.eye(T).reshape(-1) * 0.8 + 0.2, np.full(T, 0.1)]) if n_task_par else []])
    bounds = [(MIN_LENGTHSCALE, None)] * da + [(MIN_INFERRED_NOISE_LEVEL, None), (None, None)] + \
             [(1e-6, None)] * n_task_par

    def task_cov(x):
        if not n_task_par:
            return np.ones((1, 1)), None, None
        W = x[da + 2: da + 2 + T * T].reshape(T, T)
        v = x[da + 2 + T * T:]
        return W @ W.T + np.diag(v), W, v

    def objective(x):
        B, W, _ = task_cov(x)
        theta = np.concatenate([x[: da + 2], B.reshape(-1)])
        val, g, ok = mll(theta)
        if not ok or not np.isfinite(val):
            return 1e10, np.zeros_like(x)
        ls, nz = x[:da], x[da]
        lp = ((conc_l - 1.0) * np.log(ls) - rate_l * ls).sum() + (conc_n - 1.0) * math.log(nz) - rate_n * nz
        grad = np.zeros_like(x)
        grad[:da] = g[:da] + (conc_l - 1.0) / ls - rate_l
        grad[da] = g[da] + (conc_n - 1.0) / nz - rate_n
        grad[da + 1] = g[da + 1]
        if n_task_par:
            gB = g[da + 2:].reshape(T, T)
            grad[da + 2: da + 2 + T * T] = ((gB + gB.T) @ W).reshape(-1)
            grad[da + 2 + T * T:] = np.diag(gB)
        return -(val + lp) / n, -grad / n

    res = minimize(objective, x0, jac=True, method="L-BFGS-B", bounds=bounds,
                   options={"maxiter": max_iter, "ftol": 1e-10, "gtol": 1e-7})
    B, _, _ = task_cov(res.x)
    return {"lengthscale": res.x[:da].copy(), "noise": float(res.x[da]), "mean_const": float(res.x[da + 1]),
            "task_covar": B if n_task_par else None, "objective": float(res.fun), "n_iter": int(res.nit),
            "n_eval": int(res.nfev), "backend": "device"}


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
            fit = fit_map_hyperparameters if self.fit_backend == "host" else fit_map_hyperparameters_device
            kw = {} if self.fit_backend == "host" else {"device": self.device}
            hp = fit(Xn, (train_y - train_y.mean()) / ys, active, tids, n_tasks, self.max_fit_iter, **kw)
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
