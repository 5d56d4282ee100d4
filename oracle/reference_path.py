"""CPU restatement (torch float64) of the reference's recommend-time scoring path.

TEST INFRASTRUCTURE ONLY -- never imported by ``baybe_b200``.  PARITY UNPINNED (see
``oracle/__init__.py``): the arithmetic lives in botorch 0.16.1 / gpytorch 1.14.3 /
linear-operator 0.6, which are not vendored in ``/root/reference``; every function
cites the reference call site that reaches it and restates the library's published
algorithm.

Path restated:

    Campaign.recommend                      /root/reference/baybe/campaign.py:495
    -> BayesianRecommender.recommend        baybe/recommenders/pure/bayesian/base.py:129
    -> GaussianProcessSurrogate._fit        baybe/surrogates/gaussian_process/core.py:272-341
    -> BotorchAcquisitionFunctionBuilder    baybe/acquisition/_builder.py:195-265
    -> recommend_discrete_without_subsets   baybe/recommenders/pure/bayesian/botorch/discrete.py:78-142
       -> botorch.optim.optimize_acqf_discrete   (discrete.py:124-126)

All tensors are float64, the reference default (baybe/settings.py:228,318-322).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

DTYPE = torch.float64

# botorch.models.utils.gpytorch_modules.MIN_INFERRED_NOISE_LEVEL (presets/baybe.py:129-144)
MIN_INFERRED_NOISE_LEVEL = 1e-4
# gpytorch.settings.min_variance for float64 (used by MultivariateNormal.variance)
MIN_VARIANCE = 1e-10
# botorch.utils.safe_math TAU / qLogEI defaults (acqf built at _builder.py:206)
TAU_RELU = 1e-6
TAU_MAX = 1e-2
FATPLUS_ALPHA = 1e-1
FATMAX_ALPHA = 2.0
# botorch.optim.optimize_acqf_discrete default max_batch_size (discrete.py:124-126)
MAX_BATCH_SIZE = 2048
# default MC sample count of botorch MC acquisition functions (_builder.py:267-274 pokes
# the same ``_default_sample_shape`` attribute for Thompson sampling)
DEFAULT_MC_SAMPLES = 512

KERNEL_FAMILIES = ("matern12", "matern32", "matern52", "rbf")


# --------------------------------------------------------------------------------------
# Specs
# --------------------------------------------------------------------------------------
@dataclass
class KernelSpec:
    """Evaluated kernel hyper-parameters (what ``Kernel.to_gpytorch`` builds,
    baybe/kernels/base.py:113-194, after fitting).

    family        one of KERNEL_FAMILIES (Matern nu=1/2,3/2,5/2; RBF; baybe/kernels/basic.py:48,166)
    lengthscale   ARD lengthscales, one per active dim (kernels/base.py:235-240)
    active_dims   comp-rep column indices the base kernel acts on (kernels/base.py:223-232)
    outputscale   ScaleKernel s_f^2 or None (kernels/composite.py:21; default preset has none,
                  presets/baybe.py:57-107)
    task_idx      comp-rep column of the task parameter or None (searchspace/core.py:272-283)
    task_covar    evaluated PositiveIndexKernel matrix B = W W^T + diag(v), (T,T)
                  (components/kernel.py:298-337; kernels/basic.py:239-248)
    """

    family: str
    lengthscale: np.ndarray
    active_dims: list[int]
    outputscale: float | None = None
    task_idx: int | None = None
    task_covar: np.ndarray | None = None

    def __post_init__(self):
        if self.family not in KERNEL_FAMILIES:
            raise ValueError(f"unknown kernel family {self.family!r}")
        self.lengthscale = np.asarray(self.lengthscale, dtype=np.float64).reshape(-1)
        self.active_dims = [int(i) for i in self.active_dims]
        if len(self.lengthscale) != len(self.active_dims):
            raise ValueError("one lengthscale per active dim required (ARD)")
        if (self.task_idx is None) != (self.task_covar is None):
            raise ValueError("task_idx and task_covar must be given together")
        if self.task_covar is not None:
            self.task_covar = np.asarray(self.task_covar, dtype=np.float64)


@dataclass
class AcqSpec:
    """One acquisition function + its context (baybe/acquisition/acqfs.py; _builder.py:195-265).

    kind       abbreviation as in the reference: qLogEI qEI qUCB qSR qPI (Monte Carlo) or
               UCB EI LogEI PI PM PSTD (analytic)
    best_f     max_i o(mu(x_i)) over the training inputs (_builder.py:256-265)
    beta       UCB/qUCB trade-off (acqfs.py:270,288)
    obj_scale, obj_shift   the affine objective o = a*y + b (objectives/single.py:66-91;
               minimisation is a = -1, targets/numerical.py:615-621)
    maximize   PSTD sign (acqfs.py:168-177)
    """

    kind: str
    best_f: float = 0.0
    beta: float = 0.2
    obj_scale: float = 1.0
    obj_shift: float = 0.0
    maximize: bool = True
    tau_relu: float = TAU_RELU
    tau_max: float = TAU_MAX
    tau_pi: float = 1e-3

    MC_KINDS = ("qLogEI", "qEI", "qUCB", "qSR", "qPI")
    ANALYTIC_KINDS = ("UCB", "EI", "LogEI", "PI", "PM", "PSTD")

    @property
    def is_mc(self) -> bool:
        return self.kind in self.MC_KINDS

    def __post_init__(self):
        if self.kind not in self.MC_KINDS + self.ANALYTIC_KINDS:
            raise ValueError(f"unsupported acquisition function {self.kind!r}")


@dataclass
class GPModel:
    """Fitted-model state: everything ``SingleTaskGP`` caches after the first prediction."""

    spec: KernelSpec
    lo: torch.Tensor  # (d,) scaling bounds, lower (searchspace/core.py:247-251)
    rng: torch.Tensor  # (d,) hi-lo with degenerate ranges replaced by 1 and task col = 1
    num_idx: list[int]  # normalised columns = all but the task column (core.py:104-111)
    Xn: torch.Tensor  # (n,d) normalised training inputs
    y_mean: float
    y_std: float
    mean_const: torch.Tensor  # (T,) constant mean per task (T=1 without task parameter)
    noise: torch.Tensor  # (T,) homoskedastic noise per task
    L: torch.Tensor  # (n,n) chol(K + noise)
    alpha: torch.Tensor  # (n,)   K^-1 (y~ - c)
    R: torch.Tensor  # (n,n)   L^-T, so that K^-1 = R R^T (fast_pred_var cache)
    jitter: float = 0.0
    extra: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return self.Xn.shape[0]

    @property
    def d(self) -> int:
        return self.Xn.shape[1]


# --------------------------------------------------------------------------------------
# A.1 transforms
# --------------------------------------------------------------------------------------
def _normalise(X: torch.Tensor, lo: torch.Tensor, rng: torch.Tensor) -> torch.Tensor:
    """BoTorch ``Normalize(d, bounds=scaling_bounds, indices=non-task)``
    (gaussian_process/core.py:301-305).  ``rng`` already holds 1 for the task column and
    for degenerate ranges, ``lo`` holds 0 for the task column."""
    return (X - lo) / rng


def _task_ids(X: torch.Tensor, task_idx: int | None) -> torch.Tensor:
    if task_idx is None:
        return torch.zeros(X.shape[0], dtype=torch.long)
    return X[:, task_idx].round().long()


# --------------------------------------------------------------------------------------
# A.2 kernels
# --------------------------------------------------------------------------------------
def _sq_dist(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """gpytorch ``Distance._sq_dist``: ||a||^2 + ||b||^2 - 2 a.b after subtracting a
    common mean, clamped at 0."""
    adj = a.mean(dim=0, keepdim=True)
    a = a - adj
    b = b - adj
    an = (a * a).sum(-1, keepdim=True)
    bn = (b * b).sum(-1, keepdim=True)
    res = an + bn.T - 2.0 * (a @ b.T)
    return res.clamp_min_(0.0)


def kernel_matrix(spec: KernelSpec, X1n: torch.Tensor, X2n: torch.Tensor) -> torch.Tensor:
    """k(X1, X2) on *normalised* inputs: Matern/RBF ARD [x ScaleKernel] [x task kernel]
    (modules built at kernels/base.py:173-178, components/kernel.py:337)."""
    ls = torch.as_tensor(spec.lengthscale, dtype=DTYPE)
    a = X1n[:, spec.active_dims] / ls
    b = X2n[:, spec.active_dims] / ls
    r2 = _sq_dist(a, b)
    if spec.family == "rbf":
        k = torch.exp(-0.5 * r2)
    else:
        r = r2.clamp_min(1e-30).sqrt()
        if spec.family == "matern12":
            k = torch.exp(-r)
        elif spec.family == "matern32":
            s = math.sqrt(3.0) * r
            k = (1.0 + s) * torch.exp(-s)
        else:
            s = math.sqrt(5.0) * r
            k = (1.0 + s + (5.0 / 3.0) * r2) * torch.exp(-s)
    if spec.outputscale is not None:
        k = k * float(spec.outputscale)
    if spec.task_idx is not None:
        B = torch.as_tensor(spec.task_covar, dtype=DTYPE)
        t1 = _task_ids(X1n, spec.task_idx)
        t2 = _task_ids(X2n, spec.task_idx)
        k = k * B[t1][:, t2]
    return k


def _kernel_diag(spec: KernelSpec, Xn: torch.Tensor) -> torch.Tensor:
    k = torch.ones(Xn.shape[0], dtype=DTYPE)
    if spec.outputscale is not None:
        k = k * float(spec.outputscale)
    if spec.task_idx is not None:
        B = torch.as_tensor(spec.task_covar, dtype=DTYPE)
        t = _task_ids(Xn, spec.task_idx)
        k = k * B[t, t]
    return k


# --------------------------------------------------------------------------------------
# A.3 training-side caches
# --------------------------------------------------------------------------------------
def _psd_safe_cholesky(K: torch.Tensor) -> tuple[torch.Tensor, float]:
    """linear_operator ``psd_safe_cholesky``: plain attempt, then jitter 1e-8 * 10^i."""
    L, info = torch.linalg.cholesky_ex(K)
    if info.item() == 0:
        return L, 0.0
    jitter_prev = 0.0
    Kp = K.clone()
    for i in range(3):
        jitter_new = 1e-8 * (10**i)
        Kp.diagonal().add_(jitter_new - jitter_prev)
        jitter_prev = jitter_new
        L, info = torch.linalg.cholesky_ex(Kp)
        if info.item() == 0:
            return L, jitter_new
    raise RuntimeError("matrix not positive definite after adding jitter up to 1e-6")


def build_model(
    spec: KernelSpec,
    train_x: np.ndarray | torch.Tensor,
    train_y: np.ndarray | torch.Tensor,
    bounds: np.ndarray | torch.Tensor,
    noise: float | np.ndarray,
    mean_const: float | np.ndarray = 0.0,
) -> GPModel:
    """Assemble what ``SingleTaskGP(train_x, train_y, Normalize, Standardize, mean, kernel,
    likelihood)`` (gaussian_process/core.py:331-339) holds once hyper-parameters are fixed.

    train_x  (n,d) raw comp-rep rows; train_y (n,) raw targets;
    bounds   (2,d) ``searchspace.scaling_bounds`` (lower row, upper row);
    noise    GaussianLikelihood noise (floored at MIN_INFERRED_NOISE_LEVEL), scalar or (T,);
    mean_const  ConstantMean value in *standardised* units, scalar or (T,).
    """
    X = torch.as_tensor(np.asarray(train_x), dtype=DTYPE)
    y = torch.as_tensor(np.asarray(train_y), dtype=DTYPE).reshape(-1)
    bounds = torch.as_tensor(np.asarray(bounds), dtype=DTYPE)
    n, d = X.shape
    num_idx = [i for i in range(d) if i != spec.task_idx]
    lo = bounds[0].clone()
    rng = (bounds[1] - bounds[0]).clone()
    # botorch Normalize(min_range=1e-8): a column whose range is (almost) zero is left unscaled  [U: botorch 0.16.1
    # input.py, Normalize.__init__ / _update_coefficients -- threshold restated from memory]
    rng = torch.where(rng.abs() < 1e-8, torch.ones_like(rng), rng)
    if spec.task_idx is not None:
        lo[spec.task_idx] = 0.0
        rng[spec.task_idx] = 1.0
    Xn = _normalise(X, lo, rng)

    # Standardize(m=1): unbiased std, floor 1e-8 -> 1
    y_mean = float(y.mean())
    y_std = float(y.std(unbiased=True)) if n > 1 else 1.0
    if not (y_std >= 1e-8):
        y_std = 1.0
    yt = (y - y_mean) / y_std

    T = 1 if spec.task_covar is None else spec.task_covar.shape[0]
    noise_t = torch.as_tensor(np.broadcast_to(np.asarray(noise, dtype=np.float64), (T,)).copy())
    noise_t = noise_t.clamp_min(MIN_INFERRED_NOISE_LEVEL)
    mean_t = torch.as_tensor(np.broadcast_to(np.asarray(mean_const, dtype=np.float64), (T,)).copy())
    tid = _task_ids(Xn, spec.task_idx)

    K = kernel_matrix(spec, Xn, Xn)
    K = K + torch.diag(noise_t[tid])
    L, jitter = _psd_safe_cholesky(K)
    resid = (yt - mean_t[tid]).unsqueeze(-1)
    alpha = torch.cholesky_solve(resid, L).squeeze(-1)
    eye = torch.eye(n, dtype=DTYPE)
    Linv = torch.linalg.solve_triangular(L, eye, upper=False)
    R = Linv.T.contiguous()
    return GPModel(
        spec=spec, lo=lo, rng=rng, num_idx=num_idx, Xn=Xn, y_mean=y_mean, y_std=y_std,
        mean_const=mean_t, noise=noise_t, L=L, alpha=alpha, R=R, jitter=jitter,
    )


# --------------------------------------------------------------------------------------
# A.4 posterior
# --------------------------------------------------------------------------------------
def _standardised_posterior(model: GPModel, Xn: torch.Tensor):
    Ks = kernel_matrix(model.spec, Xn, model.Xn)  # (B,n)
    tid = _task_ids(Xn, model.spec.task_idx)
    mean = model.mean_const[tid] + Ks @ model.alpha
    V = Ks @ model.R  # (B,n)
    return Ks, mean, V


def posterior(
    model: GPModel, X: np.ndarray | torch.Tensor, chunk: int | None = None
) -> tuple[torch.Tensor, torch.Tensor]:
    """Marginal posterior mean/variance at a t-batch of single points, original units
    (``SingleTaskGP.posterior`` with ``fast_pred_var``; entered via core.py:268-269).
    No observation noise.  ``chunk`` only bounds memory, the numbers do not depend on it."""
    X = torch.as_tensor(np.asarray(X), dtype=DTYPE)
    N = X.shape[0]
    mu = torch.empty(N, dtype=DTYPE)
    var = torch.empty(N, dtype=DTYPE)
    step = N if not chunk else chunk
    for s in range(0, N, max(step, 1)):
        Xn = _normalise(X[s : s + step], model.lo, model.rng)
        _, m, V = _standardised_posterior(model, Xn)
        v = _kernel_diag(model.spec, Xn) - (V * V).sum(-1)
        v = v.clamp_min(MIN_VARIANCE)
        mu[s : s + step] = model.y_mean + model.y_std * m
        var[s : s + step] = (model.y_std**2) * v
    return mu, var


def posterior_joint(model: GPModel, Xq: np.ndarray | torch.Tensor):
    """Joint posterior (mean (q,), covariance (q,q)) of one q-batch, original units."""
    Xq = torch.as_tensor(np.asarray(Xq), dtype=DTYPE)
    Xn = _normalise(Xq, model.lo, model.rng)
    _, m, V = _standardised_posterior(model, Xn)
    cov = kernel_matrix(model.spec, Xn, Xn) - V @ V.T
    # x1 is x2 -> exact prior diagonal (gpytorch zeroes the self-distance diagonal)
    cov.diagonal().copy_(_kernel_diag(model.spec, Xn) - (V * V).sum(-1))
    return model.y_mean + model.y_std * m, (model.y_std**2) * cov


def best_f_from_training(model: GPModel, train_x: np.ndarray | torch.Tensor, acq: AcqSpec) -> float:
    """best_f = max_i o(mu(x_i)) over the training inputs -- the posterior *mean*, not the
    observed targets (_builder.py:141-161,256-265)."""
    mu, _ = posterior(model, train_x)
    return float((acq.obj_scale * mu + acq.obj_shift).max())


# --------------------------------------------------------------------------------------
# A.6 sampler
# --------------------------------------------------------------------------------------
def sobol_normal_samples(n_samples: int, dim: int, seed: int) -> torch.Tensor:
    """botorch ``draw_sobol_normal_samples`` as used by ``SobolQMCNormalSampler``:
    scrambled Sobol -> v = 0.5 + (1-eps)(u-0.5) -> z = sqrt(2) erfinv(2v-1).  (S, dim)."""
    eng = torch.quasirandom.SobolEngine(dimension=dim, scramble=True, seed=seed)
    u = eng.draw(n_samples, dtype=DTYPE)
    v = 0.5 + (1.0 - torch.finfo(DTYPE).eps) * (u - 0.5)
    return torch.erfinv(2.0 * v - 1.0) * math.sqrt(2.0)


# --------------------------------------------------------------------------------------
# A.5 acquisition functions
# --------------------------------------------------------------------------------------
def _phi(u):
    return torch.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)


def _Phi(u):
    return 0.5 * torch.erfc(-u / math.sqrt(2.0))


def _log_phi(u):
    return -0.5 * (u * u + math.log(2.0 * math.pi))


def _log1mexp(x):
    # log(1 - exp(x)) for x < 0
    return torch.where(x > -math.log(2.0), torch.log(-torch.expm1(x)), torch.log1p(-torch.exp(x)))


def _log_ei_helper(u: torch.Tensor) -> torch.Tensor:
    """botorch ``_log_ei_helper``: log(phi(u) + u Phi(u)), stable in the left tail."""
    bound = -1.0
    u_upper = torch.where(u < bound, torch.full_like(u, bound), u)
    log_ei_upper = torch.log(_phi(u_upper) + u_upper * _Phi(u_upper))
    neg_inv_sqrt_eps = -1e6
    u_lower = torch.where(u > bound, torch.full_like(u, bound), u)
    u_eps = torch.where(u_lower < neg_inv_sqrt_eps, torch.full_like(u, neg_inv_sqrt_eps), u_lower)
    # log(|u| Phi(u)/phi(u)) = log(|u| erfcx(-u/sqrt2)) + log(sqrt(pi/2))
    w = torch.log(torch.special.erfcx(-u_eps / math.sqrt(2.0)) * u_eps.abs()) + 0.5 * math.log(math.pi / 2.0)
    tail = torch.where(u > neg_inv_sqrt_eps, _log1mexp(w), -2.0 * torch.log(u_lower.abs()))
    log_ei_lower = _log_phi(u) + tail
    return torch.where(u > bound, log_ei_upper, log_ei_lower)


def _log_fatplus(x: torch.Tensor, tau: float) -> torch.Tensor:
    """log of botorch ``fatplus(x, tau) = tau*(softplus(x/tau) + 0.1/(1+(x/tau)^2))``."""
    t = x / tau
    log_sp = torch.where(t > 0, torch.log(t + torch.log1p(torch.exp(-t.abs()))),
                         torch.log(torch.log1p(torch.exp(torch.minimum(t, torch.zeros_like(t))))))
    # for very negative t, log1p(exp(t)) underflows -> use log(softplus(t)) ~ t
    log_sp = torch.where(t < -30.0, t, log_sp)
    log_cauchy = math.log(FATPLUS_ALPHA) - torch.log1p(t * t)
    return math.log(tau) + torch.logaddexp(log_sp, log_cauchy)


def _fatmax(x: torch.Tensor, tau: float, dim: int = -1) -> torch.Tensor:
    """botorch ``fatmax``: M + tau*log sum_i (1 + (M-x_i)/(alpha tau))^-alpha, alpha=2."""
    M = x.amax(dim=dim, keepdim=True)
    par = (1.0 + (M - x) / (tau * FATMAX_ALPHA)).pow(-FATMAX_ALPHA)
    return (M + tau * par.sum(dim=dim, keepdim=True).log()).squeeze(dim)


def _mc_reduce(acq: AcqSpec, obj: torch.Tensor, mean_obj: torch.Tensor | None = None) -> torch.Tensor:
    """obj: (S, B, q) objective samples -> (B,) acquisition values."""
    S = obj.shape[0]
    k = acq.kind
    if k == "qLogEI":
        li = _log_fatplus(obj - acq.best_f, acq.tau_relu)  # (S,B,q)
        li = _fatmax(li, acq.tau_max, dim=-1) if obj.shape[-1] > 1 else li.squeeze(-1)
        return torch.logsumexp(li, dim=0) - math.log(S)
    if k == "qEI":
        return (obj - acq.best_f).clamp_min(0.0).amax(-1).mean(0)
    if k == "qSR":
        return obj.amax(-1).mean(0)
    if k == "qPI":
        return torch.sigmoid((obj - acq.best_f) / acq.tau_pi).amax(-1).mean(0)
    if k == "qUCB":
        # botorch qUpperConfidenceBound._sample_forward: ``mean = obj.mean(dim=0)`` -- the SAMPLE mean over
        # the MC dimension, not the posterior mean (they differ by so*mean(z) for scrambled Sobol samples)
        c = math.sqrt(acq.beta * math.pi / 2.0)
        m = obj.mean(dim=0, keepdim=True)
        return (m + c * (obj - m).abs()).amax(-1).mean(0)
    raise ValueError(k)


def _analytic(acq: AcqSpec, mu: torch.Tensor, var: torch.Tensor) -> torch.Tensor:
    """Analytic acquisition values with an affine posterior transform (_builder.py:224-236)."""
    a, b = acq.obj_scale, acq.obj_shift
    m = a * mu + b
    s = abs(a) * var.sqrt()
    k = acq.kind
    if k == "PM":
        return m
    if k == "PSTD":
        # botorch PosteriorStandardDeviation has no posterior_transform scaling issue for |a|=1
        return s if acq.maximize else -s
    if k == "UCB":
        return m + math.sqrt(acq.beta) * s
    u = (m - acq.best_f) / s
    if k == "EI":
        return s * (_phi(u) + u * _Phi(u))
    if k == "LogEI":
        return _log_ei_helper(u) + torch.log(s)
    if k == "PI":
        return _Phi(u)
    raise ValueError(k)


def acq_from_moments(acq: AcqSpec, mu: torch.Tensor, var: torch.Tensor, z: torch.Tensor | None = None) -> torch.Tensor:
    """q=1 acquisition values from marginal posterior moments (original units): the step after
    ``model.posterior`` inside the acquisition function's forward.  z: (S,) base samples for the MC kinds."""
    mu = torch.as_tensor(mu, dtype=DTYPE).reshape(-1)
    var = torch.as_tensor(var, dtype=DTYPE).reshape(-1)
    if not acq.is_mc:
        return _analytic(acq, mu, var)
    assert z is not None, "MC acquisition functions need base samples"
    zc = z.reshape(-1, 1).to(DTYPE)  # (S,1)
    y = mu.unsqueeze(0) + var.sqrt().unsqueeze(0) * zc  # (S,B)
    obj = (acq.obj_scale * y + acq.obj_shift).unsqueeze(-1)
    return _mc_reduce(acq, obj, None)


def acq_values(
    model: GPModel, acq: AcqSpec, X: np.ndarray | torch.Tensor, z: torch.Tensor | None = None,
    chunk: int | None = MAX_BATCH_SIZE,
) -> torch.Tensor:
    """Acquisition value of every row of X as its own q=1 batch (no pending points):
    ``acqf(X.unsqueeze(-2))`` evaluated in chunks like ``_split_batch_eval_acqf``.
    z: (S,) or (S,1) shared base samples for the MC kinds."""
    X = torch.as_tensor(np.asarray(X), dtype=DTYPE)
    N = X.shape[0]
    out = torch.empty(N, dtype=DTYPE)
    step = N if not chunk else chunk
    for s in range(0, N, max(step, 1)):
        mu, var = posterior(model, X[s : s + step])
        out[s : s + step] = acq_from_moments(acq, mu, var, z)
    return out


def _chol_with_jitter(cov: torch.Tensor) -> torch.Tensor:
    """Batched linear_operator ``psd_safe_cholesky`` of (B,r,r) covariances: jitter
    1e-8 * 10^i is added only to the matrices of the batch that failed."""
    L, info = torch.linalg.cholesky_ex(cov)
    if int(info.max()) == 0:
        return L
    jitter_prev = 0.0
    covp = cov.clone()
    for i in range(3):
        jitter_new = 1e-8 * (10**i)
        add = (info > 0).to(cov.dtype) * (jitter_new - jitter_prev)
        covp.diagonal(dim1=-2, dim2=-1).add_(add.unsqueeze(-1))
        jitter_prev = jitter_new
        L, info = torch.linalg.cholesky_ex(covp)
        if int(info.max()) == 0:
            return L
    raise RuntimeError("joint covariance not positive definite")


def acq_values_joint(
    model: GPModel, acq: AcqSpec, X: np.ndarray | torch.Tensor, X_pending: np.ndarray | torch.Tensor,
    z: torch.Tensor, chunk: int | None = MAX_BATCH_SIZE,
) -> torch.Tensor:
    """MC acquisition value of [x*; X_pending] for every row x* of X (the candidate comes
    first: botorch ``concatenate_pending_points`` does ``cat([X, X_pending], dim=-2)``).
    z: (S, 1+p) shared base samples; samples are ``mean + chol(cov) z``."""
    if not acq.is_mc:
        raise ValueError("pending points need a Monte Carlo acquisition function")
    X = torch.as_tensor(np.asarray(X), dtype=DTYPE)
    P = torch.as_tensor(np.asarray(X_pending), dtype=DTYPE).reshape(-1, X.shape[1])
    p = P.shape[0]
    if p == 0:
        return acq_values(model, acq, X, z[:, 0], chunk)
    N = X.shape[0]
    z = z.to(DTYPE)
    assert z.shape[1] == p + 1
    Pn = _normalise(P, model.lo, model.rng)
    _, mP, VP = _standardised_posterior(model, Pn)
    covPP = kernel_matrix(model.spec, Pn, Pn) - VP @ VP.T
    covPP.diagonal().copy_(_kernel_diag(model.spec, Pn) - (VP * VP).sum(-1))
    out = torch.empty(N, dtype=DTYPE)
    step = N if not chunk else chunk
    s2 = model.y_std**2
    for s in range(0, N, max(step, 1)):
        Xn = _normalise(X[s : s + step], model.lo, model.rng)
        B = Xn.shape[0]
        _, m, V = _standardised_posterior(model, Xn)
        var = (_kernel_diag(model.spec, Xn) - (V * V).sum(-1))
        cross = kernel_matrix(model.spec, Xn, Pn) - V @ VP.T  # (B,p)
        cov = torch.empty(B, p + 1, p + 1, dtype=DTYPE)
        cov[:, 0, 0] = var
        cov[:, 0, 1:] = cross
        cov[:, 1:, 0] = cross
        cov[:, 1:, 1:] = covPP
        cov = cov * s2
        mean = torch.cat([m.unsqueeze(-1), mP.unsqueeze(0).expand(B, p)], dim=-1)
        mean = model.y_mean + model.y_std * mean  # (B,1+p)
        Lc = _chol_with_jitter(cov)  # (B,1+p,1+p)
        y = mean.unsqueeze(0) + torch.einsum("bij,sj->sbi", Lc, z)  # (S,B,1+p)
        obj = acq.obj_scale * y + acq.obj_shift
        out[s : s + step] = _mc_reduce(acq, obj, None)
    return out


def acq_values_qnei(
    model: GPModel, acq: AcqSpec, X: np.ndarray | torch.Tensor, X_pending: np.ndarray | torch.Tensor | None,
    z: torch.Tensor, X_baseline: np.ndarray | torch.Tensor | None = None,
) -> torch.Tensor:
    """qNoisyExpectedImprovement of [x*; X_pending] for every row x* of X (baybe/acquisition/acqfs.py:227-232;
    ``X_baseline`` = the training inputs, acquisition/_builder.py:319-324; no baseline pruning):
        mean_s relu( max_{x*, pending} o(f_s) - max_{baseline} o(f_s) )
    over joint posterior samples f = mean + chol(cov) z  [U: botorch qNoisyExpectedImprovement._sample_forward].
    The joint Cholesky is taken in the order [baseline; pending; x*] -- z: (S, nb + p + 1), last column = the new
    point.  botorch orders the q-batch first, i.e. it pairs the same base samples with other rows: the two agree in
    distribution, single values differ by Monte-Carlo error; THIS order is the one the device path reproduces."""
    X = torch.as_tensor(np.asarray(X), dtype=DTYPE)
    d = X.shape[1]
    Xb = (model.Xn * model.rng + model.lo) if X_baseline is None else torch.as_tensor(np.asarray(X_baseline), dtype=DTYPE)
    P = torch.empty(0, d, dtype=DTYPE) if X_pending is None else torch.as_tensor(np.asarray(X_pending), dtype=DTYPE).reshape(-1, d)
    C = torch.cat([Xb, P], dim=0)
    nb, p = Xb.shape[0], P.shape[0]
    m = nb + p
    z = z.to(DTYPE)
    assert z.shape[1] == m + 1
    Cn = _normalise(C, model.lo, model.rng)
    _, mC, VC = _standardised_posterior(model, Cn)
    covCC = kernel_matrix(model.spec, Cn, Cn) - VC @ VC.T
    covCC.diagonal().copy_(_kernel_diag(model.spec, Cn) - (VC * VC).sum(-1))
    covCC = covCC * model.y_std**2
    LC = _chol_with_jitter(covCC.unsqueeze(0))[0]
    muC = model.y_mean + model.y_std * mC
    FC = muC.unsqueeze(0) + z[:, :m] @ LC.T  # (S, m)
    oC = acq.obj_scale * FC + acq.obj_shift
    best = oC[:, :nb].amax(-1)
    pend_max = oC[:, nb:].amax(-1) if p > 0 else torch.full_like(best, -float("inf"))
    Xn = _normalise(X, model.lo, model.rng)
    _, mX, VX = _standardised_posterior(model, Xn)
    varX = (_kernel_diag(model.spec, Xn) - (VX * VX).sum(-1)) * model.y_std**2
    covXC = (kernel_matrix(model.spec, Xn, Cn) - VX @ VC.T) * model.y_std**2  # (N, m)
    R = torch.linalg.solve_triangular(LC, covXC.T, upper=False).T  # (N, m) = covXC L_C^-T
    cond_sd = (varX - (R * R).sum(-1)).clamp_min(0.0).sqrt()
    muX = model.y_mean + model.y_std * mX
    fx = muX.unsqueeze(0) + z[:, :m] @ R.T + z[:, m : m + 1] * cond_sd.unsqueeze(0)  # (S, N)
    ox = acq.obj_scale * fx + acq.obj_shift
    top = torch.maximum(ox, pend_max.unsqueeze(-1))
    return (top - best.unsqueeze(-1)).clamp_min(0.0).mean(0)


# --------------------------------------------------------------------------------------
# A.7 optimize_acqf_discrete
# --------------------------------------------------------------------------------------
def optimize_acqf_discrete(
    model: GPModel, acq: AcqSpec, choices: np.ndarray | torch.Tensor, q: int,
    sampler_seed: int = 0, n_samples: int = DEFAULT_MC_SAMPLES,
    X_pending: np.ndarray | torch.Tensor | None = None, chunk: int | None = MAX_BATCH_SIZE,
) -> tuple[list[int], list[float]]:
    """botorch ``optimize_acqf_discrete(acqf, q, choices, max_batch_size=2048, unique=True)``
    (call site discrete.py:124-126): sequential greedy over q rounds; each round scores all
    remaining choices jointly with the pending set, takes ``argmax`` (first maximum), appends
    the winner to X_pending and removes it from the choices.  Returns (row indices into
    ``choices`` in selection order, their acquisition values)."""
    X = torch.as_tensor(np.asarray(choices), dtype=DTYPE)
    N, d = X.shape
    if q > 1 and not acq.is_mc:
        raise ValueError("q>1 needs a Monte Carlo acquisition function (discrete.py:110-114)")
    base_pending = (
        torch.empty(0, d, dtype=DTYPE) if X_pending is None
        else torch.as_tensor(np.asarray(X_pending), dtype=DTYPE).reshape(-1, d)
    )
    alive = torch.ones(N, dtype=torch.bool)
    chosen: list[int] = []
    values: list[float] = []
    for _ in range(q):
        pend = torch.cat([base_pending, X[chosen]], dim=0)
        idx = torch.nonzero(alive).squeeze(-1)
        if acq.is_mc:
            z = sobol_normal_samples(n_samples, 1 + pend.shape[0], sampler_seed)
            vals = acq_values_joint(model, acq, X[idx], pend, z, chunk)
        else:
            vals = acq_values(model, acq, X[idx], None, chunk)
        j = int(torch.argmax(vals))
        chosen.append(int(idx[j]))
        values.append(float(vals[j]))
        alive[idx[j]] = False
    return chosen, values
