"""Glue between the synthetic workloads and the CPU oracle (test infrastructure only)."""
from __future__ import annotations

import math

import numpy as np
import torch

import oracle
from baybe_b200.synthetic import Workload


def oracle_model(w: Workload) -> oracle.GPModel:
    active = [j for j in range(w.candidates.shape[1]) if w.lengthscale[j] > 0 and j != w.task_col]
    spec = oracle.KernelSpec(
        family=w.family, lengthscale=w.lengthscale[active], active_dims=active,
        outputscale=w.outputscale, task_idx=w.task_col, task_covar=w.task_covar,
    )
    return oracle.build_model(spec, w.train_x, w.train_y, w.bounds, noise=w.noise,
                              mean_const=w.mean_const)


def oracle_acq(kind: str, **kw) -> oracle.AcqSpec:
    return oracle.AcqSpec(kind=kind, **kw)


# --------------------------------------------------------------------------------------
# Per-row score bounds (no outlier allowance)
# --------------------------------------------------------------------------------------
MU_RTOL = 5e-5      # stated posterior-mean tolerance: |d mu|  <= MU_RTOL * max(1, |mu|_inf)
VAR_RTOL = 2e-5     # stated posterior-variance tolerance: |d var| <= VAR_RTOL * prior variance (original units)


def posterior_tolerances(om: oracle.GPModel, mu_ref) -> tuple[float, float]:
    """(|d mu|, |d var|) the CUDA posterior is allowed against the float64 oracle (tests/test_gpu_parity.py)."""
    prior = float(om.spec.outputscale or 1.0)
    if om.spec.task_covar is not None:
        prior *= float(np.max(np.diag(om.spec.task_covar)))
    return MU_RTOL * max(1.0, float(abs(mu_ref).max())), VAR_RTOL * prior * om.y_std**2


def score_bounds(om: oracle.GPModel, acq: oracle.AcqSpec, X, z, base_atol: float = 2e-4, base_rtol: float = 1e-4,
                 safety: float = 1.5):
    """Oracle scores of every row of X and a HARD per-row bound on |cuda score - oracle score|:

        bound_i = base_atol + base_rtol*|ref_i| + safety * max_corner |score(mu_i +- dmu, var_i +- dvar) - ref_i|

    i.e. the float32 arithmetic of the acquisition itself plus the score change the *stated* posterior
    tolerance (dmu, dvar) can cause at that row -- log-scale acquisition values of far-tail rows are steep in
    (mu, sigma), which is why a flat tolerance needed an outlier allowance in round 1.  Returns (ref, bound)."""
    import torch

    mu, var = oracle.posterior(om, X)
    ref = oracle.reference_path.acq_from_moments(acq, mu, var, z)
    dmu, dvar = posterior_tolerances(om, mu)
    dev = torch.zeros_like(ref)
    for sm in (-1.0, 1.0):
        for sv in (-1.0, 1.0):
            v = (var + sv * dvar).clamp_min(oracle.reference_path.MIN_VARIANCE * om.y_std**2)
            c = oracle.reference_path.acq_from_moments(acq, mu + sm * dmu, v, z)
            dev = torch.maximum(dev, (c - ref).abs())
    return ref, base_atol + base_rtol * ref.abs() + safety * dev


# --------------------------------------------------------------------------------------
# Oracle-backed stand-in for the device engine (CPU tests of the host-side binding only)
# --------------------------------------------------------------------------------------
class OracleBackedGP:
    """Same constructor and methods as ``baybe_b200.engine.DeviceGP``, computed by the float64 oracle on the
    CPU.  It exists so that the HOST logic around the engine (the BayBE plugin, masks, index plumbing, greedy
    bookkeeping) can be exercised without a GPU; the numbers of the CUDA path are checked by the ``-m gpu``
    tests.  Never imported by the product."""

    def __init__(self, train_x, train_y, bounds, family, lengthscale, noise, mean_const=0.0, outputscale=None,
                 task_col=None, task_covar=None, device=None):
        import torch

        tx = np.asarray(train_x, dtype=np.float64)
        ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (tx.shape[1],))
        active = [j for j in range(tx.shape[1]) if ls[j] > 0 and j != task_col]
        spec = oracle.KernelSpec(family=family, lengthscale=ls[active], active_dims=active, outputscale=outputscale,
                                 task_idx=task_col, task_covar=task_covar)
        self.om = oracle.build_model(spec, tx, np.asarray(train_y, dtype=np.float64), np.asarray(bounds), noise=noise,
                                     mean_const=mean_const)
        self.train_x = tx
        self.n, self.d = tx.shape
        self.device = torch.device("cpu")

    @staticmethod
    def _spec(acq) -> "oracle.AcqSpec":
        return oracle.AcqSpec(kind=acq.kind, best_f=acq.best_f, beta=acq.beta, obj_scale=acq.obj_scale,
                              obj_shift=acq.obj_shift, maximize=acq.maximize)

    def close(self):
        pass

    def prepare(self, x):
        import torch

        return torch.as_tensor(np.asarray(x, dtype=np.float64))

    def posterior(self, x):
        mu, var = oracle.posterior(self.om, np.asarray(x, dtype=np.float64))
        return mu.float(), var.float()

    def best_f(self, acq) -> float:
        return oracle.best_f_from_training(self.om, self.train_x, self._spec(acq))

    def argmax(self, scores, keep, index_offset=0):
        import torch

        from baybe_b200.engine import pack_best

        s = scores.double().clone()
        if keep is not None:
            s[keep == 0] = float("nan")
        best = -(1 << 63)
        ok = ~torch.isnan(s)
        if bool(ok.any()):
            m = float(s[ok].max())
            i = int(torch.nonzero(ok & (s == m))[0])
            best = pack_best(float(np.float32(m)), i + int(index_offset))
        return torch.tensor([best], dtype=torch.int64)

    def score(self, acq, x, z, keep=None, index_offset=0, want_scores=True):
        vals = oracle.acq_values(self.om, self._spec(acq), np.asarray(x, dtype=np.float64),
                                 None if z is None else z.reshape(-1))
        scores = vals.float()
        return (scores if want_scores else None), self.argmax(scores, keep, index_offset)

    def score_joint(self, acq, x, pending, z):
        return oracle.acq_values_joint(self.om, self._spec(acq), np.asarray(x, dtype=np.float64),
                                       np.asarray(pending, dtype=np.float64), z).float()


# --------------------------------------------------------------------------------------
# Host twin of the device fit criterion (cross-check only)
# --------------------------------------------------------------------------------------
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
    """Float64 torch-autograd twin of ``baybe_b200.surrogates.DeviceMLL`` (same constructor and call signature):
    the independent cross-check of the device fit criteria ("mll": exact marginal log likelihood, "loo":
    leave-one-out pseudo-likelihood) and their gradients.  TEST INFRASTRUCTURE -- it used to live in the product
    package as ``fit_backend="host"``."""

    def __init__(self, Xa: np.ndarray, y_std: np.ndarray, task_ids=None, n_tasks: int = 1,
                 family: str = "matern52", device=None, criterion: str = "mll"):
        self.criterion = criterion
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
        if self.criterion == "loo":
            # gpytorch LeaveOneOutPseudoLikelihood: sigma_i^2 = 1/[K^-1]_ii, mu_i = y_i - alpha_i sigma_i^2
            kap = torch.cholesky_inverse(L).diagonal()
            a = alpha.squeeze(-1)
            mll = (0.5 * kap.log() - 0.5 * a * a / kap).sum() - 0.5 * n * math.log(2 * math.pi)
        else:
            mll = -0.5 * (r * alpha).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * math.log(2 * math.pi)
        mll.backward()
        return float(mll.detach()), t.grad.numpy().copy(), True


