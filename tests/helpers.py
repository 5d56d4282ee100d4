"""Glue between the synthetic workloads and the CPU oracle (test infrastructure only)."""
from __future__ import annotations

import numpy as np

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
