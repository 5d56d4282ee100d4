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
