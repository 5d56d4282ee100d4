"""Synthetic workloads of the BASELINE.json configurations (numpy only; shared by tests and
bench.py).  No dataset or checkpoint exists offline, so candidates, measurements and
hyper-parameters are generated from fixed seeds as SURVEY.md section 8(d) prescribes."""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from baybe_b200.bits import pack_bits  # noqa: F401  (re-exported for tests and bench)


@dataclass
class Workload:
    name: str
    candidates: np.ndarray  # (N, d) float64 comp-rep rows
    train_idx: np.ndarray  # (n,) rows of `candidates` that were "measured"
    train_x: np.ndarray  # (n, d)
    train_y: np.ndarray  # (n,)
    bounds: np.ndarray  # (2, d) scaling bounds
    family: str
    lengthscale: np.ndarray  # (d,) ; <= 0 marks inactive columns
    noise: np.ndarray  # (T,)
    mean_const: np.ndarray  # (T,)
    outputscale: float | None = None
    task_col: int | None = None
    task_covar: np.ndarray | None = None

    def gp_kwargs(self) -> dict:
        return dict(
            train_x=self.train_x, train_y=self.train_y, bounds=self.bounds, family=self.family,
            lengthscale=self.lengthscale, noise=self.noise, mean_const=self.mean_const,
            outputscale=self.outputscale, task_col=self.task_col, task_covar=self.task_covar,
        )


def smooth_target(x: np.ndarray, seed: int = 7) -> np.ndarray:
    """Hartmann-like smooth function on [0,1]^d: sum of four anisotropic Gaussian bumps."""
    d = x.shape[1]
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0.15, 0.85, size=(4, d))
    widths = rng.uniform(0.5, 3.0, size=(4, d)) / d * 6.0
    amps = np.array([1.0, 1.2, 3.0, 3.2])
    y = np.zeros(x.shape[0])
    for a, c, w in zip(amps, centres, widths):
        y += a * np.exp(-((x - c) ** 2 * w).sum(axis=1))
    return y


def prior_mode_lengthscale(d_active: int) -> float:
    """Mode of the BayBE-preset Gamma(3, rate(d)) lengthscale prior
    (/root/reference/baybe/surrogates/gaussian_process/presets/baybe.py:95-99)."""
    return math.exp(math.sqrt(2.0) - 3.0) * math.sqrt(d_active)


PRIOR_MODE_NOISE = math.exp(-5.0)  # mode of Gamma(2, rate=e^5) (presets/baybe.py:134-144)


def numeric_grid_workload(N: int, d: int = 20, n: int = 256, levels: int = 11, seed: int = 0,
                          family: str = "matern52", noise: float | None = None,
                          lengthscale: float | np.ndarray | None = None,
                          outputscale: float | None = None, name: str | None = None) -> Workload:
    """BASELINE config 2 generator: d NumericalDiscreteParameters with `levels` values in [0,1],
    N rows sampled uniformly from the product grid, n training rows subsampled from them,
    prior-mode hyper-parameters."""
    rng = np.random.default_rng(seed)
    grid = np.linspace(0.0, 1.0, levels)
    cand = grid[rng.integers(0, levels, size=(N, d))]
    n = min(n, N)
    idx = rng.choice(N, size=n, replace=False)
    tx = cand[idx]
    ty = smooth_target(tx) + 0.01 * rng.standard_normal(n)
    ls = prior_mode_lengthscale(d) if lengthscale is None else lengthscale
    return Workload(
        name=name or f"grid{levels}^{d}_N{N}_n{n}_{family}",
        candidates=cand, train_idx=idx, train_x=tx, train_y=ty,
        bounds=np.stack([np.zeros(d), np.ones(d)]), family=family,
        lengthscale=np.broadcast_to(np.asarray(ls, dtype=np.float64), (d,)).copy(),
        noise=np.array([PRIOR_MODE_NOISE if noise is None else noise]),
        mean_const=np.array([0.0]), outputscale=outputscale,
    )


def task_workload(N_per_task: int, n_tasks: int = 4, d_num: int = 20, n_per_task: int = 128,
                  seed: int = 0, family: str = "matern52") -> Workload:
    """BASELINE config 5 generator: the config-2 grid replicated over `n_tasks` tasks with an
    integer task column appended (TaskParameter INT encoding,
    /root/reference/baybe/parameters/categorical.py:87-91) and a fixed positive ICM matrix
    B = W W^T + diag(v)."""
    rng = np.random.default_rng(seed)
    grid = np.linspace(0.0, 1.0, 11)
    blocks, tidx, txs, tys = [], [], [], []
    W = rng.uniform(0.3, 1.0, size=(n_tasks, n_tasks))
    v = rng.uniform(0.05, 0.3, size=n_tasks)
    B = W @ W.T + np.diag(v)
    B = B / B.max()
    off = 0
    for t in range(n_tasks):
        xb = grid[rng.integers(0, 11, size=(N_per_task, d_num))]
        blocks.append(np.concatenate([xb, np.full((N_per_task, 1), float(t))], axis=1))
        sel = rng.choice(N_per_task, size=min(n_per_task, N_per_task), replace=False)
        tidx.append(off + sel)
        txs.append(blocks[-1][sel])
        tys.append((1.0 + 0.15 * t) * smooth_target(xb[sel]) + 0.1 * t + 0.01 * rng.standard_normal(len(sel)))
        off += N_per_task
    cand = np.concatenate(blocks)
    d = d_num + 1
    ls = np.full(d, prior_mode_lengthscale(d_num))
    ls[-1] = -1.0
    lo = np.zeros(d)
    hi = np.ones(d)
    hi[-1] = n_tasks - 1
    return Workload(
        name=f"task{n_tasks}x{N_per_task}_n{n_tasks * n_per_task}", candidates=cand,
        train_idx=np.concatenate(tidx), train_x=np.concatenate(txs), train_y=np.concatenate(tys),
        bounds=np.stack([lo, hi]), family=family, lengthscale=ls,
        noise=np.full(n_tasks, PRIOR_MODE_NOISE), mean_const=np.zeros(n_tasks),
        task_col=d - 1, task_covar=B,
    )


def fingerprint_workload(N: int, d: int = 2048, n: int = 512, density: float = 0.05, seed: int = 0,
                         family: str = "rbf", ls_factor: float = 0.4, outputscale: float | None = 1.0,
                         noise: float | None = None) -> Workload:
    """BASELINE config 4 generator: binary substance fingerprints X ~ Bernoulli(density)
    of width d, n training rows subsampled from them, ScaleKernel(RBF) with ARD lengthscale
    sqrt(d) * ls_factor on every bit.  The target is a smooth function of 16 random bit-group counts.
    ``candidates`` holds the unpacked 0/1 matrix (float64); ``pack_bits`` gives the device layout."""
    rng = np.random.default_rng(seed)
    cand = (rng.random((N, d)) < density).astype(np.float64)
    n = min(n, N)
    idx = rng.choice(N, size=n, replace=False)
    tx = cand[idx]
    groups = rng.integers(0, 16, size=d)
    feats = np.stack([tx[:, groups == g].sum(axis=1) for g in range(16)], axis=1)
    feats = feats / max(1.0, density * d / 16.0) / 2.0  # ~0.5 on average
    ty = smooth_target(np.clip(feats, 0.0, 1.0)) + 0.01 * rng.standard_normal(n)
    return Workload(
        name=f"fingerprint{d}_N{N}_n{n}_{family}",
        candidates=cand, train_idx=idx, train_x=tx, train_y=ty,
        bounds=np.stack([np.zeros(d), np.ones(d)]), family=family,
        lengthscale=np.full(d, math.sqrt(d) * ls_factor),
        noise=np.array([PRIOR_MODE_NOISE if noise is None else noise]),
        mean_const=np.array([0.0]), outputscale=outputscale,
    )


def mixed_small_workload(seed: int = 0) -> Workload:
    """BASELINE config 1 shape: 3 parameters (one-hot categorical with 3 levels, two numerical
    with 8 levels) -> ~192 candidates in 5 comp-rep columns, 15 training points."""
    rng = np.random.default_rng(seed)
    cats = np.eye(3)
    a = np.array([1.0, 2.0, 5.0, 10.0, 20.0, 50.0, 80.0, 100.0])
    b = np.linspace(90.0, 160.0, 8)
    rows = [np.concatenate([c, [x, y]]) for c in cats for x in a for y in b]
    cand = np.array(rows)
    idx = rng.choice(len(cand), size=15, replace=False)
    tx = cand[idx]
    ty = 50 + 10 * tx[:, 0] - 5 * tx[:, 2] + 0.3 * tx[:, 3] - 0.002 * (tx[:, 4] - 120.0) ** 2 + rng.standard_normal(15)
    lo = np.array([0, 0, 0, 1.0, 90.0])
    hi = np.array([1, 1, 1, 100.0, 160.0])
    return Workload(
        name="cfg1_3param_192cand_n15", candidates=cand, train_idx=idx, train_x=tx, train_y=ty,
        bounds=np.stack([lo, hi]), family="matern52",
        lengthscale=np.full(5, prior_mode_lengthscale(5)), noise=np.array([PRIOR_MODE_NOISE]),
        mean_const=np.array([0.1]),
    )
