"""GPU tests of the device-side hyper-parameter fit: bb_fit_eval's marginal log
likelihood / bb_fit_eval_loo's leave-one-out pseudo-likelihood and their gradients against float64 torch
autograd on the CPU (tests/helpers.py::HostMLL), and the fitted hyper-parameters of the device-driven MAP fit
against the same fit driven by the host twin (same objective, same optimiser)."""
from __future__ import annotations

import math

import numpy as np
import pytest
import torch

from baybe_b200.surrogates import DeviceMLL, fit_map, fit_map_hyperparameters_device
from tests.helpers import HostMLL
from baybe_b200.synthetic import numeric_grid_workload, task_workload

pytestmark = pytest.mark.gpu


def _kernel(family, d2):
    if family == "rbf":
        return torch.exp(-0.5 * d2)
    r = d2.clamp_min(1e-30).sqrt()
    if family == "matern12":
        return torch.exp(-r)
    if family == "matern32":
        s = math.sqrt(3.0) * r
        return (1.0 + s) * torch.exp(-s)
    s = math.sqrt(5.0) * r
    return (1.0 + s + (5.0 / 3.0) * d2) * torch.exp(-s)


def _host_mll(X, y, tid, T, family, theta):
    """Independent float64 restatement: returns (mll, gradient) by autograd."""
    n, d = X.shape
    t = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    ls, nz, c, B = t[:d], t[d], t[d + 1], t[d + 2:].reshape(T, T)
    Xt = torch.as_tensor(X)
    diff = (Xt[:, None, :] - Xt[None, :, :]) / ls
    d2 = (diff * diff).sum(-1)
    K = _kernel(family, d2)
    K = K * (1.0 - torch.eye(n, dtype=torch.float64)) + torch.eye(n, dtype=torch.float64)  # exact unit diagonal
    tt = torch.zeros(n, dtype=torch.long) if tid is None else torch.as_tensor(tid, dtype=torch.long)
    K = K * B[tt][:, tt] + nz * torch.eye(n, dtype=torch.float64)
    L = torch.linalg.cholesky(K)
    r = (torch.as_tensor(y) - c).unsqueeze(-1)
    a = torch.cholesky_solve(r, L)
    mll = -0.5 * (r * a).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * math.log(2 * math.pi)
    mll.backward()
    return float(mll), t.grad.numpy().copy()


@pytest.mark.parametrize("family", ["matern52", "matern32", "rbf", "matern12"])
@pytest.mark.parametrize("tasks", [False, True])
def test_device_mll_and_gradient_match_autograd(family, tasks, cuda_device):
    rng = np.random.default_rng(3)
    if tasks:
        w = task_workload(N_per_task=60, n_tasks=3, d_num=5, n_per_task=25, seed=4)
        X = np.delete(w.train_x, w.task_col, axis=1)
        tid = np.rint(w.train_x[:, w.task_col]).astype(np.int32)
        T = 3
        A = rng.uniform(0.2, 1.0, size=(T, T))
        B = A @ A.T + np.diag(rng.uniform(0.1, 0.5, T))
    else:
        w = numeric_grid_workload(N=400, d=7, n=90, seed=5)
        X, tid, T, B = w.train_x, None, 1, np.array([[1.3]])
    y = (w.train_y - w.train_y.mean()) / w.train_y.std(ddof=1)
    d = X.shape[1]
    mll = DeviceMLL(X, y, tid, T, family, cuda_device)
    for trial in range(3):
        theta = np.concatenate([rng.uniform(0.3, 2.0, d), [rng.uniform(1e-3, 0.1)], [rng.normal(0, 0.3)], B.reshape(-1)])
        val, grad, ok = mll(theta)
        ref_val, ref_grad = _host_mll(X, y, tid, T, family, theta)
        assert ok
        assert abs(val - ref_val) <= 1e-9 * max(1.0, abs(ref_val))
        assert np.abs(grad - ref_grad).max() <= 1e-8 * max(1.0, np.abs(ref_grad).max()), (grad, ref_grad)


def test_not_positive_definite_is_reported(cuda_device):
    X = np.zeros((6, 2))  # six identical points, no noise to speak of: singular Gram matrix
    mll = DeviceMLL(X, np.zeros(6), None, 1, "rbf", cuda_device)
    _, _, ok = mll(np.array([1.0, 1.0, -1e-3, 0.0, 1.0]))
    assert not ok


@pytest.mark.parametrize("tasks", [False, True])
def test_device_fit_equals_host_fit(tasks, cuda_device):
    if tasks:
        w = task_workload(N_per_task=50, n_tasks=2, d_num=4, n_per_task=30, seed=7)
        active = [j for j in range(w.train_x.shape[1]) if j != w.task_col]
        tid, T = np.rint(w.train_x[:, w.task_col]).astype(int), 2
    else:
        w = numeric_grid_workload(N=500, d=6, n=80, seed=8)
        active, tid, T = list(range(6)), None, 1
    y = (w.train_y - w.train_y.mean()) / w.train_y.std(ddof=1)
    host = fit_map(w.train_x, y, active, tid, T, 200, mll_factory=HostMLL)
    dev = fit_map_hyperparameters_device(w.train_x, y, active, tid, T, 200, device=cuda_device)
    # a task parameter switches both to the leave-one-out pseudo-likelihood (presets/baybe.py:270-281)
    assert dev["criterion"] == host["criterion"] == ("loo" if tasks else "mll")
    # same objective, same optimiser, float64 on both sides: both runs stop within L-BFGS-B's termination
    # tolerance of the same optimum (the task parameters W, v have flat directions, so only the well-determined
    # hyper-parameters are compared by value)
    # (with tasks the objective is the leave-one-out pseudo-likelihood over 8 extra task parameters with flat
    # directions: the two float64 runs, whose objectives agree to 1e-9 at equal arguments
    # (test_device_loo_pseudo_likelihood_and_gradient_match_autograd), end the 200-iteration budget at different
    # points of the plateau -- both must have improved on the start point by a wide margin and land close)
    if tasks:
        assert abs(dev["objective"] - host["objective"]) <= 0.1 * max(1.0, abs(host["objective"]))
        assert dev["objective"] < 0.0 and host["objective"] < 0.0
    else:
        assert abs(dev["objective"] - host["objective"]) <= 2e-5 * max(1.0, abs(host["objective"]))
    if not tasks:
        assert np.allclose(dev["lengthscale"], host["lengthscale"], rtol=2e-2, atol=1e-3)
        assert np.isclose(dev["noise"], host["noise"], rtol=2e-2, atol=1e-5)
        assert np.isclose(dev["mean_const"], host["mean_const"], rtol=2e-2, atol=1e-3)


def test_surrogate_device_fit_and_host_fitted_hyperparameters_recommend_the_same_point(cuda_device):
    import pandas as pd

    from baybe_b200.recommenders import B200Recommender
    from baybe_b200.searchspace import NumericalDiscreteParameter, NumericalTarget, SearchSpace, SingleTargetObjective
    from baybe_b200.surrogates import GaussianProcessSurrogate

    ss = SearchSpace.from_product([NumericalDiscreteParameter(f"x{j}", np.linspace(0, 1, 7)) for j in range(4)])
    rows = ss.discrete.exp_rep.sample(n=25, random_state=2)
    comp = ss.transform(rows).to_numpy()
    meas = rows.assign(Yield=np.sin(3 * comp[:, 0]) + comp[:, 1] ** 2 - 0.5 * comp[:, 2] + 0.01 * np.arange(25) % 3)
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    rec = B200Recommender(surrogate_model=GaussianProcessSurrogate())  # MAP fit on the device
    torch.manual_seed(5)
    out_dev = rec.recommend(2, ss, obj, meas)
    hp = rec.surrogate_model.fitted_hyperparameters
    assert hp["criterion"] == "mll" and hp["noise"] >= 1e-4 and (hp["lengthscale"] >= 2.5e-2).all()
    # the same MAP objective driven by the float64 autograd twin on the host, its optimum handed over as fixed
    # hyper-parameters: same recommendation
    y = meas["Yield"].to_numpy()
    host = fit_map(comp, (y - y.mean()) / y.std(ddof=1), list(range(4)), mll_factory=HostMLL)
    rec2 = B200Recommender(surrogate_model=GaussianProcessSurrogate(hyperparameters={
        "lengthscale": host["lengthscale"], "noise": host["noise"], "mean_const": host["mean_const"]}))
    torch.manual_seed(5)
    out_host = rec2.recommend(2, ss, obj, meas)
    assert list(out_dev.index) == list(out_host.index)
    assert isinstance(out_dev, pd.DataFrame)


@pytest.mark.parametrize("family", ["matern52", "rbf"])
@pytest.mark.parametrize("tasks", [False, True])
def test_device_loo_pseudo_likelihood_and_gradient_match_autograd(family, tasks, cuda_device):
    """bb_fit_eval_loo against gpytorch's LeaveOneOutPseudoLikelihood formula evaluated by float64 autograd
    (sigma_i^2 = 1/[K^-1]_ii, mu_i = y_i - alpha_i sigma_i^2; tests/helpers.py::HostMLL, criterion="loo")."""
    rng = np.random.default_rng(11)
    if tasks:
        w = task_workload(N_per_task=60, n_tasks=3, d_num=5, n_per_task=25, seed=4)
        X = np.delete(w.train_x, w.task_col, axis=1)
        tid = np.rint(w.train_x[:, w.task_col]).astype(np.int32)
        T = 3
        A = rng.uniform(0.2, 1.0, size=(T, T))
        B = A @ A.T + np.diag(rng.uniform(0.1, 0.5, T))
    else:
        w = numeric_grid_workload(N=400, d=7, n=90, seed=5)
        X, tid, T, B = w.train_x, None, 1, np.array([[1.3]])
    y = (w.train_y - w.train_y.mean()) / w.train_y.std(ddof=1)
    d = X.shape[1]
    dev = DeviceMLL(X, y, tid, T, family, cuda_device, criterion="loo")
    ref = HostMLL(X, y, tid, T, family, criterion="loo")
    for trial in range(3):
        theta = np.concatenate([rng.uniform(0.3, 2.0, d), [rng.uniform(1e-3, 0.1)], [rng.normal(0, 0.3)], B.reshape(-1)])
        val, grad, ok = dev(theta)
        ref_val, ref_grad, ref_ok = ref(theta)
        assert ok and ref_ok
        assert abs(val - ref_val) <= 1e-9 * max(1.0, abs(ref_val)), (val, ref_val)
        assert np.abs(grad - ref_grad).max() <= 1e-7 * max(1.0, np.abs(ref_grad).max()), (grad, ref_grad)
    # brute force: refit n times leaving one point out (n small), exact predictive log densities
    n_small = 12
    Xs, ys = X[:n_small], y[:n_small]
    ts = None if tid is None else tid[:n_small]
    theta = np.concatenate([np.full(d, 0.9), [0.05], [0.1], B.reshape(-1)])
    val, _, ok = DeviceMLL(Xs, ys, ts, T, family, cuda_device, criterion="loo")(theta)
    Xt = torch.as_tensor(Xs)
    diff = (Xt[:, None, :] - Xt[None, :, :]) / 0.9
    K = _kernel(family, (diff * diff).sum(-1))
    tt = torch.zeros(n_small, dtype=torch.long) if ts is None else torch.as_tensor(ts, dtype=torch.long)
    K = K * torch.as_tensor(B)[tt][:, tt] + 0.05 * torch.eye(n_small, dtype=torch.float64)
    yt = torch.as_tensor(ys) - 0.1
    total = 0.0
    for i in range(n_small):
        keep = [j for j in range(n_small) if j != i]
        Kk = K[keep][:, keep]
        ki = K[i, keep]
        sol = torch.linalg.solve(Kk, torch.stack([yt[keep], ki], dim=1))
        mu_i = float(ki @ sol[:, 0])
        var_i = float(K[i, i] - ki @ sol[:, 1])
        total += -0.5 * math.log(2 * math.pi * var_i) - 0.5 * (float(yt[i]) - mu_i) ** 2 / var_i
    assert ok and abs(val - total) <= 1e-8 * max(1.0, abs(total)), (val, total)


@pytest.mark.parametrize("preset", ["CHEN", "EDBO", "custom_rbf"])
def test_device_fit_equals_host_fit_for_other_presets(preset, cuda_device):
    """Output scale, other priors/start values and kernel families go through the same device objective."""
    from baybe_b200.kernels import RBFKernel, ScaleKernel, gp_preset, resolve_kernel
    from baybe_b200.priors import GammaPrior, LogNormalPrior

    w = numeric_grid_workload(N=500, d=5, n=70, seed=9)
    y = (w.train_y - w.train_y.mean()) / w.train_y.std(ddof=1)
    if preset == "custom_rbf":
        cfg = resolve_kernel(ScaleKernel(RBFKernel(LogNormalPrior(0.2, 0.8)), GammaPrior(2.0, 1.0)), GammaPrior(1.1, 0.05))
    else:
        cfg = gp_preset(preset, 5)
    host = fit_map(w.train_x, y, list(range(5)), None, 1, 200, config=cfg, mll_factory=HostMLL)
    dev = fit_map_hyperparameters_device(w.train_x, y, list(range(5)), None, 1, 200, device=cuda_device, config=cfg)
    assert dev["family"] == cfg.family and dev["outputscale"] is not None
    assert abs(dev["objective"] - host["objective"]) <= 2e-5 * max(1.0, abs(host["objective"]))
    assert np.allclose(dev["lengthscale"], host["lengthscale"], rtol=5e-2, atol=1e-3)
    assert np.isclose(dev["outputscale"], host["outputscale"], rtol=5e-2)
    assert np.isclose(dev["noise"], host["noise"], rtol=5e-2, atol=1e-5)


def test_surrogate_with_kernel_object_and_preset_names(cuda_device):
    from baybe_b200.kernels import MaternKernel, ScaleKernel
    from baybe_b200.priors import GammaPrior
    from baybe_b200.recommenders import B200Recommender
    from baybe_b200.searchspace import NumericalDiscreteParameter, NumericalTarget, SearchSpace, SingleTargetObjective
    from baybe_b200.surrogates import GaussianProcessSurrogate

    ss = SearchSpace.from_product([NumericalDiscreteParameter(f"x{j}", np.linspace(0, 1, 6)) for j in range(3)])
    rows = ss.discrete.exp_rep.sample(n=20, random_state=4)
    comp = ss.transform(rows).to_numpy()
    meas = rows.assign(Yield=np.cos(2 * comp[:, 0]) + comp[:, 1] - comp[:, 2] ** 2)
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    kern = ScaleKernel(MaternKernel(1.5, GammaPrior(3.0, 6.0)), GammaPrior(2.0, 0.5))
    for kf in (kern, "EDBO", "CHEN"):
        rec = B200Recommender(surrogate_model=GaussianProcessSurrogate(kernel_or_factory=kf))
        out = rec.recommend(2, ss, obj, meas)
        hp = rec.surrogate_model.fitted_hyperparameters
        assert len(out) == 2 and hp["outputscale"] > 0
        assert hp["family"] == ("matern32" if kf is kern else "matern52")
        assert rec.surrogate_model.device_gp.family == hp["family"]
    with pytest.raises(TypeError):
        GaussianProcessSurrogate(kernel_or_factory=3.0).fit(ss, obj, meas)
