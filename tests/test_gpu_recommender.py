"""GPU tests of the plugin surfaces end to end: B200Recommender.recommend / posterior_stats /
acquisition_values against the CPU oracle's optimize_acqf_discrete on BASELINE config 1
(3-parameter discrete space, 15 measurements) and on a transfer-learning (task) space."""
from __future__ import annotations

import numpy as np
import pandas as pd
import pytest
import torch

import oracle
from baybe_b200.acquisition import qLogEI
from baybe_b200.recommenders import B200Recommender, NotEnoughPointsLeftError
from baybe_b200.searchspace import (CategoricalParameter, NumericalDiscreteParameter, NumericalTarget,
                                    SearchSpace, SingleTargetObjective, TaskParameter)
from baybe_b200.surrogates import GaussianProcessSurrogate

pytestmark = pytest.mark.gpu


def _space():
    return SearchSpace.from_product([
        CategoricalParameter("Granularity", ["coarse", "medium", "fine"]),
        NumericalDiscreteParameter("Pressure", [1, 2, 5, 10, 20, 50, 80, 100]),
        NumericalDiscreteParameter("Temperature", np.linspace(90, 160, 8)),
    ])


def _measure(ss: SearchSpace, n: int, seed: int) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    rows = ss.discrete.exp_rep.sample(n=n, random_state=seed)
    comp = ss.transform(rows).to_numpy()
    y = 50 + 10 * comp[:, 0] - 5 * comp[:, 2] + 0.3 * comp[:, 3] - 0.002 * (comp[:, 4] - 120) ** 2
    return rows.assign(Yield=y + rng.standard_normal(n))


HP = {"lengthscale": np.array([0.9, 1.1, 0.8, 0.6, 0.7]), "noise": 0.02, "mean_const": 0.1}


def _oracle_model(ss, meas, hp, task=None):
    comp = ss.transform(meas).to_numpy()
    active = [j for j in range(comp.shape[1]) if j != ss.task_idx]
    spec = oracle.KernelSpec("matern52", hp["lengthscale"], active, task_idx=ss.task_idx,
                             task_covar=hp.get("task_covar"))
    return oracle.build_model(spec, comp, meas["Yield"].to_numpy(), ss.scaling_bounds.to_numpy(),
                              noise=hp["noise"], mean_const=hp["mean_const"])


@pytest.mark.parametrize("minimize", [False, True])
@pytest.mark.parametrize("batch_size", [1, 3])
def test_recommend_matches_oracle_greedy(minimize, batch_size, cuda_device):
    ss = _space()
    meas = _measure(ss, 15, seed=3)
    obj = SingleTargetObjective(NumericalTarget("Yield", minimize=minimize))
    rec = B200Recommender(surrogate_model=GaussianProcessSurrogate(hyperparameters=HP))
    torch.manual_seed(1337)
    out = rec.recommend(batch_size, ss, obj, meas)
    assert list(out.columns) == list(ss.discrete.exp_rep.columns) and len(out) == batch_size
    assert out.index.isin(ss.discrete.exp_rep.index).all() and out.index.is_unique
    # oracle with the same sampler seed (drawn from torch's global RNG like botorch does)
    torch.manual_seed(1337)
    seed = int(torch.randint(0, 1_000_000, (1,)))
    om = _oracle_model(ss, meas, HP)
    acq = oracle.AcqSpec("qLogEI", obj_scale=-1.0 if minimize else 1.0)
    acq.best_f = oracle.best_f_from_training(om, ss.transform(meas).to_numpy(), acq)
    idx, vals = oracle.optimize_acqf_discrete(om, acq, ss.discrete.comp_rep.to_numpy(), q=batch_size,
                                              sampler_seed=seed)
    assert list(out.index) == list(ss.discrete.comp_rep.index[idx])
    assert np.allclose(rec._last_acq_values, vals, rtol=2e-3, atol=5e-3)


def test_posterior_stats_and_acquisition_values(cuda_device):
    ss = _space()
    meas = _measure(ss, 15, seed=4)
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    sur = GaussianProcessSurrogate(hyperparameters=HP)
    sur.fit(ss, obj, meas)
    cands = ss.discrete.exp_rep.iloc[::7]
    stats = sur.posterior_stats(cands, stats=("mean", "std", "var", 0.9))
    assert list(stats.columns) == ["Yield_mean", "Yield_std", "Yield_var", "Yield_Q_0.9"]
    assert stats.index.equals(cands.index) and not stats.isna().any().any()
    om = _oracle_model(ss, meas, HP)
    mu, var = oracle.posterior(om, ss.transform(cands).to_numpy())
    assert np.allclose(stats["Yield_mean"], mu.numpy(), rtol=1e-5, atol=1e-3)
    assert np.allclose(stats["Yield_var"], var.numpy(), rtol=1e-3, atol=1e-3)
    assert np.allclose(stats["Yield_Q_0.9"], mu.numpy() + var.sqrt().numpy() * 1.2815515655446004, rtol=1e-4, atol=1e-2)
    with pytest.raises(ValueError):
        sur.posterior_stats(cands, stats=(1.5,))
    rec = B200Recommender(surrogate_model=sur)
    torch.manual_seed(5)
    acqv = rec.acquisition_values(cands, ss, obj, meas)
    assert isinstance(acqv, pd.Series) and acqv.index.equals(cands.index) and np.isfinite(acqv).all()
    # per-row values against the oracle with the same sampler seed (a14: acquisition/base.py:112-159)
    torch.manual_seed(5)
    seed = int(torch.randint(0, 1_000_000, (1,)))
    oacq = oracle.AcqSpec("qLogEI")
    oacq.best_f = oracle.best_f_from_training(om, ss.transform(meas).to_numpy(), oacq)
    z = oracle.sobol_normal_samples(512, 1, seed)[:, 0]
    from tests.helpers import score_bounds

    ref, bound = score_bounds(om, oacq, ss.transform(cands).to_numpy(), z)
    err = np.abs(acqv.to_numpy() - ref.numpy())
    assert (err <= bound.numpy()).all(), float(err.max())


def test_joint_posterior_of_a_small_batch(cuda_device):
    """Surrogate.posterior(candidates, joint=True) (surrogates/base.py:213-247): mean and FULL covariance of one
    q-batch against the oracle's joint posterior."""
    ss = _space()
    meas = _measure(ss, 15, seed=6)
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    sur = GaussianProcessSurrogate(hyperparameters=HP)
    sur.fit(ss, obj, meas)
    batch = ss.discrete.exp_rep.iloc[[3, 50, 51, 120, 191]]
    post = sur.posterior(batch, joint=True)
    om = _oracle_model(ss, meas, HP)
    m_ref, cov_ref = oracle.posterior_joint(om, ss.transform(batch).to_numpy())
    assert post.mean.shape == (1, 5, 1) and post.covariance.shape == (5, 5)
    assert np.allclose(post.mean.reshape(-1).double().cpu().numpy(), m_ref.numpy(), rtol=1e-5, atol=1e-3)
    scale = float(cov_ref.diagonal().max())
    assert np.abs(post.covariance.double().cpu().numpy() - cov_ref.numpy()).max() <= 2e-4 * scale
    assert np.allclose(post.variance.reshape(-1).double().cpu().numpy(), cov_ref.diagonal().numpy(), rtol=1e-3, atol=2e-4 * scale)
    _ = post.mvn  # a valid multivariate normal
    with pytest.raises(NotImplementedError):
        sur.posterior(ss.discrete.exp_rep.iloc[:40], joint=True)


def test_filtered_candidate_set_becomes_a_position_mask(cuda_device):
    """a1/a2: ``Campaign.recommend`` hands the recommender a FilteredSubspaceDiscrete whose get_candidates() returns
    ``exp_rep.loc[mask]`` (campaign.py:549-572, _filtered.py:41-43).  The engine must only pick rows of that subset,
    and pick the best of THEM."""
    ss = _space()
    meas = _measure(ss, 15, seed=3)
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    rec = B200Recommender(surrogate_model=GaussianProcessSurrogate(hyperparameters=HP))
    torch.manual_seed(11)
    full = rec.recommend(1, ss, obj, meas)
    torch.manual_seed(11)
    vals = rec.acquisition_values(ss.discrete.exp_rep, ss, obj, meas)
    assert full.index[0] == vals.idxmax()
    # drop the 40 best rows and every third row: the recommendation must be the best remaining one
    drop = set(vals.sort_values(ascending=False).index[:40]) | set(ss.discrete.exp_rep.index[::3])
    mask = ~ss.discrete.exp_rep.index.isin(list(drop))
    candidates_exp = ss.discrete.exp_rep.loc[mask]
    surrogate = rec.get_surrogate(ss, obj, meas)
    cfg = rec._get_acquisition_function(obj).to_engine(surrogate, ss, obj, meas, None)
    rec._context = (cfg, ss, None)
    torch.manual_seed(11)
    idxs = rec._recommend_discrete(ss.discrete, candidates_exp, 3)
    assert set(idxs) <= set(candidates_exp.index) and len(set(idxs)) == 3
    assert idxs[0] == vals.loc[candidates_exp.index].idxmax()
    with pytest.raises(ValueError):
        rec._recommend_discrete(ss.discrete, candidates_exp.rename(index={candidates_exp.index[0]: 10**9}), 1)


def test_fit_is_cached_and_refit_on_new_data(cuda_device):
    ss = _space()
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    sur = GaussianProcessSurrogate(hyperparameters=HP)
    meas = _measure(ss, 10, seed=1)
    sur.fit(ss, obj, meas)
    gp1 = sur.device_gp
    sur.fit(ss, obj, meas.copy())
    assert sur.device_gp is gp1  # unchanged context -> no retraining (surrogates/base.py:419-424)
    sur.fit(ss, obj, _measure(ss, 12, seed=2))
    assert sur.device_gp is not gp1


def test_pending_experiments_are_not_recommended_again(cuda_device):
    """reference: tests/test_pending_experiments.py:100-127"""
    ss = _space()
    meas = _measure(ss, 15, seed=6)
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    rec = B200Recommender(surrogate_model=GaussianProcessSurrogate(hyperparameters=HP),
                          acquisition_function=qLogEI())
    torch.manual_seed(11)
    rec1 = rec.recommend(3, ss, obj, meas)
    torch.manual_seed(11)
    rec2 = rec.recommend(3, ss, obj, meas, pending_experiments=rec1)
    assert len(set(rec1.index) & set(rec2.index)) == 0


def test_full_fit_and_transfer_learning_space(cuda_device):
    ss = SearchSpace.from_product([
        NumericalDiscreteParameter("x0", np.linspace(0, 1, 9)),
        NumericalDiscreteParameter("x1", np.linspace(0, 1, 9)),
        TaskParameter("Function", ["source", "target"]),
    ])
    rng = np.random.default_rng(0)
    rows = ss.discrete.exp_rep.sample(n=30, random_state=0)
    shift = np.where(rows["Function"] == "target", 0.3, 0.0)
    y = np.sin(3 * rows["x0"]) + rows["x1"] ** 2 + shift + 0.02 * rng.standard_normal(30)
    meas = rows.assign(Yield=y)
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    rec = B200Recommender()  # MAP-fitted hyper-parameters incl. the task covariance
    torch.manual_seed(3)
    out = rec.recommend(2, ss, obj, meas)
    hp = rec.surrogate_model.fitted_hyperparameters
    assert hp["task_covar"].shape == (2, 2) and hp["noise"] >= 1e-4
    # parity of the scoring path under the fitted hyper-parameters
    om = _oracle_model(ss, meas, hp)
    torch.manual_seed(3)
    seed = int(torch.randint(0, 1_000_000, (1,)))
    acq = oracle.AcqSpec("qLogEI")
    acq.best_f = oracle.best_f_from_training(om, ss.transform(meas).to_numpy(), acq)
    idx, _ = oracle.optimize_acqf_discrete(om, acq, ss.discrete.comp_rep.to_numpy(), q=2, sampler_seed=seed)
    assert list(out.index) == list(ss.discrete.comp_rep.index[idx])
    with pytest.raises(NotEnoughPointsLeftError):
        rec.recommend(10_000, ss, obj, meas)


@pytest.mark.parametrize("batch_size", [1, 3])
def test_recommend_on_binary_fingerprint_space_uses_bit_layout(batch_size, cuda_device):
    """A comp-rep that is all 0/1 and >= 256 columns wide is kept bit-packed on the device and scored by
    the wide-feature path; the recommendation equals the oracle's arg-max over the float matrix."""
    from baybe_b200 import recommenders as R

    rng = np.random.default_rng(5)
    d, N, n = 320, 1500, 40
    bits = (rng.random((N, d)) < 0.1).astype(np.float64)
    cols = [f"b{j}" for j in range(d)]
    df = pd.DataFrame(bits, columns=cols)
    ss = SearchSpace.from_dataframe(df, [NumericalDiscreteParameter(c, [0.0, 1.0]) for c in cols])
    rows = ss.discrete.exp_rep.sample(n=n, random_state=1)
    comp = ss.transform(rows).to_numpy()
    y = comp[:, :40].sum(axis=1) - 0.5 * comp[:, 40:80].sum(axis=1) + 0.05 * rng.standard_normal(n)
    meas = rows.assign(Yield=y)
    hp = {"family": "rbf", "lengthscale": np.full(d, 6.0), "noise": 0.01, "mean_const": 0.0, "outputscale": 1.3}
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    rec = B200Recommender(surrogate_model=GaussianProcessSurrogate(hyperparameters=hp))
    torch.manual_seed(11)
    out = rec.recommend(batch_size, ss, obj, meas)
    x_dev, _ = R._cache.get(ss.discrete, cuda_device, 0, N)
    assert x_dev.dtype == torch.uint8 and x_dev.shape == (N, d // 8)
    torch.manual_seed(11)
    seed = int(torch.randint(0, 1_000_000, (1,)))
    spec = oracle.KernelSpec("rbf", hp["lengthscale"], list(range(d)), outputscale=1.3)
    om = oracle.build_model(spec, comp, meas["Yield"].to_numpy(), ss.scaling_bounds.to_numpy(), noise=0.01,
                            mean_const=0.0)
    acq = oracle.AcqSpec("qLogEI")
    acq.best_f = oracle.best_f_from_training(om, comp, acq)
    idx, vals = oracle.optimize_acqf_discrete(om, acq, ss.discrete.comp_rep.to_numpy(), q=batch_size,
                                              sampler_seed=seed)
    assert list(out.index) == list(ss.discrete.comp_rep.index[idx])
    assert np.allclose(rec._last_acq_values, vals, rtol=2e-3, atol=5e-3)


def test_joint_acquisition_value_of_a_batch(cuda_device):
    """AcquisitionFunction.evaluate(jointly=True) / BayesianRecommender.joint_acquisition_value
    (acquisition/base.py:112-159, bayesian/base.py:239-277): one q-batch value for the whole candidate set."""
    from baybe_b200.acquisition import IncompatibleAcquisitionFunctionError, UpperConfidenceBound

    ss = _space()
    meas = _measure(ss, 15, seed=6)
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    rec = B200Recommender(surrogate_model=GaussianProcessSurrogate(hyperparameters=HP))
    batch = ss.discrete.exp_rep.iloc[[5, 77, 130, 9]]
    pend = ss.discrete.exp_rep.iloc[[40]]
    torch.manual_seed(21)
    got = rec.joint_acquisition_value(batch, ss, obj, meas, pending_experiments=pend)
    torch.manual_seed(21)
    seed = int(torch.randint(0, 1_000_000, (1,)))
    om = _oracle_model(ss, meas, HP)
    acq = oracle.AcqSpec("qLogEI")
    acq.best_f = oracle.best_f_from_training(om, ss.transform(meas).to_numpy(), acq)
    rows = np.concatenate([ss.transform(batch).to_numpy(), ss.transform(pend).to_numpy()], axis=0)
    z = oracle.sobol_normal_samples(512, len(rows), seed)
    ref = float(oracle.acq_values_joint(om, acq, rows[:1], rows[1:], z)[0])
    assert isinstance(got, float) and abs(got - ref) <= 5e-3 + 2e-3 * abs(ref)
    # a single point valued "jointly" is its ordinary q=1 value
    torch.manual_seed(21)
    one = rec.joint_acquisition_value(batch.iloc[:1], ss, obj, meas)
    torch.manual_seed(21)
    assert abs(one - float(rec.acquisition_values(batch.iloc[:1], ss, obj, meas).iloc[0])) < 1e-6
    with pytest.raises(IncompatibleAcquisitionFunctionError):
        rec.joint_acquisition_value(batch, ss, obj, meas, acquisition_function=UpperConfidenceBound())
