"""CPU tests of the host-side mirror of the reference's plugin surfaces (no GPU needed)."""
from __future__ import annotations

import numpy as np
import pandas as pd
import pytest
import torch

from baybe_b200.acquisition import (UCB, IncompatibleAcquisitionFunctionError, PosteriorStandardDeviation,
                                    convert_acqf, qLogEI, qUCB)
from baybe_b200.engine import AcqConfig, pack_best, sobol_normal_samples, unpack_best
from baybe_b200.recommenders import B200Recommender, shard_bounds
from baybe_b200.searchspace import (CategoricalParameter, NumericalDiscreteParameter, NumericalTarget,
                                    SearchSpace, SingleTargetObjective, TaskParameter, objective_affine)
from baybe_b200.surrogates import GaussianProcessSurrogate, ModelNotTrainedError, fit_map_hyperparameters


def _space():
    return SearchSpace.from_product([
        CategoricalParameter("Granularity", ["coarse", "medium", "fine"]),
        NumericalDiscreteParameter("Pressure", [1, 5, 10]),
        NumericalDiscreteParameter("Temperature", [90, 105, 120, 160]),
    ])


def test_searchspace_encodings_and_bounds():
    ss = _space()
    assert len(ss.discrete.exp_rep) == 36 and ss.discrete.comp_rep.shape == (36, 5)
    assert ss.comp_rep_columns == ("Granularity_coarse", "Granularity_medium", "Granularity_fine",
                                   "Pressure", "Temperature")
    b = ss.scaling_bounds.to_numpy()
    assert b[0].tolist() == [0, 0, 0, 1, 90] and b[1].tolist() == [1, 1, 1, 10, 160]
    row = ss.transform(pd.DataFrame({"Granularity": ["fine"], "Pressure": [5], "Temperature": [160]}))
    assert row.to_numpy().tolist() == [[0, 0, 1, 5, 160]]
    assert ss.task_idx is None and ss.n_tasks == 1
    with pytest.raises(ValueError):
        ss.transform(pd.DataFrame({"Granularity": ["x"], "Pressure": [5], "Temperature": [160]}))


def test_task_parameter_uses_sorted_integer_codes():
    ss = SearchSpace.from_product([NumericalDiscreteParameter("x", [0, 1]),
                                   TaskParameter("Function", ["zeta", "alpha"])])
    assert ss.task_idx == 1 and ss.n_tasks == 2
    comp = ss.transform(pd.DataFrame({"x": [0.0, 1.0], "Function": ["zeta", "alpha"]}))
    assert comp["Function"].tolist() == [1.0, 0.0]  # sorted labels: alpha=0, zeta=1


def test_objective_orientation():
    assert objective_affine(SingleTargetObjective(NumericalTarget("y")))[:2] == (1.0, 0.0)
    assert objective_affine(SingleTargetObjective(NumericalTarget("y", minimize=True)))[:2] == (-1.0, -0.0)


def test_acquisition_specs_follow_the_reference_flags():
    assert qLogEI().supports_batching and qLogEI().supports_pending_experiments and not qLogEI().is_analytic
    assert UCB().is_analytic and not UCB().supports_batching
    assert convert_acqf("qLogEI") == qLogEI() and convert_acqf("qUpperConfidenceBound") == qUCB()
    assert qUCB(beta=1).beta == 1.0 and PosteriorStandardDeviation(maximize=False).maximize is False
    with pytest.raises(ValueError):
        convert_acqf("qKG")
    with pytest.raises(IncompatibleAcquisitionFunctionError):
        UCB().to_engine(None, None, SingleTargetObjective(NumericalTarget("y")), None, pd.DataFrame({"a": [1]}))
    with pytest.raises(ValueError):
        AcqConfig(kind="nope")


def test_packed_key_order_is_score_then_lowest_index():
    keys = [pack_best(s, i) for s, i in [(1.5, 7), (1.5, 3), (-2.0, 0), (0.0, 9), (-0.0, 1), (float("-inf"), 2)]]
    assert max(keys) == keys[1]
    assert unpack_best(keys[1]) == (1.5, 3)
    assert unpack_best(-(1 << 63)) == (float("-inf"), -1)
    rng = np.random.default_rng(0)
    vals = rng.standard_normal(1000).astype(np.float32)
    vals[[10, 500]] = vals.max() + 1
    best = max(pack_best(float(v), i) for i, v in enumerate(vals))
    assert unpack_best(best)[1] == 10


def test_shard_bounds_cover_all_rows_once():
    for n, w in [(10, 3), (1_000_000, 8), (7, 8), (0, 2)]:
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_sobol_base_samples_match_botorch_recipe():
    z = sobol_normal_samples(512, 3, seed=5)
    assert z.shape == (512, 3) and z.dtype == torch.float64
    assert torch.equal(z, sobol_normal_samples(512, 3, seed=5))
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1) < 0.03


def test_map_fit_recovers_sensible_hyperparameters():
    rng = np.random.default_rng(1)
    X = rng.uniform(size=(60, 3))
    y = np.sin(6 * X[:, 0]) + 0.02 * rng.standard_normal(60)  # only dim 0 matters
    hp = fit_map_hyperparameters(X, (y - y.mean()) / y.std(ddof=1), [0, 1, 2])
    ls = hp["lengthscale"]
    assert ls[0] < ls[1] and ls[0] < ls[2]  # ARD: the relevant dimension gets the short lengthscale
    assert 1e-4 <= hp["noise"] < 0.1 and np.all(ls >= 2.5e-2)


def test_surrogate_and_recommender_fail_loudly_without_fit_or_gpu():
    s = GaussianProcessSurrogate()
    with pytest.raises(ModelNotTrainedError):
        s.posterior(pd.DataFrame({"a": [1]}))
    with pytest.raises(ImportError):
        s.to_botorch()
    r = B200Recommender()
    ss = _space()
    with pytest.raises(NotImplementedError):
        r.recommend(1, ss)  # no objective
    obj = SingleTargetObjective(NumericalTarget("Yield"))
    with pytest.raises(NotImplementedError):
        r.recommend(1, ss, obj, pd.DataFrame())  # no data
    meas = ss.discrete.exp_rep.iloc[:5].assign(Yield=[1.0, 2.0, 3.0, 2.5, 0.5])
    with pytest.raises(IncompatibleAcquisitionFunctionError):
        B200Recommender(acquisition_function="UCB").recommend(2, ss, obj, meas)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            r.recommend(1, ss, obj, meas)


def test_bit_packing_round_trip_and_layout():
    """BB_BITS_U8: feature j lives in bit (j & 7) of byte j >> 3; padding bits of the last byte are 0."""
    from baybe_b200.bits import pack_bits, unpack_bits

    rng = np.random.default_rng(0)
    x = (rng.random((37, 77)) < 0.3).astype(np.float64)
    p = pack_bits(x)
    assert p.dtype == np.uint8 and p.shape == (37, 10)
    assert np.array_equal(unpack_bits(p, 77), x)
    j = 13
    assert np.array_equal((p[:, j >> 3] >> (j & 7)) & 1, x[:, j].astype(np.uint8))
    assert (p[:, -1] >> 5 == 0).all()  # 77 = 9*8 + 5 valid bits in the last byte


def test_fingerprint_workload_shape():
    from baybe_b200.synthetic import fingerprint_workload

    w = fingerprint_workload(N=300, d=256, n=64, seed=3)
    assert w.candidates.shape == (300, 256) and set(np.unique(w.candidates)) <= {0.0, 1.0}
    assert w.family == "rbf" and w.outputscale == 1.0 and w.train_x.shape == (64, 256)
    assert np.allclose(w.lengthscale, np.sqrt(256) * 0.4)
