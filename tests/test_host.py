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
from baybe_b200.surrogates import GaussianProcessSurrogate, ModelNotTrainedError, fit_map
from tests.helpers import HostMLL


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
    hp = fit_map(X, (y - y.mean()) / y.std(ddof=1), [0, 1, 2], mll_factory=HostMLL)
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


def test_priors_match_scipy_and_their_gradients():
    """Priors of baybe/priors/basic.py: log densities (up to a constant) against scipy.stats, gradients
    against central differences."""
    from scipy import stats

    from baybe_b200.priors import (GammaPrior, HalfCauchyPrior, HalfNormalPrior, LogNormalPrior, NormalPrior,
                                   SmoothedBoxPrior)

    x = np.array([0.07, 0.6, 1.3, 4.0])
    cases = [
        (GammaPrior(3.0, 2.0), stats.gamma(a=3.0, scale=0.5)),
        (LogNormalPrior(0.3, 0.7), stats.lognorm(s=0.7, scale=np.exp(0.3))),
        (NormalPrior(0.5, 1.2), stats.norm(0.5, 1.2)),
        (HalfNormalPrior(0.8), stats.halfnorm(scale=0.8)),
        (HalfCauchyPrior(1.5), stats.halfcauchy(scale=1.5)),
    ]
    for prior, ref in cases:
        diff = prior.log_prob(x) - ref.logpdf(x)
        assert np.allclose(diff, diff[0], atol=1e-12), type(prior).__name__  # equal up to the constant
        fd = (prior.log_prob(x + 1e-6) - prior.log_prob(x - 1e-6)) / 2e-6
        assert np.allclose(prior.grad(x), fd, rtol=1e-6, atol=1e-8), type(prior).__name__
    assert GammaPrior(3.0, 2.0).mode == 1.0 and np.isclose(LogNormalPrior(0.3, 0.7).mode, np.exp(0.3 - 0.49))
    box = SmoothedBoxPrior(0.5, 2.0, 0.1)
    assert box.log_prob(np.array([1.0]))[0] == 0.0 and box.log_prob(np.array([2.3]))[0] == pytest.approx(-4.5)
    with pytest.raises(ValueError):
        SmoothedBoxPrior(2.0, 1.0)


def test_kernel_specs_and_presets():
    """kernels/basic.py + composite.py mirrors and the presets built from them."""
    import math

    from baybe_b200.kernels import MaternKernel, RBFKernel, ScaleKernel, gp_preset, resolve_kernel
    from baybe_b200.priors import GammaPrior

    assert MaternKernel("3/2").family == "matern32" and MaternKernel().family == "matern52"
    with pytest.raises(ValueError):
        MaternKernel(2.0)
    with pytest.raises(NotImplementedError):
        ScaleKernel(ScaleKernel(RBFKernel()))
    cfg = resolve_kernel(ScaleKernel(RBFKernel(GammaPrior(3.0, 6.0)), GammaPrior(2.0, 0.15), 5.0))
    assert (cfg.family, cfg.outputscale, cfg.lengthscale_initial_value, cfg.outputscale_initial_value) == \
           ("rbf", True, pytest.approx(1.0 / 3.0), 5.0)
    b = gp_preset("BAYBE", 20)  # presets/baybe.py:95-99,134-144
    assert b.family == "matern52" and not b.outputscale and b.lengthscale_lower == 2.5e-2
    assert b.lengthscale_initial_value == pytest.approx(math.exp(math.sqrt(2.0) - 3.0) * math.sqrt(20))
    assert b.noise_initial_value == pytest.approx(math.exp(-5.0))
    c = gp_preset("chen", 9)  # presets/chen.py:35-61
    assert c.outputscale and c.lengthscale_initial_value == pytest.approx(0.4 * 3 + 4.0) and c.noise_prior is None
    assert gp_preset("EDBO", 3).lengthscale_initial_value == 0.2 and gp_preset("EDBO", 30).outputscale_initial_value == 20.0
    with pytest.raises(ValueError):
        gp_preset("nope", 3)


def test_host_fit_with_presets_improves_on_the_start_point():
    from baybe_b200.kernels import gp_preset
    from baybe_b200.synthetic import numeric_grid_workload

    w = numeric_grid_workload(N=200, d=3, n=30, seed=2)
    y = (w.train_y - w.train_y.mean()) / w.train_y.std(ddof=1)
    for name in ("BAYBE", "CHEN", "EDBO"):
        cfg = gp_preset(name, 3)
        hp = fit_map(w.train_x, y, [0, 1, 2], None, 1, 60, config=cfg, mll_factory=HostMLL)
        assert hp["family"] == "matern52" and (hp["outputscale"] is not None) == cfg.outputscale
        assert hp["noise"] >= 1e-4 and (hp["lengthscale"] >= cfg.lengthscale_lower).all()
        # the optimum is at least as good as the start point
        mll = HostMLL(w.train_x, y, None, 1, cfg.family)
        os0 = cfg.outputscale_initial_value if cfg.outputscale else 1.0
        v0, _, ok = mll(np.concatenate([np.full(3, cfg.lengthscale_initial_value), [cfg.noise_initial_value, 0.0, os0]]))
        lp0 = sum(p.log_prob(np.asarray(x)).sum() for p, x in [
            (cfg.lengthscale_prior, np.full(3, cfg.lengthscale_initial_value)), (cfg.noise_prior, cfg.noise_initial_value),
            (cfg.outputscale_prior if cfg.outputscale else None, os0)] if p is not None)
        assert ok and hp["objective"] <= -(v0 + lp0) / 30 + 1e-9


def test_packed_keys_and_topk_merge_properties():
    """Property tests (hypothesis): the int64 key order is (score, then lowest index) for any finite floats, and the
    cross-rank top-k merge equals a global sort."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from baybe_b200.engine import pack_best, unpack_best
    from baybe_b200.recommenders import merge_topk_across_ranks

    f32 = st.floats(width=32, allow_nan=False, allow_infinity=False)

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.tuples(f32, st.integers(0, 2**31 - 1)), min_size=1, max_size=20))
    def keys(pairs):
        best_key = max(pack_best(s, i) for s, i in pairs)
        smax = max(np.float32(s) for s, _ in pairs)
        want_idx = min(i for s, i in pairs if np.float32(s) == smax or (smax == 0 and np.float32(s) == 0))
        val, idx = unpack_best(best_key)
        assert np.float32(val) == smax or (smax == 0 and val == 0)
        if not (smax == 0):  # +0.0 / -0.0 are distinct keys (bitwise order); everything else ties on equal floats
            assert idx == want_idx

    @settings(max_examples=100, deadline=None)
    @given(st.lists(f32, min_size=1, max_size=40), st.integers(1, 8))
    def merge(scores, k):
        s = torch.tensor(scores, dtype=torch.float32)
        kk = min(k, len(scores))
        v, i = torch.topk(s, kk)
        pad = k - kk
        v = torch.cat([v, torch.full((pad,), -float("inf"))])
        i = torch.cat([i, torch.full((pad,), -1, dtype=torch.int64)])
        gv, gi = merge_topk_across_ranks(v, i, k)
        order = np.lexsort((np.arange(len(scores)), -np.asarray(scores, dtype=np.float32)))[:kk]
        assert np.array_equal(gv[:kk].numpy(), np.asarray(scores, dtype=np.float32)[order])
        assert (gi[kk:] == -1).all()

    keys()
    merge()


def test_layout_detection_never_copies_and_reads_strides():
    """a one-row slice of a column-major matrix -- the layout the reference hands over,
    baybe/utils/dataframe.py:68-81 -- must be reported as column-major with ld = N (it used to be taken for a
    row, and a local .contiguous() hid the mismatch from the caller's data_ptr)."""
    from baybe_b200 import _lib
    from baybe_b200.engine import _as_device_matrix, _layout_of

    L = _lib.LAYOUT
    N, d = 7, 5
    row = torch.arange(N * d, dtype=torch.float32).reshape(N, d)
    col = row.t().contiguous().t()  # strides (1, N)
    assert _layout_of(row) == (L["row_f32"], d)
    assert _layout_of(col) == (L["col_f32"], N)
    assert _layout_of(col[:1]) == (L["col_f32"], N)      # one row of a column-major matrix
    assert _layout_of(row[:1]) == (L["row_f32"], d)
    assert _layout_of(row.double()[:, :3]) == (L["row_f64"], d)  # padded leading dimension
    assert _layout_of(row[:, :1]) == (L["row_f32"], d)   # one column, rows d apart
    assert _layout_of(torch.zeros(1, 1)) == (L["row_f32"], 1)
    with pytest.raises(ValueError):
        _layout_of(row[:, ::2])
    # the only place a copy may happen is _as_device_matrix, whose result the caller uses
    fixed = _as_device_matrix(row[:1, ::2], torch.device("cpu"), 3)
    assert fixed.is_contiguous() and _layout_of(fixed) == (L["row_f32"], 3)
    same = _as_device_matrix(col[:1], torch.device("cpu"), d)
    assert same.data_ptr() == col.data_ptr() and _layout_of(same) == (L["col_f32"], N)


def test_model_registry_holds_models_weakly():
    """the handle registry must not keep a dropped DeviceGP (and its device blob) alive."""
    import gc
    import weakref

    from baybe_b200 import engine

    class Dummy:
        pass

    assert isinstance(engine._registry, weakref.WeakValueDictionary)
    obj = Dummy()
    engine._registry[-1] = obj
    assert engine._registry.get(-1) is obj
    del obj
    gc.collect()
    assert engine._registry.get(-1) is None


def test_hybrid_search_by_scoring_finds_the_optimum_of_a_known_function():
    """baybe_b200.hybrid.HybridSearch with a stand-in scorer on the CPU: the sweep over
    (discrete configurations) x (shared Sobol points) plus the shrinking-box refinement must land on the maximiser of
    a smooth function whose optimum lies in the interior of one configuration's box."""
    import torch

    from baybe_b200.hybrid import HybridSearch

    disc = torch.tensor([[0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 1.0], [0.5, 0.5]], dtype=torch.float32)
    opt_c = torch.tensor([0.3141, 0.7182, 0.55])

    class Scorer:
        calls = 0

        def score(self, rows):
            Scorer.calls += 1
            dpart = -((rows[:, 0] - 1.0) ** 2 + (rows[:, 1] - 0.0) ** 2)  # configuration (1, 0) is best
            cpart = -((rows[:, 2:] - opt_c) ** 2).sum(-1) * (1.0 + rows[:, 0])
            return dpart + cpart

    lo, hi = torch.zeros(3), torch.ones(3)
    row, val = HybridSearch(n_sobol=256, n_seeds=16, n_local=64, n_rounds=8).best_point(Scorer(), disc, lo, hi, seed=3)
    assert torch.equal(row[:2], torch.tensor([1.0, 0.0]))
    assert float((row[2:] - opt_c).abs().max()) < 5e-3 and val > -1e-4
    assert Scorer.calls == 1 + 8  # one global sweep + one sweep per refinement round
    # deterministic for a seed
    row2, val2 = HybridSearch(n_sobol=256, n_seeds=16, n_local=64, n_rounds=8).best_point(Scorer(), disc, lo, hi, seed=3)
    assert torch.equal(row, row2) and val == val2


def test_qnei_spec_is_mirrored_and_kept_off_the_fused_kernels():
    from baybe_b200 import AcqConfig
    from baybe_b200.acquisition import convert_acqf, qNoisyExpectedImprovement

    spec = convert_acqf("qNEI")
    assert isinstance(spec, qNoisyExpectedImprovement) and spec.supports_batching and spec.prune_baseline
    cfg = AcqConfig(kind="qNEI", obj_scale=-1.0)
    assert cfg.is_mc
    with pytest.raises(NotImplementedError):
        cfg.to_c()


def test_default_preset_dispatches_on_search_space_content():
    """presets/baybe.py:151-197: Substance parameters -> Chen preset, everything else -> BayBE preset; task spaces
    select the leave-one-out criterion (presets/baybe.py:270-281)."""
    from types import SimpleNamespace

    from baybe_b200.kernels import gp_preset
    from baybe_b200.surrogates import default_preset_name

    class SubstanceParameter:  # the dispatch keys on the reference's class name
        pass

    class NumericalDiscreteParameter:
        pass

    plain = SimpleNamespace(parameters=(NumericalDiscreteParameter(), NumericalDiscreteParameter()))
    chem = SimpleNamespace(parameters=(NumericalDiscreteParameter(), SubstanceParameter()))
    assert default_preset_name(plain) == "BAYBE" and default_preset_name(chem) == "CHEN"
    assert default_preset_name(SimpleNamespace()) == "BAYBE"
    chen, baybe = gp_preset("CHEN", 12), gp_preset("BAYBE", 12)
    assert chen != baybe  # different kernel / prior / likelihood numbers
