"""Kernel specifications with the reference's class names and fields
(``/root/reference/baybe/kernels/basic.py:48-190``, ``composite.py:21-60``) and the GP presets built
from them (``surrogates/gaussian_process/presets/{baybe,chen,edbo}.py``).  A specification only
selects what the engine's kernels implement -- stationary ARD Matern-1/2, 3/2, 5/2 or RBF, optionally
wrapped in one ScaleKernel -- plus the hyper-priors and start values of the MAP fit."""
from __future__ import annotations

import math
from fractions import Fraction

from attrs import define, field
from attrs.validators import in_, instance_of
from attrs.validators import optional as optional_v

from baybe_b200.priors import GammaPrior, Prior

__all__ = ["Kernel", "MaternKernel", "RBFKernel", "ScaleKernel", "GPConfig", "gp_preset", "resolve_kernel"]


class Kernel:
    """Base class (kernels/base.py)."""


def _nu(value) -> float:
    return float(Fraction(value)) if isinstance(value, str) else float(value)


@define(frozen=True)
class MaternKernel(Kernel):
    nu: float = field(converter=_nu, validator=in_([0.5, 1.5, 2.5]), default=2.5)
    lengthscale_prior: Prior | None = field(default=None, validator=optional_v(instance_of(Prior)))
    lengthscale_initial_value: float | None = field(default=None)

    @property
    def family(self) -> str:
        return {0.5: "matern12", 1.5: "matern32", 2.5: "matern52"}[self.nu]


@define(frozen=True)
class RBFKernel(Kernel):
    lengthscale_prior: Prior | None = field(default=None, validator=optional_v(instance_of(Prior)))
    lengthscale_initial_value: float | None = field(default=None)

    @property
    def family(self) -> str:
        return "rbf"


@define(frozen=True)
class ScaleKernel(Kernel):
    base_kernel: Kernel = field(validator=instance_of(Kernel))
    outputscale_prior: Prior | None = field(default=None, validator=optional_v(instance_of(Prior)))
    outputscale_initial_value: float | None = field(default=None)
    outputscale_trainable: bool = field(default=True, validator=instance_of(bool))

    def __attrs_post_init__(self):
        if not isinstance(self.base_kernel, (MaternKernel, RBFKernel)):
            raise NotImplementedError("the B200 engine supports ScaleKernel over one stationary ARD kernel "
                                      f"(Matern or RBF), got {type(self.base_kernel).__name__}")


_SOFTPLUS0 = math.log(2.0)  # gpytorch's default start value of positive parameters


@define(frozen=True)
class GPConfig:
    """Everything the MAP fit needs besides the data."""

    family: str
    lengthscale_prior: Prior | None
    lengthscale_initial_value: float
    lengthscale_lower: float
    outputscale: bool
    outputscale_prior: Prior | None
    outputscale_initial_value: float
    outputscale_trainable: bool
    noise_prior: Prior | None
    noise_initial_value: float
    noise_lower: float = 1e-4  # botorch MIN_INFERRED_NOISE_LEVEL / gpytorch's default noise constraint


def resolve_kernel(kernel: Kernel, noise_prior: Prior | None = None, noise_initial_value: float | None = None,
                   lengthscale_lower: float = 1e-6) -> GPConfig:
    """GPConfig of a user-given kernel object (GaussianProcessSurrogate(kernel_or_factory=<Kernel>),
    surrogates/gaussian_process/core.py:147-186)."""
    scale = kernel if isinstance(kernel, ScaleKernel) else None
    base = scale.base_kernel if scale is not None else kernel
    if not isinstance(base, (MaternKernel, RBFKernel)):
        raise NotImplementedError(f"kernel {type(kernel).__name__} is not supported by the B200 engine")
    lp = base.lengthscale_prior
    ls0 = base.lengthscale_initial_value
    if ls0 is None:
        ls0 = lp.mode if lp is not None and lp.mode is not None else _SOFTPLUS0
    os0 = _SOFTPLUS0
    if scale is not None:
        os0 = scale.outputscale_initial_value
        if os0 is None:
            op = scale.outputscale_prior
            os0 = op.mode if op is not None and op.mode is not None else _SOFTPLUS0
    nz0 = noise_initial_value
    if nz0 is None:
        nz0 = noise_prior.mode if noise_prior is not None and noise_prior.mode is not None else _SOFTPLUS0
    return GPConfig(base.family, lp, float(ls0), lengthscale_lower, scale is not None,
                    None if scale is None else scale.outputscale_prior, float(os0),
                    True if scale is None else scale.outputscale_trainable, noise_prior, float(nz0))


def gp_preset(name: str, n_dims: int) -> GPConfig:
    """Presets by name; `n_dims` is the number of active (non-task) comp-rep columns."""
    key = name.upper()
    if key == "BAYBE":  # presets/baybe.py:57-144 (the default)
        lp = GammaPrior(3.0, 2.0 / math.exp(math.sqrt(2.0) - 3.0) / math.sqrt(n_dims))
        npri = GammaPrior(2.0, 1.0 / math.exp(-5.0))
        return GPConfig("matern52", lp, lp.mode, 2.5e-2, False, None, 1.0, False, npri, max(npri.mode, 1e-4))
    if key == "CHEN":  # presets/chen.py:35-61
        ls = 0.4 * math.sqrt(n_dims) + 4.0
        kern = ScaleKernel(MaternKernel(2.5, GammaPrior(2.0 * ls, 2.0), ls), GammaPrior(1.0 * ls, 1.0), ls)
        return resolve_kernel(kern, None, None)  # plain GaussianLikelihood(): no noise prior (chen.py:76-77)
    if key == "EDBO":  # presets/edbo.py (non-substance branches)
        if n_dims < 5:
            kern = ScaleKernel(MaternKernel(2.5, GammaPrior(1.2, 1.1), 0.2), GammaPrior(5.0, 0.5), 8.0)
            return resolve_kernel(kern, GammaPrior(1.05, 0.5), 0.1)
        kern = ScaleKernel(MaternKernel(2.5, GammaPrior(3.0, 1.0), 2.0), GammaPrior(5.0, 0.2), 20.0)
        return resolve_kernel(kern, GammaPrior(1.5, 0.1), 5.0)
    raise ValueError(f"unknown GP preset {name!r} (available: BAYBE, CHEN, EDBO)")
