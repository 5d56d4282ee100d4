"""Prior distributions with the reference's names and parameterisations
(``/root/reference/baybe/priors/basic.py:17-110``); each provides the log density (up to its
normalising constant, which does not move the MAP optimum) and its derivative in float64 numpy --
what the hyper-parameter fit adds to the marginal likelihood evaluated on the GPU."""
from __future__ import annotations

import math

import numpy as np
from attrs import define, field
from attrs.validators import gt


class Prior:
    """Base class (priors/base.py)."""

    def log_prob(self, x: np.ndarray) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError

    def grad(self, x: np.ndarray) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError

    @property
    def mode(self) -> float | None:
        return None


@define(frozen=True)
class GammaPrior(Prior):
    concentration: float = field(converter=float, validator=gt(0.0))
    rate: float = field(converter=float, validator=gt(0.0))

    def log_prob(self, x):
        return (self.concentration - 1.0) * np.log(x) - self.rate * x

    def grad(self, x):
        return (self.concentration - 1.0) / x - self.rate

    @property
    def mode(self):
        return (self.concentration - 1.0) / self.rate if self.concentration > 1.0 else None


@define(frozen=True)
class LogNormalPrior(Prior):
    loc: float = field(converter=float)
    scale: float = field(converter=float, validator=gt(0.0))

    def log_prob(self, x):
        lx = np.log(x)
        return -lx - 0.5 * ((lx - self.loc) / self.scale) ** 2

    def grad(self, x):
        return -(1.0 + (np.log(x) - self.loc) / self.scale**2) / x

    @property
    def mode(self):
        return math.exp(self.loc - self.scale**2)


@define(frozen=True)
class NormalPrior(Prior):
    loc: float = field(converter=float)
    scale: float = field(converter=float, validator=gt(0.0))

    def log_prob(self, x):
        return -0.5 * ((x - self.loc) / self.scale) ** 2

    def grad(self, x):
        return -(x - self.loc) / self.scale**2

    @property
    def mode(self):
        return self.loc


@define(frozen=True)
class HalfNormalPrior(Prior):
    scale: float = field(converter=float, validator=gt(0.0))

    def log_prob(self, x):
        return -0.5 * (x / self.scale) ** 2

    def grad(self, x):
        return -x / self.scale**2


@define(frozen=True)
class HalfCauchyPrior(Prior):
    scale: float = field(converter=float, validator=gt(0.0))

    def log_prob(self, x):
        return -np.log1p((x / self.scale) ** 2)

    def grad(self, x):
        return -2.0 * x / (self.scale**2 + x**2)


@define(frozen=True)
class SmoothedBoxPrior(Prior):
    """gpytorch's SmoothedBoxPrior: flat on [a, b], Gaussian shoulders of width sigma outside."""

    a: float = field(converter=float)
    b: float = field(converter=float)
    sigma: float = field(converter=float, default=0.01, validator=gt(0.0))

    def __attrs_post_init__(self):
        if self.b <= self.a:
            raise ValueError(f"For {type(self).__name__}, the upper bound `b` (provided: {self.b}) "
                             f"must be larger than the lower bound `a` (provided: {self.a}).")

    def _excess(self, x):
        c, r = 0.5 * (self.a + self.b), 0.5 * (self.b - self.a)
        return np.clip(np.abs(x - c) - r, 0.0, None), np.sign(x - c)

    def log_prob(self, x):
        e, _ = self._excess(np.asarray(x, dtype=np.float64))
        return -0.5 * (e / self.sigma) ** 2

    def grad(self, x):
        e, sgn = self._excess(np.asarray(x, dtype=np.float64))
        return -sgn * e / self.sigma**2
