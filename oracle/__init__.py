"""CPU oracle for the recommend-time hot path (TEST INFRASTRUCTURE ONLY).

This package restates, in plain torch-CPU float64, the arithmetic that
``Campaign.recommend()`` executes inside BoTorch/GPyTorch for a purely discrete
search space.  It exists to check the CUDA
path; nothing in ``baybe_b200/`` may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` are allowed to call it.

PARITY UNPINNED: the reference delegates this arithmetic to botorch==0.16.1 /
gpytorch==1.14.3 / linear-operator==0.6 (``/root/reference/uv.lock``), none of
which is vendored under ``/root/reference`` or installable in this image, and the
reference's own tests hold no golden vectors for this path.
The oracle therefore *defines* the behaviour; it is cross-checked against
independent implementations where one exists (scikit-learn's GP posterior, scipy's
normal distribution, closed-form n=1 cases) in ``tests/test_oracle.py``.
"""

from oracle.reference_path import (  # noqa: F401
    AcqSpec,
    GPModel,
    KernelSpec,
    acq_values,
    acq_values_joint,
    acq_values_qnei,
    best_f_from_training,
    build_model,
    kernel_matrix,
    optimize_acqf_discrete,
    posterior,
    posterior_joint,
    sobol_normal_samples,
)
