"""torchebm_b200: B200-native (sm_100a) MCMC negative-sampling hot path behind torchebm's sampler API.

Importing the package never touches the GPU; the CUDA library (lib/libebm_b200.so, built by
`python -m torchebm_b200.build`) is loaded on first use and its absence is an error, not a fallback.

Two class families share the fused bodies (samplers.py / losses.py mixins):
  * when the reference package `torchebm` is importable, the names below are subclasses of the reference's own classes
    (`dropin.py`) -- `issubclass(torchebm_b200.LangevinDynamics, torchebm.core.BaseSampler)` -- and the schedulers ARE the
    reference's; requests outside the fused kernels' scope run the reference's own code;
  * otherwise (or with EBM_B200_STANDALONE=1) they are this package's standalone mirrors of that API surface.
`torchebm_b200.REFERENCE_DERIVED` says which.
"""

from . import _lib
from ._ref import reference as _reference
from .core import (BaseModel, DoubleWellModel, GaussianModel, HarmonicModel, MixtureOfGaussiansModel, MLPEnergy,
                   RastriginModel, energy_descriptor, mark_mlp_energy)
from .diagnostics import ess_from_chain, ess_from_diagnostics
from .integrators import HeunIntegrator, energy_drift

REFERENCE_DERIVED = _reference() is not None

if REFERENCE_DERIVED:
    from torchebm.core import (BaseSampler, BaseScheduler, ConstantScheduler, CosineScheduler, ExponentialDecayScheduler,
                               LinearScheduler)
    from torchebm.core import BaseContrastiveDivergence
    from torchebm.integrators import HeunIntegrator  # noqa: F811  (no fused step at the integrator level: the reference's own)

    from .dropin import (ContrastiveDivergence, EulerMaruyamaIntegrator, GradientDescentSampler, HamiltonianMonteCarlo,
                         LangevinDynamics, LeapfrogIntegrator, NesterovSampler, install, uninstall)
else:
    from .core import BaseScheduler, ConstantScheduler, CosineScheduler, ExponentialDecayScheduler, LinearScheduler
    from .integrators import EulerMaruyamaIntegrator, LeapfrogIntegrator
    from .losses import BaseContrastiveDivergence, ContrastiveDivergence
    from .samplers import BaseSampler, GradientDescentSampler, HamiltonianMonteCarlo, LangevinDynamics, NesterovSampler

    def install() -> None:
        raise ImportError("torchebm_b200.install() needs the reference package `torchebm` on sys.path")

    uninstall = install

__all__ = [
    "BaseModel", "BaseScheduler", "ConstantScheduler", "CosineScheduler", "DoubleWellModel", "ExponentialDecayScheduler",
    "GaussianModel", "HarmonicModel", "LinearScheduler", "MixtureOfGaussiansModel", "MLPEnergy", "RastriginModel",
    "energy_descriptor", "mark_mlp_energy", "EulerMaruyamaIntegrator", "HeunIntegrator", "LeapfrogIntegrator", "energy_drift",
    "BaseContrastiveDivergence", "ContrastiveDivergence", "BaseSampler", "HamiltonianMonteCarlo", "LangevinDynamics",
    "GradientDescentSampler", "NesterovSampler", "REFERENCE_DERIVED", "install", "uninstall", "ess_from_chain",
    "ess_from_diagnostics",
]
