"""torchebm_b200: B200-native (sm_100a) MCMC negative-sampling hot path behind torchebm's sampler API.

Importing the package never touches the GPU; the CUDA library (lib/libebm_b200.so, built by
`python -m torchebm_b200.build`) is loaded on first use and its absence is an error, not a fallback.
"""

from . import _lib
from .core import (BaseModel, BaseScheduler, ConstantScheduler, CosineScheduler, DoubleWellModel,
                   ExponentialDecayScheduler, GaussianModel, HarmonicModel, LinearScheduler, MixtureOfGaussiansModel,
                   MLPEnergy, RastriginModel, energy_descriptor, mark_mlp_energy)
from .integrators import EulerMaruyamaIntegrator, HeunIntegrator, LeapfrogIntegrator, energy_drift
from .losses import BaseContrastiveDivergence, ContrastiveDivergence
from .samplers import BaseSampler, GradientDescentSampler, HamiltonianMonteCarlo, LangevinDynamics, NesterovSampler

__all__ = [
    "BaseModel", "BaseScheduler", "ConstantScheduler", "CosineScheduler", "DoubleWellModel", "ExponentialDecayScheduler",
    "GaussianModel", "HarmonicModel", "LinearScheduler", "MixtureOfGaussiansModel", "MLPEnergy", "RastriginModel",
    "energy_descriptor", "mark_mlp_energy", "EulerMaruyamaIntegrator", "HeunIntegrator", "LeapfrogIntegrator", "energy_drift",
    "BaseContrastiveDivergence", "ContrastiveDivergence", "BaseSampler", "HamiltonianMonteCarlo", "LangevinDynamics",
    "GradientDescentSampler", "NesterovSampler",
]
