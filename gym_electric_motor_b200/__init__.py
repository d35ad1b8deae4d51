"""gym_electric_motor_b200 — B200-native vectorised physical-system step for gym-electric-motor (GEM).

    import gym_electric_motor_b200 as gem
    env = gem.make("Cont-CC-PMSM-v0", num_envs=1 << 20, ode_solver=gem.physical_systems.RK4Solver())
    (state, ref), _ = env.reset(seed=0)
    (state, ref), reward, terminated, truncated, _ = env.step(actions)      # torch tensors on the GPU, one launch

The package mirrors the reference's surface for the hot path only: `make`, the batched
`ElectricMotorEnvironment`, the SCML component classes, reference generators, reward function and constraints.
All compute runs in libgemb200.so (hand-written sm_100a CUDA, C-ABI in include/gemb200.h); importing this package
does not need a GPU, creating an environment does.
"""
from . import envs, physical_system_wrappers, physical_systems, reference_generators, reward_functions, vector, visualization  # noqa: F401
from .constraints import Constraint, ConstraintMonitor, LimitConstraint, SquaredConstraint  # noqa: F401
from .core import Callback, ElectricMotorEnvironment, ElectricMotorVisualization  # noqa: F401
from .envs import env_ids, make  # noqa: F401
from .physical_systems import PhysicalSystem  # noqa: F401
from .reference_generators import ReferenceGenerator  # noqa: F401
from .reward_functions import RewardFunction, WeightedSumOfErrors  # noqa: F401

__version__ = "0.1.0"
