"""Environments (namespaced like pddp.envs).  `KnownDynamicsEnv` stands where the reference has `GymEnv`: the
example environments' ground truth is their closed-form model (pddp/examples/*/env.py step()), stepped here by
`pddp_env_step_known` for a batch of instances on the device; there is no gym dependency and no rendering."""
from .base import Env, KnownDynamicsEnv

__all__ = ["Env", "KnownDynamicsEnv"]
