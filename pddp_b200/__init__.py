"""pddp_b200 -- B200-native (sm_100a) iteration hot path of anassinator/pddp:
linearise -> backward Riccati -> rollout with parallel line search, behind pddp's controller /
model / cost API.  CUDA kernels live in pddp_b200/csrc and are reached through the C ABI declared
in include/pddp_b200.h; there is no CPU fallback."""
from . import controllers, costs, encoding, envs, examples, models  # noqa: F401
from .controllers import PDDPController, iLQRController, iLQRState  # noqa: F401
from .encoding import GaussianVariable, StateEncoding  # noqa: F401

__version__ = "0.1.0"
