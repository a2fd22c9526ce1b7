"""pddp_b200 -- B200-native (sm_100a) iteration hot path of anassinator/pddp:
linearise -> backward Riccati -> rollout with parallel line search, behind pddp's controller /
model / cost API and under pddp's own module paths (`import pddp_b200 as pddp`).  CUDA kernels live in
pddp_b200/csrc and are reached through the C ABI declared in include/pddp_b200.h; there is no CPU fallback."""
from . import utils  # noqa: F401  (first: everything else imports utils.encoding)
from . import controllers, costs, envs, examples, models  # noqa: F401
from .controllers import PDDPController, iLQRController, iLQRState  # noqa: F401
from .utils.encoding import StateEncoding  # noqa: F401
from .utils.gaussian_variable import GaussianVariable  # noqa: F401

__version__ = "0.2.0"
__all__ = ["controllers", "costs", "envs", "models", "utils", "GaussianVariable", "StateEncoding"]


def install_as(name="pddp"):
    """Registers this package and every submodule under another top-level name, so that a script written
    against the reference (`import pddp`, `from pddp.controllers.ilqr import *`, `pddp.utils.encoding...`)
    runs unchanged:  `import pddp_b200; pddp_b200.install_as("pddp")`  before its imports."""
    import sys
    for key, module in list(sys.modules.items()):
        if key == __name__ or key.startswith(__name__ + "."):
            sys.modules[name + key[len(__name__):]] = module
    return sys.modules[name]
