"""Example problems, namespaced like pddp.examples.<problem>.{<Problem>DynamicsModel, <Problem>Cost, <Problem>Env}."""
from types import SimpleNamespace

from .costs import CartpoleCost, DoubleCartpoleCost, PendulumCost, RendezvousCost
from .envs import CartpoleEnv, DoubleCartpoleEnv, PendulumEnv, RendezvousEnv
from .models import (CartpoleDynamicsModel, DoubleCartpoleDynamicsModel, PendulumDynamicsModel,
                     RendezvousDynamicsModel)

pendulum = SimpleNamespace(PendulumDynamicsModel=PendulumDynamicsModel, PendulumCost=PendulumCost, PendulumEnv=PendulumEnv)
cartpole = SimpleNamespace(CartpoleDynamicsModel=CartpoleDynamicsModel, CartpoleCost=CartpoleCost, CartpoleEnv=CartpoleEnv)
double_cartpole = SimpleNamespace(DoubleCartpoleDynamicsModel=DoubleCartpoleDynamicsModel,
                                  DoubleCartpoleCost=DoubleCartpoleCost, DoubleCartpoleEnv=DoubleCartpoleEnv)
rendezvous = SimpleNamespace(RendezvousDynamicsModel=RendezvousDynamicsModel, RendezvousCost=RendezvousCost,
                             RendezvousEnv=RendezvousEnv)
