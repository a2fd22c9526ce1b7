"""Example problems, namespaced like pddp.examples.<problem>.{<Problem>DynamicsModel, <Problem>Cost}."""
from types import SimpleNamespace

from .costs import CartpoleCost, DoubleCartpoleCost, PendulumCost
from .models import CartpoleDynamicsModel, DoubleCartpoleDynamicsModel, PendulumDynamicsModel

pendulum = SimpleNamespace(PendulumDynamicsModel=PendulumDynamicsModel, PendulumCost=PendulumCost)
cartpole = SimpleNamespace(CartpoleDynamicsModel=CartpoleDynamicsModel, CartpoleCost=CartpoleCost)
double_cartpole = SimpleNamespace(DoubleCartpoleDynamicsModel=DoubleCartpoleDynamicsModel,
                                  DoubleCartpoleCost=DoubleCartpoleCost)
