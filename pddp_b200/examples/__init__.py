"""Example problems, namespaced like pddp.examples (cartpole, double_cartpole, pendulum, rendezvous)."""
from . import cartpole, double_cartpole, pendulum, rendezvous
from .problems import SampleProblems

__all__ = ["SampleProblems", "cartpole", "double_cartpole", "pendulum", "rendezvous"]
