"""BNN training loss.  The device trainer (csrc/bnn_train.cu) evaluates exactly this expression and its
gradient inside its forward/backward kernel; this host function is the same formula for callers that want the
number (plain torch ops on whatever device the tensors are on)."""
import math


def gaussian_log_likelihood(targets, pred_means, pred_stds=None):
    """log N(targets; pred_means, diag(pred_stds^2)) summed over the last dimension, with the reference's
    constant: ONE 0.5*log(2*pi) per row, not one per dimension.  ref: pddp/models/bnn/losses.py:20-38"""
    deltas = pred_means - targets
    if pred_stds is None:
        return -(deltas ** 2).sum(dim=-1) * 0.5
    return -((deltas / pred_stds) ** 2).sum(dim=-1) * 0.5 - pred_stds.log().sum(dim=-1) - math.log(2 * math.pi) * 0.5
