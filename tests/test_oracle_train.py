"""The oracle's restatement of BNN training (explicit forward / backward / Adam-amsgrad formulas) against what the
reference's own fit() produced: first-step loss and gradients, parameters after the last step."""
import pytest
import torch

import pddp_oracle as O
from train_util import augmented_inputs, load, train_tags


@pytest.mark.parametrize("tag", train_tags())
def test_training_restatement_matches_the_reference(tag):
    fx = load(tag)
    dt = fx["dtype"]
    tol = 1e-9 if dt == torch.float64 else 2e-4
    p, g0, losses = O.bnn_train(fx["p_init"], augmented_inputs(fx), fx["dX"], fx["batch_idx"], fx["noise"].to(dt),
                                [int(h) for h in fx["hidden"]], 4, int(fx["dropout"]), fx["lr"], fx["reg_scale"],
                                reg=fx["reg"], rate=fx["rate"], X_mean=fx["X_mean"], X_std_inv=fx["X_std_inv"],
                                dX_mean=fx["dX_mean"], dX_std=fx["dX_std"])
    assert abs(float(losses[0]) - fx["loss0"]) <= tol * max(1.0, abs(fx["loss0"]))
    assert float((g0 - fx["grads0"]).abs().max()) <= tol * max(1.0, float(fx["grads0"].abs().max()))
    # Adam divides by sqrt(v): early steps amplify rounding differences of tiny gradients, hence the looser bound
    assert float((p - fx["p_final"]).abs().max()) <= (1e-7 if dt == torch.float64 else 2e-4)
    assert float((p - fx["p_init"]).abs().max()) > 1e-3
