"""BNN training on the device (`pddp_bnn_train` behind `BNNDynamicsModel.fit`) against the
REFERENCE's own fit(): the fixtures hold the mini-batches and the dropout noise the reference's run consumed
(oracle/make_golden_train.py), so the comparison is step for step -- loss and every gradient of the first step,
and all parameters (weights, biases, the learned logit_p) after the last Adam(amsgrad) step."""
import pytest
import torch

from train_util import augmented_inputs, load, train_tags

pytestmark = pytest.mark.gpu


def _model(fx, P):
    from pddp_b200.models.bnn import BDropout, bnn_dynamics_model_factory
    hidden = [int(h) for h in fx["hidden"]]
    kwargs = {} if int(fx["dropout"]) == 0 else {"dropout_layers": BDropout, "initial_p": fx["rate"]}
    default = torch.get_default_dtype()
    torch.set_default_dtype(fx["dtype"])          # like the reference run: temperature = 0.1 in the run's own precision
    try:
        model = bnn_dynamics_model_factory(4, 1, hidden, [2], [0, 1, 3], **kwargs)(n_particles=10)
    finally:
        torch.set_default_dtype(default)
    model._store_flat(fx["p_init"])
    return model


@pytest.mark.parametrize("tag", train_tags())
def test_training_matches_the_reference(tag):
    import pddp_b200 as P
    import pddp_oracle as O
    fx = load(tag)
    dt = fx["dtype"]
    tol = 1e-9 if dt == torch.float64 else 2e-4
    args = dict(batch_size=int(fx["batch_size"]), reg_scale=fx["reg_scale"], learning_rate=fx["lr"], quiet=True,
                return_diagnostics=True)
    # first step: loss and every gradient
    model = _model(fx, P)
    model.train()
    loss, grads = model.fit(fx["X"], fx["U"], fx["dX"], n_iter=1, batch_indices=fx["batch_idx"][:1], noise=fx["noise"][:1],
                            **args)
    assert torch.allclose(model.X_mean.cpu(), fx["X_mean"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(model.dX_std.cpu(), fx["dX_std"], rtol=1e-6, atol=1e-7)
    assert abs(float(loss[0]) - fx["loss0"]) <= tol * max(1.0, abs(fx["loss0"]))
    g = grads.cpu()
    assert float((g - fx["grads0"]).abs().max()) <= tol * max(1.0, float(fx["grads0"].abs().max()))
    # the whole run: parameters after the last step
    model = _model(fx, P)
    loss, _ = model.fit(fx["X"], fx["U"], fx["dX"], n_iter=int(fx["n_iter"]), batch_indices=fx["batch_idx"],
                        noise=fx["noise"], **args)
    p = model.flat_parameters().cpu()
    assert float((p - fx["p_final"]).abs().max()) <= (1e-7 if dt == torch.float64 else 2e-4)
    assert float((p - fx["p_init"]).abs().max()) > 1e-3
    # and step by step against the oracle's restatement (same formulas on the CPU)
    po, _, lo = O.bnn_train(fx["p_init"], augmented_inputs(fx), fx["dX"], fx["batch_idx"], fx["noise"].to(dt),
                            [int(h) for h in fx["hidden"]], 4, int(fx["dropout"]), fx["lr"], fx["reg_scale"], reg=fx["reg"],
                            rate=fx["rate"], X_mean=fx["X_mean"], X_std_inv=fx["X_std_inv"], dX_mean=fx["dX_mean"],
                            dX_std=fx["dX_std"])
    assert float((loss.cpu() - lo).abs().max()) <= (1e-8 if dt == torch.float64 else 1e-3) * max(1.0, float(lo.abs().max()))
    # eval-mode state after training: masks are redrawn at the next use (modules.py:565-573)
    assert all(d.mask is None or d.mask.dim() < 2 or d.mask.shape[0] == 1 for d in (model.model.drop_0, model.model.drop_1))


def test_training_with_the_device_generator_learns():
    """No recorded noise: masks come from the kernel's counter-based generator.  A learnable synthetic target
    (dX = linear map of the augmented state and the action) must be fitted: the loss falls and the mean
    prediction error drops well below the target's spread."""
    from pddp_b200.models.bnn import bnn_dynamics_model_factory
    g = torch.Generator().manual_seed(0)
    n = 600
    X = torch.randn(n, 4, generator=g)
    U = torch.randn(n, 1, generator=g)
    A = 0.3 * torch.randn(6, 4, generator=g)
    feats = torch.cat([X[:, [0, 1, 3]], X[:, 2:3].sin(), X[:, 2:3].cos(), U], -1)
    dX = feats @ A + 0.01 * torch.randn(n, 4, generator=g)
    model = bnn_dynamics_model_factory(4, 1, [200, 200], [2], [0, 1, 3])(n_particles=50)
    lp_before = float(model.model.drop_0.logit_p)
    loss, _ = model.fit(X.cuda(), U.cuda(), dX.cuda(), n_iter=1500, learning_rate=1e-3, quiet=True, seed=3,
                        return_diagnostics=True)
    loss = loss.cpu()
    assert bool(torch.isfinite(loss).all())
    assert float(loss[-100:].mean()) < float(loss[:20].mean()) - 1.0
    assert float(model.model.drop_0.logit_p) != lp_before                      # concrete dropout rate is learned
    assert model.X_mean.shape == (6,) and model.model.fc_0.weight.device.type == "cpu"   # parameters written back in place
    # same seed, same run -> bit-identical parameters (counter-based noise, no atomics on the gradient path)
    a = bnn_dynamics_model_factory(4, 1, [64, 64], [2], [0, 1, 3])(n_particles=8)
    b = bnn_dynamics_model_factory(4, 1, [64, 64], [2], [0, 1, 3])(n_particles=8)
    b._store_flat(a.flat_parameters())
    for m in (a, b):
        m.fit(X.cuda(), U.cuda(), dX.cuda(), n_iter=50, quiet=True, seed=11)
    assert torch.equal(a.flat_parameters(), b.flat_parameters())
