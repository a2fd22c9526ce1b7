"""PDDP controller (mirror of pddp/controllers/pddp.py): the trial loop around the iLQR hot path -- collect data
with the current controller on the environment, train the BNN on it (on the device: `pddp_bnn_train`), optimise the
controls on the learned model (`super().fit`, the GPU iteration), run the result as MPC, repeat."""
import torch

from ..utils.encoding import StateEncoding, decode_mean
from .ilqr import _trajectory_cost, iLQRController


class PDDPController(iLQRController):
    """ref: pddp/controllers/pddp.py:32-206.  The trial loop (collect data with the controller on the
    env, retrain the model) is host-side and unchanged in spirit; every `super().fit()` inside it
    is the GPU hot path."""

    def __init__(self, env, model, cost, model_opts={}, cost_opts={}, training_opts={}, **kwargs):
        super().__init__(env, model, cost, model_opts, cost_opts, **kwargs)
        self._training_opts = training_opts

    def fit(self, U, encoding=StateEncoding.DEFAULT, quiet=False, on_trial=None, max_trials=None,
            n_initial_sample_trajectories=2, sampling_noise=1.0, train_on_start=True, max_dataset_size=1000,
            resample_model=True, u_min=None, u_max=None, **kwargs):
        U = U.detach()
        dataset, total_trials = None, 0
        if train_on_start:
            for i in range(n_initial_sample_trajectories):
                self.env.reset()
                Ui = U if i == 0 else sampling_noise * torch.rand_like(U)
                if i > 0 and u_min is not None and u_max is not None:
                    Ui = (u_max - u_min) * Ui + u_min
                new_data, _ = _apply_controller(self.env, self.cost, Ui, U.shape[0], encoding, False, quiet,
                                                self._cost_opts, u_min=u_min, u_max=u_max)
                dataset = _concat_datasets(dataset, new_data, max_dataset_size)
                if callable(on_trial):
                    on_trial(total_trials, new_data[0], new_data[1])
                total_trials += 1
            self.model.train()
            self.model.fit(*dataset, quiet=quiet, **self._training_opts)
        while True:
            self.env.reset()
            self.model.eval()
            if resample_model and hasattr(self.model, "resample"):
                self.model.resample()
            Z, U, state = super().fit(U, encoding=encoding, quiet=quiet, u_min=u_min, u_max=u_max, **kwargs)
            if not self.training:
                break
            new_data, _ = _apply_controller(self.env, self.cost, self, 2 * U.shape[0], encoding, True, quiet,
                                            self._cost_opts, u_min=u_min, u_max=u_max, **kwargs)
            if callable(on_trial):
                on_trial(total_trials, new_data[0], new_data[1])
            dataset = _concat_datasets(dataset, new_data, max_dataset_size)
            self.model.train()
            self.model.fit(*dataset, quiet=quiet, **self._training_opts)
            total_trials += 1
            if max_trials is not None and total_trials >= max_trials:
                break
        return Z, U, state


def _apply_controller(env, cost, controller, H, encoding, mpc=False, quiet=False, cost_opts={}, **kwargs):
    """ref: pddp/controllers/pddp.py:209-247 -> ((X, U, dX), J): runs `controller` (a feedback / MPC
    controller, or a tensor of open-loop controls) on `env` for H steps and returns the trial's dataset
    and cost.  With a device environment (pddp_b200.envs) nothing leaves the GPU: the simulator step is
    `pddp_env_step_known`, every MPC step is one batched iteration of the hot path, and B instances
    ([B, nz] states, [B, nu] actions) run at once; the dataset is then [B*H, ...] and J is [B]."""
    Z, U = [], []
    open_loop = controller if isinstance(controller, torch.Tensor) else None
    dev = open_loop.device if open_loop is not None else controller._U_nominal.device
    for i in range(H):
        z = env.get_state().encode(encoding).to(dev)
        Z.append(z)
        if open_loop is not None:
            u = open_loop[:, i] if open_loop.dim() == 3 else open_loop[i]
        else:
            u = controller(z, i, encoding, mpc, **kwargs)
        U.append(u)
        env.apply(u)
    Z.append(env.get_state().encode(encoding).to(dev))
    Z, U = torch.stack(Z).detach(), torch.stack(U).detach()          # [H+1, (B,) nz], [H, (B,) nu]
    J = _trajectory_cost(cost, Z, U, encoding, cost_opts) if Z.is_cuda else None
    X = decode_mean(Z, encoding, getattr(env, "state_size", None))
    X, dX = X[:-1], X[1:] - X[:-1]
    if Z.dim() == 3:                                                 # instance-major rows, like B trials back to back
        X, U, dX = (t.transpose(0, 1).reshape(-1, t.shape[-1]) for t in (X, U, dX))
    return (X, U, dX), J


def _concat_datasets(first, second, max_dataset_size=None):
    """ref: pddp/controllers/pddp.py:250-267"""
    if first is None:
        return second
    out = tuple(torch.cat([a, b]) for a, b in zip(first, second))
    return tuple(t[-max_dataset_size:] for t in out) if max_dataset_size is not None else out
