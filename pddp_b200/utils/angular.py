"""Angular state helpers (host side).  The hot path augments inside the kernels: `augment_state` feeds the BNN
(csrc/bnn_common.cuh) and the closed-form moment matching of sin / cos, `augment_encoded_state`, is part of every
cost evaluation (csrc/core.cuh `cost_state`).  ref: pddp/utils/angular.py"""
import torch


def complementary_indices(indices, size):
    """ref: pddp/utils/angular.py:26-44"""
    keep = torch.ones(size, dtype=torch.bool)
    keep[torch.as_tensor(indices, dtype=torch.long)] = False
    return keep.nonzero().flatten()


def infer_augmented_state_size(angular_indices, non_angular_indices):
    """ref: pddp/utils/angular.py:329-340"""
    return len(non_angular_indices) + 2 * len(angular_indices)


def infer_reduced_state_size(angular_indices, non_angular_indices):
    """ref: pddp/utils/angular.py:343-353"""
    return len(non_angular_indices) + len(angular_indices)


def augment_state(x, angular_indices, non_angular_indices):
    """[x_nonang..., sin a1, cos a1, sin a2, ...]   ref: pddp/utils/angular.py:251-286"""
    ang = [int(i) for i in angular_indices]
    non = [int(i) for i in non_angular_indices]
    if not ang:
        return x
    a = x[..., ang]
    sc = torch.stack([a.sin(), a.cos()], -1).reshape(*x.shape[:-1], 2 * len(ang))
    return torch.cat([x[..., non], sc], -1)


def reduce_state(x_, angular_indices, non_angular_indices):
    """Inverse of augment_state (angles by atan2).  ref: pddp/utils/angular.py:289-326"""
    ang = [int(i) for i in angular_indices]
    non = [int(i) for i in non_angular_indices]
    if not ang:
        return x_
    x = torch.empty(*x_.shape[:-1], len(ang) + len(non), dtype=x_.dtype, device=x_.device)
    x[..., non] = x_[..., :len(non)]
    sc = x_[..., len(non):].reshape(*x_.shape[:-1], len(ang), 2)
    x[..., ang] = torch.atan2(sc[..., 0], sc[..., 1])
    return x
