"""Known-answer pins the reference's own tests hold for this path, replayed on the oracle.

ref: tests/utils/test_encoding.py:18-19,160 ; tests/costs/test_quadratic.py:41-52 ;
     tests/utils/test_angular.py:64-68 ; tests/utils/test_autodiff.py:26-36."""
import pytest
import torch

import pddp_oracle as O

ENCODINGS = [O.FULL_COVARIANCE_MATRIX, O.UPPER_TRIANGULAR_CHOLESKY, O.VARIANCE_ONLY,
             O.STANDARD_DEVIATION_ONLY, O.IGNORE_UNCERTAINTY]


def random_gaussian(D, seed=0):
    g = torch.Generator().manual_seed(seed)
    m = torch.randn(D, generator=g, dtype=torch.float64)
    A = torch.randn(D, D, generator=g, dtype=torch.float64)
    return m, A @ A.T + 0.1 * torch.eye(D, dtype=torch.float64)


def test_encoded_sizes():
    assert [O.encoded_size(5, e) for e in ENCODINGS] == [30, 20, 10, 10, 5]
    for e in ENCODINGS:
        assert O.state_size(O.encoded_size(5, e), e) == 5


@pytest.mark.parametrize("enc", ENCODINGS)
def test_encode_decode_round_trip(enc):
    m, C = random_gaussian(5)
    z = O.encode(m, C=C, enc=enc)
    assert z.shape[-1] == O.encoded_size(5, enc)
    assert torch.allclose(O.decode_mean(z, enc), m)
    if enc in (O.FULL_COVARIANCE_MATRIX, O.UPPER_TRIANGULAR_CHOLESKY):
        assert torch.allclose(O.decode_covar(z, enc), C, atol=1e-9)
        U = O.decode_covar_sqrt(z, enc)
        assert torch.allclose(U.T @ U, C, atol=1e-9)
        assert torch.allclose(U, torch.triu(U))
    elif enc == O.IGNORE_UNCERTAINTY:
        assert torch.allclose(O.decode_covar(z, enc), 1e-6 * torch.eye(5, dtype=torch.float64))
    else:
        assert torch.allclose(O.decode_var(z, enc), torch.diagonal(C), atol=1e-9)


@pytest.mark.parametrize("enc", ENCODINGS)
def test_qrcost_hessian_is_Q_plus_Qt(enc):
    D, nu = 4, 2
    g = torch.Generator().manual_seed(1)
    Q = torch.randn(D, D, generator=g, dtype=torch.float64)
    R = torch.randn(nu, nu, generator=g, dtype=torch.float64)
    spec = O.QRCostSpec(Q, R, Q, torch.zeros(D, dtype=torch.float64),
                        torch.zeros(nu, dtype=torch.float64), D)
    m, C = random_gaussian(D, 2)
    z = O.encode(m, C=C, enc=enc)
    u = torch.randn(nu, generator=g, dtype=torch.float64)
    l, l_z, l_u, l_zz, l_uz, l_uu = O.cost_derivatives(spec, z, u, False, enc)
    assert torch.allclose(l_zz[:D, :D], Q + Q.T, atol=1e-9)
    assert torch.allclose(l_uu, R + R.T, atol=1e-9)
    assert torch.allclose(l_uz, torch.zeros(nu, z.shape[0], dtype=torch.float64))


@pytest.mark.parametrize("enc", ENCODINGS[:4])
def test_zero_variance_augmentation_matches_plain(enc):
    m = torch.tensor([0.3, -1.2, 2.0, 0.7], dtype=torch.float64)
    z = O.encode(m, V=torch.zeros(4, dtype=torch.float64) + (1e-300 if enc == 1 else 0.0), enc=enc)
    za = O.augment_encoded_state(z, (2,), (0, 1, 3), enc, 4)
    assert torch.allclose(O.decode_mean(za, enc, 5), O.augment_state(m, (2,), (0, 1, 3)),
                          atol=1e-6)


def test_identity_backprop_jacobian_of_polynomial():
    """The replicate-rows + identity-cotangent trick gives [1, 2x, 3x^2] exactly."""
    x = torch.full((3, 1), 2.0, dtype=torch.float64, requires_grad=True)
    y = torch.stack([x[0, 0], x[1, 0] ** 2, x[2, 0] ** 3])
    J, = torch.autograd.grad(y, x, torch.ones(3, dtype=torch.float64))
    assert J.flatten().tolist() == [1.0, 4.0, 12.0]
