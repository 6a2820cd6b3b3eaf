"""Known-answer tests for the restated EKF / PF: a linear-Gaussian
system has a closed-form Kalman filter; forward-mode vs autograd Jacobians."""
import numpy as np
import torch

from oracle import crossmodal_port as port  # installs the oracle shims on sys.path

import torchfilter  # noqa: E402  (the oracle shim)
from multimodalfilter_b200.synthetic import fill_parameters


class _LinearDynamics(torchfilter.base.DynamicsModel):
    def __init__(self, A, B, Q_tril):
        super().__init__(state_dim=A.shape[0])
        self.A, self.B, self.Q_tril = A, B, Q_tril

    def forward(self, *, initial_states, controls):
        N = initial_states.shape[0]
        nxt = initial_states @ self.A.T + controls @ self.B.T
        return nxt, self.Q_tril[None].expand(N, *self.Q_tril.shape)


class _DirectSensor(torchfilter.base.VirtualSensorModel):
    def __init__(self, R_tril):
        super().__init__(state_dim=R_tril.shape[0])
        self.R_tril = R_tril

    def forward(self, *, observations):
        N = observations.shape[0]
        return observations, self.R_tril[None].expand(N, *self.R_tril.shape)


def _closed_form_kf(A, B, Q, R, mean, cov, ys, us):
    out = []
    for y, u in zip(ys, us):
        mean = A @ mean + B @ u
        cov = A @ cov @ A.T + Q
        S = cov + R
        K = cov @ np.linalg.inv(S)
        mean = mean + K @ (y - mean)
        cov = (np.eye(len(mean)) - K) @ cov
        out.append(mean.copy())
    return np.stack(out), cov


def test_ekf_equals_closed_form_kalman_filter_fp64():
    torch.manual_seed(0)
    sd, cd, T, N = 3, 2, 12, 4
    A = torch.eye(sd, dtype=torch.float64) + 0.1 * torch.randn(sd, sd, dtype=torch.float64)
    B = torch.randn(sd, cd, dtype=torch.float64)
    Q_tril = torch.diag(torch.tensor([0.3, 0.2, 0.1], dtype=torch.float64))
    R_tril = torch.tril(0.2 * torch.randn(sd, sd, dtype=torch.float64)) + 0.5 * torch.eye(sd, dtype=torch.float64)
    ekf = torchfilter.filters.VirtualSensorExtendedKalmanFilter(
        dynamics_model=_LinearDynamics(A, B, Q_tril), virtual_sensor_model=_DirectSensor(R_tril)
    )
    mean0 = torch.randn(N, sd, dtype=torch.float64)
    cov0 = (torch.eye(sd, dtype=torch.float64) * 0.1)[None].expand(N, sd, sd)
    ys = torch.randn(T, N, sd, dtype=torch.float64)
    us = torch.randn(T, N, cd, dtype=torch.float64)
    ekf.initialize_beliefs(mean=mean0, covariance=cov0)
    est = ekf.forward_loop(observations=ys, controls=us)
    for n in range(N):
        ref, ref_cov = _closed_form_kf(
            A.numpy(), B.numpy(), (Q_tril @ Q_tril.T).numpy(), (R_tril @ R_tril.T).numpy(),
            mean0[n].numpy(), cov0[n].numpy(), ys[:, n].numpy(), us[:, n].numpy(),
        )
        np.testing.assert_allclose(est[:, n].detach().numpy(), ref, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(ekf.belief_covariance[n].detach().numpy(), ref_cov, rtol=1e-10, atol=1e-12)


def test_autograd_jacobian_matches_functional_jacobian():
    dyn = fill_parameters(port.DoorDynamicsModel(), seed=4).double()
    x = torch.randn(5, 3, dtype=torch.float64)
    u = torch.randn(5, 7, dtype=torch.float64)
    J = dyn.jacobian(initial_states=x, controls=u)
    for n in range(5):
        ref = torch.autograd.functional.jacobian(
            lambda s: dyn(initial_states=s[None], controls=u[n : n + 1])[0][0], x[n]
        )
        np.testing.assert_allclose(J[n].detach().numpy(), ref.numpy(), rtol=1e-10, atol=1e-12)


def test_particle_filter_mean_approaches_kalman_mean():
    """PF with many particles on the same linear-Gaussian system tracks the KF mean."""

    class _GaussianLikelihood(torchfilter.base.ParticleFilterMeasurementModel):
        def __init__(self, R):
            super().__init__(state_dim=R.shape[0])
            self.Rinv = torch.inverse(R)

        def forward(self, *, states, observations):
            d = observations[:, None, :] - states
            return -0.5 * torch.einsum("nmi,ij,nmj->nm", d, self.Rinv, d)

    torch.manual_seed(1)
    sd, cd, T, N, M = 2, 2, 8, 3, 20000
    A = torch.tensor([[1.0, 0.1], [0.0, 0.9]])
    B = 0.1 * torch.randn(sd, cd)
    Q_tril = torch.diag(torch.tensor([0.2, 0.2]))
    R_tril = torch.diag(torch.tensor([0.4, 0.3]))
    dyn = _LinearDynamics(A, B, Q_tril)
    pf = torchfilter.filters.ParticleFilter(
        dynamics_model=dyn, measurement_model=_GaussianLikelihood(R_tril @ R_tril.T), num_particles=M
    ).eval()
    ekf = torchfilter.filters.VirtualSensorExtendedKalmanFilter(
        dynamics_model=dyn, virtual_sensor_model=_DirectSensor(R_tril)
    )
    mean0 = torch.randn(N, sd)
    cov0 = (torch.eye(sd) * 0.1)[None].expand(N, sd, sd)
    us = torch.randn(T, N, cd)
    x, ys = mean0.clone(), []
    for t in range(T):  # observations simulated from the model itself (no outliers => no degeneracy)
        x = x @ A.T + us[t] @ B.T + torch.randn(N, sd) @ Q_tril.T
        ys.append(x + torch.randn(N, sd) @ R_tril.T)
    ys = torch.stack(ys)
    with torch.no_grad():
        pf.initialize_beliefs(mean=mean0, covariance=cov0)
        a = pf.forward_loop(observations=ys, controls=us)
    ekf.initialize_beliefs(mean=mean0, covariance=cov0)
    b = ekf.forward_loop(observations=ys, controls=us).detach()
    assert (a - b).abs().max() < 0.05
