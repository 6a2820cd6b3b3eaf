"""ORACLE (test infrastructure): torchfilter.filters restated from SURVEY.md Appendix A.3/A.5/A.6.

Upstream source absent (ref: setup.py:14) -- PARITY UNPINNED.  Reference call sites:
  * ``ParticleFilter(dynamics_model=, measurement_model=, num_particles=)``
        ref: crossmodal/push_models/pf.py:14-27, crossmodal/push_models/crossmodal_pf.py:18-40
  * ``VirtualSensorExtendedKalmanFilter(dynamics_model=, virtual_sensor_model=)``
        ref: crossmodal/door_models/kf.py:14-28
  * ``initialize_beliefs`` / ``forward_loop``  ref: crossmodal/eval_helpers.py:125-142

Oracle extension (not in upstream): ``ParticleFilter.noise`` -- when set to a noise source
(see ``oracle/noise.py``) every random draw is taken from it instead of
``torch.distributions``; with ``TorchRNGNoise`` the two paths are bit-identical on CPU
(tests/test_oracle_rng.py), which is what makes "identical uniform draws" meaningful.
"""
import math

import numpy as np
import torch
from fannypack.utils import SliceWrapper

from .base import (
    DynamicsModel,
    Filter,
    KalmanFilterBase,
    KalmanFilterMeasurementModel,
    ParticleFilterMeasurementModel,
    VirtualSensorModel,
)


def multinomial_inverse_cdf(probs: torch.Tensor, uniforms: torch.Tensor) -> torch.Tensor:
    """What ``torch.multinomial(probs, S, replacement=True)`` computes on CPU, with the uniforms
    made explicit (Appendix A.6): per row, a *sequential fp32* running sum, divided by its fp32
    total, then a lower-bound search of each float64 uniform.  Returns int64 (N, S)."""
    p = probs.detach().cpu().numpy().astype(np.float32)
    u = uniforms.detach().cpu().numpy().astype(np.float64)
    N, M = p.shape
    cdf = np.add.accumulate(p, axis=1, dtype=np.float32)
    cdf = (cdf / cdf[:, -1:]).astype(np.float32)
    out = np.empty(u.shape, dtype=np.int64)
    for n in range(N):
        out[n] = np.searchsorted(cdf[n].astype(np.float64), u[n], side="left")
    np.minimum(out, M - 1, out=out)
    return torch.from_numpy(out).to(probs.device)


class ParticleFilter(Filter):
    """Appendix A.3.  Belief = M weighted particles per trajectory."""

    def __init__(
        self,
        *,
        dynamics_model: DynamicsModel,
        measurement_model: ParticleFilterMeasurementModel,
        num_particles: int = 100,
        resample=None,
        soft_resample_alpha: float = 1.0,
        estimation_method: str = "weighted_average",
    ):
        super().__init__(state_dim=dynamics_model.state_dim)
        assert isinstance(dynamics_model, DynamicsModel)
        assert isinstance(measurement_model, ParticleFilterMeasurementModel)
        self.dynamics_model = dynamics_model
        self.measurement_model = measurement_model
        self.num_particles = num_particles
        self.resample = resample
        self.soft_resample_alpha = soft_resample_alpha
        self.estimation_method = estimation_method
        self.particle_states = None
        self.particle_log_weights = None
        self._initialized = False
        self.noise = None  # oracle extension

    # ---- belief initialisation --------------------------------------------------------------
    def initialize_beliefs(self, *, mean, covariance):
        N = mean.shape[0]
        sd, M = self.state_dim, self.num_particles
        assert mean.shape == (N, sd)
        assert covariance.shape == (N, sd, sd)
        if self.noise is None:
            draws = torch.distributions.MultivariateNormal(mean, covariance).sample((M,))
        else:
            eps = self.noise.init_eps(M, N, sd, like=mean)
            chol = torch.linalg.cholesky(covariance)
            draws = mean[None] + (chol[None] @ eps[..., None]).squeeze(-1)
        self.particle_states = draws.transpose(0, 1)
        assert self.particle_states.shape == (N, M, sd)
        self.particle_log_weights = self.particle_states.new_full((N, M), -math.log(M))
        self._initialized = True

    # ---- one filter step -----------------------------------------------------------------------
    def forward(self, *, observations, controls):
        assert self._initialized, "Particle filter not initialized: call initialize_beliefs() first"
        N, M, sd = self.particle_states.shape
        resample = self.resample if self.resample is not None else (not self.training)

        if not resample and self.num_particles != M:
            # particle count changed without resampling: tile, then fill with a random subset
            reps, extra = divmod(self.num_particles, M)
            idx = torch.arange(M, device=self.particle_states.device).repeat(reps)
            if extra:
                perm = torch.randperm(M) if self.noise is None else self.noise.randperm(M)
                idx = torch.cat([idx, perm[:extra].to(idx.device)])
            self.particle_states = self.particle_states[:, idx, :]
            logw = self.particle_log_weights[:, idx]
            self.particle_log_weights = logw - torch.logsumexp(logw, dim=1, keepdim=True)
            M = self.num_particles

        # predict: every particle through the dynamics, plus reparameterised process noise
        flat_states = self.particle_states.reshape(-1, sd)
        flat_controls = SliceWrapper(controls).map(
            lambda c: torch.repeat_interleave(c, repeats=M, dim=0)
        )
        pred, scale_trils = self.dynamics_model(initial_states=flat_states, controls=flat_controls)
        if self.noise is None:
            moved = torch.distributions.MultivariateNormal(loc=pred, scale_tril=scale_trils).rsample()
        else:
            eps = self.noise.process_eps(N * M, sd, like=pred)
            moved = pred + (scale_trils @ eps[..., None]).squeeze(-1)
        self.particle_states = moved.view(N, M, sd)

        # reweight + normalise
        logw = self.particle_log_weights + self.measurement_model(
            states=self.particle_states, observations=observations
        )
        assert logw.shape == (N, M)
        self.particle_log_weights = logw - torch.logsumexp(logw, dim=1, keepdim=True)

        # estimate (before resampling)
        if self.estimation_method == "weighted_average":
            estimate = torch.sum(
                torch.exp(self.particle_log_weights)[:, :, None] * self.particle_states, dim=1
            )
        elif self.estimation_method == "argmax":
            best = torch.argmax(self.particle_log_weights, dim=1)
            estimate = self.particle_states[torch.arange(N, device=best.device), best]
        else:
            raise AssertionError(f"unknown estimation method {self.estimation_method}")

        if resample:
            self._resample()
        return estimate

    def _resample(self):
        N, M, sd = self.particle_states.shape
        Mout = self.num_particles
        logw = self.particle_log_weights
        uniform = logw.new_full((N, Mout), -math.log(M))
        alpha = self.soft_resample_alpha
        if alpha < 1.0:
            assert Mout == M, "soft resampling keeps the particle count"
            logits = torch.logsumexp(
                torch.stack([logw + math.log(alpha), uniform + math.log(1.0 - alpha)], dim=0), dim=0
            )
            new_logw = logw - logits
        else:
            logits = logw
            new_logw = uniform

        if self.noise is None:
            idx = torch.distributions.Categorical(logits=logits).sample((Mout,)).T
        else:
            normalised = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
            probs = torch.softmax(normalised, dim=-1)
            idx = self.noise.resample_indices(probs, Mout, logits=logits)
        assert idx.shape == (N, Mout)

        self.particle_states = torch.gather(
            self.particle_states, 1, idx[:, :, None].expand(N, Mout, sd)
        )
        if alpha < 1.0:
            new_logw = torch.gather(new_logw, 1, idx)
        self.particle_log_weights = new_logw
        self.last_resample_indices = idx  # oracle extension: exposed for the parity tests


class ExtendedKalmanFilter(KalmanFilterBase):
    """Appendix A.5.  Plain-form covariance update, autograd Jacobians, ``torch.inverse``."""

    def _predict_step(self, *, controls):
        mean, cov = self._belief_mean, self._belief_covariance
        pred_mean, q_tril = self.dynamics_model(initial_states=mean, controls=controls)
        A = self.dynamics_model.jacobian(initial_states=mean, controls=controls)
        self._belief_mean = pred_mean
        self._belief_covariance = A @ cov @ A.transpose(-1, -2) + q_tril @ q_tril.transpose(-1, -2)

    def _update_step(self, *, observations):
        mean, cov = self._belief_mean, self._belief_covariance
        expected, r_tril = self.measurement_model(states=mean)
        C = self.measurement_model.jacobian(states=mean)
        Ct = C.transpose(-1, -2)
        S = C @ cov @ Ct + r_tril @ r_tril.transpose(-1, -2)
        gain = cov @ Ct @ torch.inverse(S)
        innovation = observations - expected
        self._belief_mean = mean + (gain @ innovation[:, :, None]).squeeze(-1)
        eye = torch.eye(self.state_dim, device=cov.device, dtype=cov.dtype)
        self._belief_covariance = (eye - gain @ C) @ cov


class _IdentityMeasurementModel(KalmanFilterMeasurementModel):
    """y_hat = x, C = I, noise tril installed per step by the virtual-sensor filter."""

    def __init__(self, *, state_dim: int):
        super().__init__(state_dim=state_dim, observation_dim=state_dim)
        self.scale_tril = None

    def forward(self, *, states):
        assert self.scale_tril is not None
        return states, self.scale_tril

    def jacobian(self, *, states):
        N = states.shape[0]
        eye = torch.eye(self.state_dim, device=states.device, dtype=states.dtype)
        return eye[None].expand(N, self.state_dim, self.state_dim)


class VirtualSensorExtendedKalmanFilter(ExtendedKalmanFilter):
    """EKF whose measurement is a learned 'virtual sensor' z(obs) observed through identity."""

    def __init__(self, *, dynamics_model: DynamicsModel, virtual_sensor_model: VirtualSensorModel):
        super().__init__(
            dynamics_model=dynamics_model,
            measurement_model=_IdentityMeasurementModel(state_dim=dynamics_model.state_dim),
        )
        self.virtual_sensor_model = virtual_sensor_model

    def forward(self, *, observations, controls):
        z, r_tril = self.virtual_sensor_model(observations=observations)
        self.measurement_model.scale_tril = r_tril
        return super().forward(observations=z, controls=controls)
