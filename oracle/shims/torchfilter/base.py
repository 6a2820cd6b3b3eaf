"""ORACLE (test infrastructure): torchfilter.base restated from SURVEY.md Appendix A.1/A.2/A.4.

Upstream ``torchfilter`` is an un-vendored, un-pinned dependency of the reference
(ref: setup.py:12-15) -- PARITY UNPINNED.  Call sites that constrain the behaviour:
  * constructors ``super().__init__(state_dim=...)``      ref: crossmodal/push_models/dynamics.py:14,
                                                          crossmodal/base_models/crossmodal_pf.py:49
  * ``Filter.forward_loop(observations=, controls=)``     ref: crossmodal/eval_helpers.py:139-142
  * ``DynamicsModel.forward_loop(initial_states=, controls=)`` ref: crossmodal/eval_helpers.py:135-137
  * ``_belief_covariance`` read                            ref: crossmodal/base_models/crossmodal_kf.py:180
"""
import abc

import torch
import torch.nn as nn
from fannypack.utils import SliceWrapper


class Filter(nn.Module, abc.ABC):
    """State estimator: (observation, control) stream -> state estimates."""

    def __init__(self, *, state_dim: int):
        super().__init__()
        self.state_dim = state_dim

    @abc.abstractmethod
    def initialize_beliefs(self, *, mean, covariance):
        """mean (N, sd); covariance (N, sd, sd)."""

    @abc.abstractmethod
    def forward(self, *, observations, controls):
        """One filter step; returns the (N, sd) estimate."""

    def forward_loop(self, *, observations, controls):
        """Appendix A.2: serial loop over the leading time axis, going through ``__call__``."""
        obs = SliceWrapper(observations)
        ctrl = SliceWrapper(controls)
        T, N = ctrl.shape[:2]
        assert obs.shape[:2] == (T, N)

        first = self(observations=obs[0], controls=ctrl[0])
        assert first.shape == (N, self.state_dim)
        out = first.new_zeros((T, N, self.state_dim))
        out[0] = first
        for t in range(1, T):
            out[t] = self(observations=obs[t], controls=ctrl[t])
        return out


class DynamicsModel(nn.Module, abc.ABC):
    """x_t ~ N(f(x_{t-1}, u_t), L L^T); ``forward`` returns (f, L)."""

    def __init__(self, *, state_dim: int):
        super().__init__()
        self.state_dim = state_dim

    @abc.abstractmethod
    def forward(self, *, initial_states, controls):
        """(N, sd), controls -> ((N, sd), (N, sd, sd))."""

    def forward_loop(self, *, initial_states, controls):
        ctrl = SliceWrapper(controls)
        T, N = ctrl.shape[:2]
        assert initial_states.shape == (N, self.state_dim)
        means, trils = [], []
        current = initial_states
        for t in range(T):
            current, tril = self(initial_states=current, controls=ctrl[t])
            means.append(current)
            trils.append(tril)
        return torch.stack(means, dim=0), torch.stack(trils, dim=0)

    def jacobian(self, *, initial_states, controls):
        """Appendix A.4: A[n, i, j] = d f_i / d x_j, by reverse-mode autograd on an (N*sd)-row batch."""
        with torch.enable_grad():
            N, sd = initial_states.shape
            assert sd == self.state_dim
            tiled = initial_states[:, None, :].expand(N, sd, sd).detach().clone()
            tiled.requires_grad_(True)
            ctrl_tiled = SliceWrapper(controls).map(
                lambda c: torch.repeat_interleave(c, repeats=sd, dim=0)
            )
            preds, _ = self(initial_states=tiled.reshape(N * sd, sd), controls=ctrl_tiled)
            preds = preds.reshape(N, sd, sd)
            mask = torch.eye(sd, device=preds.device, dtype=preds.dtype)[None].expand(N, sd, sd)
            (jac,) = torch.autograd.grad(preds, tiled, mask, create_graph=True)
        return jac


class ParticleFilterMeasurementModel(nn.Module, abc.ABC):
    """(states (N, M, sd), observations) -> log-likelihoods (N, M)."""

    def __init__(self, *, state_dim: int):
        super().__init__()
        self.state_dim = state_dim

    @abc.abstractmethod
    def forward(self, *, states, observations):
        pass


class KalmanFilterMeasurementModel(nn.Module, abc.ABC):
    """states (N, sd) -> (expected observation (N, od), scale_tril (N, od, od))."""

    def __init__(self, *, state_dim: int, observation_dim: int):
        super().__init__()
        self.state_dim = state_dim
        self.observation_dim = observation_dim

    @abc.abstractmethod
    def forward(self, *, states):
        pass

    def jacobian(self, *, states):
        with torch.enable_grad():
            N, sd = states.shape
            od = self.observation_dim
            tiled = states[:, None, :].expand(N, od, sd).detach().clone()
            tiled.requires_grad_(True)
            preds, _ = self(states=tiled.reshape(N * od, sd))
            preds = preds.reshape(N, od, od)
            mask = torch.eye(od, device=preds.device, dtype=preds.dtype)[None].expand(N, od, od)
            (jac,) = torch.autograd.grad(preds, tiled, mask, create_graph=True)
        return jac  # (N, od, sd)


class VirtualSensorModel(nn.Module, abc.ABC):
    """observations -> (virtual state observation z (N, sd), scale_tril (N, sd, sd))."""

    def __init__(self, *, state_dim: int):
        super().__init__()
        self.state_dim = state_dim

    @abc.abstractmethod
    def forward(self, *, observations):
        pass


class KalmanFilterBase(Filter, abc.ABC):
    """Gaussian belief holder (Appendix A.1): forward = predict ; update ; return mean."""

    def __init__(self, *, dynamics_model: DynamicsModel, measurement_model: KalmanFilterMeasurementModel):
        super().__init__(state_dim=dynamics_model.state_dim)
        assert isinstance(dynamics_model, DynamicsModel)
        assert isinstance(measurement_model, KalmanFilterMeasurementModel)
        self.dynamics_model = dynamics_model
        self.measurement_model = measurement_model
        self._belief_mean = None
        self._belief_covariance = None
        self._initialized = False

    @property
    def belief_mean(self):
        return self._belief_mean

    @belief_mean.setter
    def belief_mean(self, mean):
        assert mean.shape[1:] == (self.state_dim,)
        self._belief_mean = mean

    @property
    def belief_covariance(self):
        return self._belief_covariance

    @belief_covariance.setter
    def belief_covariance(self, covariance):
        assert covariance.shape[1:] == (self.state_dim, self.state_dim)
        self._belief_covariance = covariance

    def initialize_beliefs(self, *, mean, covariance):
        N = mean.shape[0]
        assert mean.shape == (N, self.state_dim)
        assert covariance.shape == (N, self.state_dim, self.state_dim)
        self.belief_mean = mean
        self.belief_covariance = covariance
        self._initialized = True

    def forward(self, *, observations, controls):
        assert self._initialized, "Kalman filter not initialized: call initialize_beliefs() first"
        N = self._belief_mean.shape[0]
        assert SliceWrapper(controls).shape[0] == N
        self._predict_step(controls=controls)
        self._update_step(observations=observations)
        return self.belief_mean

    @abc.abstractmethod
    def _predict_step(self, *, controls):
        pass

    @abc.abstractmethod
    def _update_step(self, *, observations):
        pass
