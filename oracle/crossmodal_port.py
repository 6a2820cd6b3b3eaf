"""ORACLE (test infrastructure): independent restatement of the reference's ``crossmodal`` models.

One task-parameterised implementation instead of the reference's per-task file pairs; class
names, constructor behaviour, ``state_dict`` keys/shapes and arithmetic follow the cited lines.
Checked against the reference's own code by ``oracle/make_golden.py`` -> ``tests/golden/``.

ref: crossmodal/base_models/{crossmodal_pf,crossmodal_kf,unimodal_kf,utility}.py
ref: crossmodal/{push,door}_models/{layers,dynamics,pf,kf,crossmodal_pf,crossmodal_kf,
     unimodal_pf,unimodal_kf}.py
"""
import abc
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from oracle import install_shims

install_shims()
import torchfilter  # noqa: E402  (the oracle shim)
from fannypack.nn import resblocks  # noqa: E402

UNITS = 64
CONTROL_DIM = 7  # ref: crossmodal/push_models/layers.py:6, door_models/layers.py:6
POS_DIM = 3
SENSORS_DIM = 7
MODALITIES = ("image", "pos", "sensors")
OBS_KEY = {"image": "image", "pos": "gripper_pos", "sensors": "gripper_sensors"}


# ------------------------------------------------------------------------------------------------
# encoders (ref: crossmodal/push_models/layers.py:11-136, crossmodal/door_models/layers.py:11-95)
# ------------------------------------------------------------------------------------------------
def vector_encoder(in_dim: int, units: int = UNITS) -> nn.Sequential:
    """Linear -> ReLU -> residual block; used for states, controls, gripper pos, F/T sensors."""
    return nn.Sequential(nn.Linear(in_dim, units), nn.ReLU(inplace=True), resblocks.Linear(units))


class _RowColumnMeans(nn.Module):
    """ref: crossmodal/push_models/layers.py:43-65 -- full-height and full-width average pools
    (window 2 along the other axis), flattened and concatenated."""

    def __init__(self, rows: int, cols: int, reduce_size: int):
        super().__init__()
        self.pool_h = nn.Sequential(nn.AvgPool2d((rows, reduce_size)), nn.Flatten())
        self.pool_w = nn.Sequential(nn.AvgPool2d((reduce_size, cols)), nn.Flatten())

    def forward(self, x):
        return torch.cat((self.pool_h(x), self.pool_w(x)), dim=-1)


def image_encoder(units: int = UNITS, spanning_avg_pool: bool = False) -> nn.Sequential:
    """32x32 single-channel image -> units.  ref: crossmodal/push_models/layers.py:68-104."""
    trunk = [
        nn.Conv2d(1, 32, kernel_size=5, padding=2),
        nn.ReLU(inplace=True),
        resblocks.Conv2d(channels=32, kernel_size=3),
        nn.Conv2d(32, 16, kernel_size=3, padding=1),
        nn.ReLU(inplace=True),
    ]
    if spanning_avg_pool:
        head = [
            nn.Conv2d(16, 2, kernel_size=3, padding=1),
            _RowColumnMeans(rows=32, cols=32, reduce_size=2),
            nn.Linear(32 * 2, units),
        ]
    else:
        head = [nn.Conv2d(16, 8, kernel_size=3, padding=1), nn.Flatten(), nn.Linear(8 * 32 * 32, units)]
    return nn.Sequential(*trunk, *head, nn.ReLU(inplace=True), resblocks.Linear(units))


class _ObservationEncoders:
    """Mixin: builds ``observation_{image,pos,sensors}_layers`` for a modality subset and
    concatenates their features in the fixed image/pos/sensors order
    (ref: crossmodal/push_models/pf.py:43-50,76-88)."""

    def _build_encoders(self, modalities, units, spanning_avg_pool=False):
        assert len(set(MODALITIES) | set(modalities)) == 3, "Received invalid modality"
        assert len(modalities) > 0, "Received empty modality list"
        self.modalities = set(modalities)
        if "image" in self.modalities:
            self.observation_image_layers = image_encoder(units, spanning_avg_pool)
        if "pos" in self.modalities:
            self.observation_pos_layers = vector_encoder(POS_DIM, units)
        if "sensors" in self.modalities:
            self.observation_sensors_layers = vector_encoder(SENSORS_DIM, units)

    def observation_features(self, observations) -> torch.Tensor:
        feats = []
        if "image" in self.modalities:
            feats.append(self.observation_image_layers(observations["image"][:, None, :, :]))
        if "pos" in self.modalities:
            feats.append(self.observation_pos_layers(observations["gripper_pos"]))
        if "sensors" in self.modalities:
            feats.append(self.observation_sensors_layers(observations["gripper_sensors"]))
        return torch.cat(feats, dim=1)


def image_is_blacked_out(observations) -> torch.Tensor:
    """ref: crossmodal/push_models/crossmodal_pf.py:97-101."""
    image = observations["image"]
    N = image.shape[0]
    return torch.sum(torch.abs(image.reshape((N, -1))), dim=1) < 1e-8


# ------------------------------------------------------------------------------------------------
# dynamics (ref: crossmodal/push_models/dynamics.py:10-64, crossmodal/door_models/dynamics.py)
# ------------------------------------------------------------------------------------------------
class _GatedResidualDynamics(torchfilter.base.DynamicsModel):
    """x' = x + h[:sd] * sigmoid(h[sd]),  h = shared(cat(control_feats, state_feats))."""

    def __init__(self, state_dim: int, units: int = UNITS):
        super().__init__(state_dim=state_dim)
        self._register_noise()
        self.state_layers = vector_encoder(state_dim, units)
        self.control_layers = vector_encoder(CONTROL_DIM, units)
        self.shared_layers = nn.Sequential(
            nn.Linear(units * 2, units),
            resblocks.Linear(units),
            resblocks.Linear(units),
            resblocks.Linear(units),
            nn.Linear(units, state_dim + 1),
        )
        self.units = units

    def _scale_tril(self) -> torch.Tensor:
        return self.Q_scale_tril

    def forward(self, *, initial_states, controls):
        N, sd = initial_states.shape[:2]
        assert sd == self.state_dim
        merged = torch.cat((self.control_layers(controls), self.state_layers(initial_states)), dim=-1)
        h = self.shared_layers(merged)
        update = h[..., :sd] * torch.sigmoid(h[..., -1:])
        return initial_states + update, self._scale_tril()[None, :, :].expand(N, sd, sd)


class PushDynamicsModel(_GatedResidualDynamics):
    def __init__(self, units: int = UNITS):
        super().__init__(state_dim=2, units=units)

    def _register_noise(self):  # ref: crossmodal/push_models/dynamics.py:17-20
        self.Q_scale_tril = nn.Parameter(
            torch.linalg.cholesky(torch.diag(torch.FloatTensor([0.02, 0.02]))), requires_grad=False
        )


class DoorDynamicsModel(_GatedResidualDynamics):
    def __init__(self, units: int = UNITS):
        super().__init__(state_dim=3, units=units)

    def _register_noise(self):  # ref: crossmodal/door_models/dynamics.py:20-23
        self.Q_scale_tril = nn.Parameter(
            torch.linalg.cholesky(torch.diag(torch.FloatTensor([0.05, 0.01, 0.01]))),
            requires_grad=False,
        )


class DoorDynamicsModelBrent(_GatedResidualDynamics):
    def __init__(self, units: int = UNITS):
        super().__init__(state_dim=3, units=units)

    def _register_noise(self):  # ref: crossmodal/door_models/dynamics.py:85-88
        self.Q_scale_tril_diag = nn.Parameter(
            torch.sqrt(torch.FloatTensor([0.05, 0.01, 0.01])) / 8.0, requires_grad=False
        )

    def _scale_tril(self):  # ref: crossmodal/door_models/dynamics.py:131-133
        return torch.diag(self.Q_scale_tril_diag)


# ------------------------------------------------------------------------------------------------
# particle-filter measurement heads (ref: crossmodal/push_models/pf.py:30-109)
# ------------------------------------------------------------------------------------------------
class _MeasurementHead(torchfilter.base.ParticleFilterMeasurementModel, _ObservationEncoders):
    STATE_DIM = 0

    def __init__(self, units: int = UNITS, modalities=frozenset(MODALITIES)):
        super().__init__(state_dim=self.STATE_DIM)
        self._build_encoders(modalities, units)
        self.state_layers = vector_encoder(self.STATE_DIM, units)
        self.shared_layers = nn.Sequential(
            nn.Linear(units * (1 + len(self.modalities)), units),
            nn.ReLU(inplace=True),
            resblocks.Linear(units),
            resblocks.Linear(units),
            nn.Linear(units, 1),
        )
        self.units = units

    def forward(self, *, states, observations):
        assert type(observations) == dict
        assert len(states.shape) == 3 and states.shape[2] == self.state_dim
        N, M, _ = states.shape
        obs_feats = self.observation_features(observations)
        obs_feats = obs_feats[:, None, :].expand(N, M, obs_feats.shape[1])
        merged = torch.cat((obs_feats, self.state_layers(states)), dim=2)
        return self.shared_layers(merged).squeeze(dim=2)


class PushMeasurementModel(_MeasurementHead):
    STATE_DIM = 2


class DoorMeasurementModel(_MeasurementHead):
    STATE_DIM = 3


# ------------------------------------------------------------------------------------------------
# particle-filter fusion (ref: crossmodal/base_models/crossmodal_pf.py:11-141)
# ------------------------------------------------------------------------------------------------
class CrossmodalWeightModel(nn.Module, abc.ABC):
    def __init__(self, modality_count: int):
        super().__init__()
        self.modality_count = modality_count

    @abc.abstractmethod
    def forward(self, *, observations) -> torch.Tensor:
        """(N, modality_count) un-normalised log-weights."""


class _EnabledModelsMixin:
    """``enabled_models`` property with the reference's validation
    (ref: crossmodal/base_models/crossmodal_pf.py:60-85); scripts also poke ``_enabled_models``."""

    def _init_enabled(self, count: int):
        self._enabled_models: List[bool] = [True] * count

    @property
    def enabled_models(self) -> List[bool]:
        return self._enabled_models

    @enabled_models.setter
    def enabled_models(self, value: List[bool]) -> None:
        assert isinstance(value, list)
        assert len(value) == len(self._enabled_models)
        for flag in value:
            assert type(flag) == bool
        self._enabled_models = value


class CrossmodalParticleFilterMeasurementModel(
    torchfilter.base.ParticleFilterMeasurementModel, _EnabledModelsMixin
):
    def __init__(self, *, measurement_models, crossmodal_weight_model: Optional[CrossmodalWeightModel], state_dim: int):
        super().__init__(state_dim=state_dim)
        self.measurement_models = nn.ModuleList(measurement_models)
        self.crossmodal_weight_model = crossmodal_weight_model
        self._init_enabled(len(self.measurement_models))

    def forward(self, *, states, observations):
        N, M, _ = states.shape
        enabled = self._enabled_models
        per_modality = torch.stack(
            [m(states=states, observations=observations) for i, m in enumerate(self.measurement_models) if enabled[i]],
            dim=2,
        )
        assert per_modality.shape == (N, M, int(np.sum(enabled)))
        if self.crossmodal_weight_model is None:
            return torch.logsumexp(per_modality, dim=2)
        log_weights = self.crossmodal_weight_model(observations=observations)[:, enabled]
        assert log_weights.shape == (N, int(np.sum(enabled)))
        # (the reference also computes a max-normalised copy that it never uses, :124-129 -- quirk Q1)
        return torch.logsumexp(log_weights[:, None, :] + per_modality, dim=2)


class _PFWeightModel(CrossmodalWeightModel, _ObservationEncoders):
    """ref: crossmodal/push_models/crossmodal_pf.py:52-104, door_models/crossmodal_pf.py:52-106."""

    RESBLOCKS = 1

    def __init__(self, know_image_blackout: bool, units: int = UNITS):
        super().__init__(modality_count=2)
        self.know_image_blackout = know_image_blackout
        self._build_encoders(MODALITIES, units)
        self.fusion_layers = nn.Sequential(
            nn.Linear(units * 3, units),
            nn.ReLU(inplace=True),
            *[resblocks.Linear(units) for _ in range(self.RESBLOCKS)],
            nn.Linear(units, self.modality_count),
        )

    def forward(self, *, observations):
        N = observations["gripper_pos"].shape[0]
        out = self.fusion_layers(self.observation_features(observations))
        assert out.shape == (N, self.modality_count)
        if self.know_image_blackout:  # quirk Q3: -inf image weight on blacked-out frames
            out[image_is_blacked_out(observations), 0] -= np.inf
        return out


class PushCrossmodalWeightModel(_PFWeightModel):
    RESBLOCKS = 1


class DoorCrossmodalWeightModel(_PFWeightModel):
    RESBLOCKS = 3


class _TaskParticleFilter(torchfilter.filters.ParticleFilter):
    def train(self, mode: bool = True):  # quirk Q8: ref: crossmodal/push_models/pf.py:24-27
        self.num_particles = 30 if mode else 300
        return super().train(mode)


def _pf_family(prefix, dynamics_cls, head_cls, weight_cls, state_dim):
    class Plain(_TaskParticleFilter):
        def __init__(self):
            super().__init__(dynamics_model=dynamics_cls(), measurement_model=head_cls(), num_particles=30)

    class Crossmodal(_TaskParticleFilter):
        def __init__(self, know_image_blackout: bool = False):
            super().__init__(
                dynamics_model=dynamics_cls(),
                measurement_model=CrossmodalParticleFilterMeasurementModel(
                    measurement_models=[head_cls(modalities={"image"}), head_cls(modalities={"pos", "sensors"})],
                    crossmodal_weight_model=weight_cls(know_image_blackout=know_image_blackout),
                    state_dim=state_dim,
                ),
                num_particles=30,
            )

    class CrossmodalSeq5(Crossmodal):
        def __init__(self):
            super().__init__(know_image_blackout=True)

    class Unimodal(_TaskParticleFilter):
        def __init__(self):
            super().__init__(
                dynamics_model=dynamics_cls(),
                measurement_model=CrossmodalParticleFilterMeasurementModel(
                    measurement_models=[head_cls(modalities={"image"}), head_cls(modalities={"pos", "sensors"})],
                    crossmodal_weight_model=None,
                    state_dim=state_dim,
                ),
                num_particles=30,
            )

    for cls, name in (
        (Plain, "ParticleFilter"),
        (Crossmodal, "CrossmodalParticleFilter"),
        (CrossmodalSeq5, "CrossmodalParticleFilterSeq5"),
        (Unimodal, "UnimodalParticleFilter"),
    ):
        cls.__name__ = cls.__qualname__ = prefix + name
    return Plain, Crossmodal, CrossmodalSeq5, Unimodal


(PushParticleFilter, PushCrossmodalParticleFilter, PushCrossmodalParticleFilterSeq5, PushUnimodalParticleFilter) = _pf_family(
    "Push", PushDynamicsModel, PushMeasurementModel, PushCrossmodalWeightModel, 2
)
(DoorParticleFilter, DoorCrossmodalParticleFilter, DoorCrossmodalParticleFilterSeq5, DoorUnimodalParticleFilter) = _pf_family(
    "Door", DoorDynamicsModelBrent, DoorMeasurementModel, DoorCrossmodalWeightModel, 3
)


# ------------------------------------------------------------------------------------------------
# Kalman side: virtual sensors (ref: crossmodal/door_models/kf.py:31-126, push_models/kf.py:31-128)
# ------------------------------------------------------------------------------------------------
class _VirtualSensor(torchfilter.base.VirtualSensorModel, _ObservationEncoders):
    STATE_DIM = 0
    SPANNING_AVG_POOL = False

    def __init__(self, units: int = UNITS, modalities=frozenset(MODALITIES), add_R_noise: float = 1e-6, noise_R_tril=None):
        super().__init__(state_dim=self.STATE_DIM)
        sd = self.STATE_DIM
        self.noise_R_tril = noise_R_tril
        self._build_encoders(modalities, units, spanning_avg_pool=self.SPANNING_AVG_POOL)
        self.shared_layers = nn.Sequential(
            nn.Linear(units * len(self.modalities), units * 2),
            nn.ReLU(inplace=True),
            resblocks.Linear(units * 2),
            resblocks.Linear(units * 2),
        )

        def small_head():
            return nn.Sequential(nn.Linear(units, sd), nn.ReLU(inplace=True), resblocks.Linear(sd), nn.Linear(sd, sd))

        self.r_layer = small_head()
        self.z_layer = small_head()
        self.units = units
        self.add_R_noise = torch.ones(sd) * add_R_noise

    def forward(self, *, observations):
        assert type(observations) == dict
        N = observations["gripper_pos"].shape[0]
        feats = self.observation_features(observations)
        assert feats.shape == (N, self.units * len(self.modalities))
        shared = self.shared_layers(feats)
        z = self.z_layer(shared[:, : self.units].clone())
        assert z.shape == (N, self.state_dim)
        if self.noise_R_tril is None:
            lt_hat = self.r_layer(shared[:, self.units :].clone())
        else:
            lt_hat = self.noise_R_tril
        lt = torch.diag_embed(lt_hat, offset=0, dim1=-2, dim2=-1)
        assert lt.shape == (N, self.state_dim, self.state_dim)
        R = lt ** 2
        if self.add_R_noise[0] > 0:
            R = R + torch.diag(self.add_R_noise).to(R.device)
        return z, torch.sqrt(R)  # elementwise sqrt of a diagonal matrix (quirk list, section 7)


class PushVirtualSensorModel(_VirtualSensor):
    STATE_DIM = 2
    SPANNING_AVG_POOL = True  # ref: crossmodal/push_models/kf.py:50-52


class DoorVirtualSensorModel(_VirtualSensor):
    STATE_DIM = 3


def _kalman_filter_cls(name, dynamics_cls, sensor_cls):
    def __init__(self, dynamics_model=None, virtual_sensor_model=None):
        if dynamics_model is None and virtual_sensor_model is None:
            dynamics_model, virtual_sensor_model = dynamics_cls(), sensor_cls()
        torchfilter.filters.VirtualSensorExtendedKalmanFilter.__init__(
            self, dynamics_model=dynamics_model, virtual_sensor_model=virtual_sensor_model
        )

    return type(name, (torchfilter.filters.VirtualSensorExtendedKalmanFilter,), {"__init__": __init__})


PushKalmanFilter = _kalman_filter_cls("PushKalmanFilter", PushDynamicsModel, PushVirtualSensorModel)
DoorKalmanFilter = _kalman_filter_cls("DoorKalmanFilter", DoorDynamicsModel, DoorVirtualSensorModel)


# ------------------------------------------------------------------------------------------------
# Kalman fusion (ref: crossmodal/base_models/{utility,crossmodal_kf,unimodal_kf}.py)
# ------------------------------------------------------------------------------------------------
def weighted_average(predictions, weights):
    """ref: crossmodal/base_models/utility.py:4-11."""
    assert predictions.shape == weights.shape
    weights = weights / (torch.sum(weights, dim=0) + 1e-9)
    return torch.sum(weights * predictions, dim=0)


class CrossmodalKalmanFilterWeightModel(nn.Module, abc.ABC):
    def __init__(self, modality_count: int, state_dim: int):
        super().__init__()
        self.modality_count = modality_count
        self.state_dim = state_dim

    @abc.abstractmethod
    def forward(self, *, observations) -> torch.Tensor:
        """(modality_count, N, state_dim) weights."""


def _mask_weights(enabled, N, state_dim, device):
    """ref: crossmodal/base_models/crossmodal_kf.py:124-131 -- 1/0 weights when a model is off."""
    w = torch.from_numpy(np.array(enabled).astype(np.float32))
    return w.unsqueeze(-1).unsqueeze(-1).repeat(1, N, state_dim).to(device)


def _measurement_level_fusion(unimodal_states, unimodal_scale_trils, state_weights):
    """ref: crossmodal/base_models/crossmodal_kf.py:219-235 and :337-354."""
    covs = unimodal_scale_trils @ unimodal_scale_trils.transpose(-1, -2)
    mean = weighted_average(unimodal_states, state_weights)
    mult = torch.prod(torch.prod(state_weights, dim=-1), dim=0).unsqueeze(-1).unsqueeze(-1)
    return mean, mult * torch.sum(covs, dim=0)


class CrossmodalKalmanFilter(torchfilter.base.Filter, _EnabledModelsMixin):
    """ref: crossmodal/base_models/crossmodal_kf.py:39-240."""

    def __init__(self, *, filter_models, crossmodal_weight_model, state_dim: int):
        super().__init__(state_dim=state_dim)
        self.filter_models = nn.ModuleList(filter_models)
        self.crossmodal_weight_model = crossmodal_weight_model
        self._init_enabled(len(self.filter_models))
        self.weighted_covariances = None

    def calculate_unimodal_states(self, observations, controls):
        on = [f for i, f in enumerate(self.filter_models) if self._enabled_models[i]]
        states = torch.stack([f(observations=observations, controls=controls) for f in on])
        covs = torch.stack([f._belief_covariance for f in on])
        return states, covs

    def calculate_weighted_states(self, state_weights, unimodal_states, unimodal_covariances):
        K, N, sd = state_weights.shape
        assert K == np.sum(self._enabled_models) and sd == self.state_dim
        mean = weighted_average(unimodal_states, state_weights)
        cw = state_weights.unsqueeze(-1).repeat((1, 1, 1, sd))
        cw = cw * cw.transpose(-1, -2)  # outer product beta beta^T
        return mean, torch.sum(cw * unimodal_covariances, 0)

    def forward(self, *, observations, controls):
        N = controls.shape[0]
        enabled = self._enabled_models
        K = int(np.sum(enabled))
        states, covs = self.calculate_unimodal_states(observations, controls)
        assert states.shape == (K, N, self.state_dim)
        assert covs.shape == (K, N, self.state_dim, self.state_dim)
        if K < len(enabled):
            weights = _mask_weights(enabled, N, self.state_dim, states.device)
        else:
            weights = self.crossmodal_weight_model(observations=observations)
        weights = weights[enabled]
        assert weights.shape == (K, N, self.state_dim)
        mean, cov = self.calculate_weighted_states(weights, states, covs)
        self.weighted_covariances = cov
        for f in self.filter_models:  # quirk Q6: attributes nobody reads => no posterior feedback
            f.states_prev = mean
            f.states_covariance_prev = cov
        return mean

    @property
    def state_covariance_estimate(self):
        return self.weighted_covariances

    def initialize_beliefs(self, *, mean, covariance):
        N = mean.shape[0]
        assert mean.shape == (N, self.state_dim)
        assert covariance.shape == (N, self.state_dim, self.state_dim)
        for f in self.filter_models:
            f.initialize_beliefs(mean=mean, covariance=covariance)

    def measurement_initialize_beliefs(self, observations):
        on = [f for i, f in enumerate(self.filter_models) if self._enabled_models[i]]
        outs = [f.virtual_sensor_model(observations=observations) for f in on]
        weights = self.crossmodal_weight_model(observations=observations)[self._enabled_models]
        mean, cov = _measurement_level_fusion(
            torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]), weights
        )
        self.initialize_beliefs(mean=mean, covariance=cov)


class CrossmodalVirtualSensorModel(torchfilter.base.VirtualSensorModel, _EnabledModelsMixin):
    """ref: crossmodal/base_models/crossmodal_kf.py:243-359."""

    def __init__(self, *, virtual_sensor_model, crossmodal_weight_model, state_dim: int):
        super().__init__(state_dim=state_dim)
        self.virtual_sensor_model = nn.ModuleList(virtual_sensor_model)
        self.crossmodal_weight_model = crossmodal_weight_model
        self._init_enabled(len(self.virtual_sensor_model))

    def forward(self, *, observations):
        enabled = self._enabled_models
        N = observations[[*observations][0]].shape[0]
        outs = [m(observations=observations) for i, m in enumerate(self.virtual_sensor_model) if enabled[i]]
        states = torch.stack([o[0] for o in outs])
        trils = torch.stack([o[1] for o in outs])
        if np.sum(enabled) < len(enabled):
            weights = _mask_weights(enabled, N, self.state_dim, states.device)
        else:
            weights = self.crossmodal_weight_model(observations=observations)
        weights = weights[enabled]
        mean, cov = _measurement_level_fusion(states, trils, weights)
        return mean, torch.linalg.cholesky(cov)


class UnimodalVirtualSensorModel(torchfilter.base.VirtualSensorModel, _EnabledModelsMixin):
    """ref: crossmodal/base_models/unimodal_kf.py:13-115 (returns a covariance, not a tril)."""

    def __init__(self, *, virtual_sensor_model, state_dim: int):
        super().__init__(state_dim=state_dim)
        self.virtual_sensor_model = nn.ModuleList(virtual_sensor_model)
        self._init_enabled(len(self.virtual_sensor_model))

    def forward(self, *, observations):
        enabled = self._enabled_models
        outs = [m(observations=observations) for i, m in enumerate(self.virtual_sensor_model) if enabled[i]]
        states = torch.stack([o[0] for o in outs])
        trils = torch.stack([o[1] for o in outs])
        covs = trils @ trils.transpose(-1, -2)
        if np.sum(enabled) == 1:
            return states[0], covs[0]
        precision = torch.stack([1.0 / (o[1] + 1e-9) for o in outs])  # elementwise, on the tril
        weights = torch.diagonal(precision, dim1=-2, dim2=-1).squeeze(1)
        assert weights.shape == states.shape
        mean = weighted_average(states, weights)
        return mean, torch.inverse(torch.sum(precision, dim=0) + 1e-9)


class UnimodalKalmanFilter(torchfilter.base.Filter, _EnabledModelsMixin):
    """ref: crossmodal/base_models/unimodal_kf.py:118-270 -- information-form fusion."""

    def __init__(self, *, filter_models, state_dim: int):
        super().__init__(state_dim=state_dim)
        self.filter_models = nn.ModuleList(filter_models)
        self._init_enabled(len(self.filter_models))
        self.weighted_covariances = None

    def forward(self, *, observations, controls):
        N = controls.shape[0]
        sd = self.state_dim
        on = [f for i, f in enumerate(self.filter_models) if self._enabled_models[i]]
        states = torch.stack([f(observations=observations, controls=controls) for f in on])
        covs = torch.stack([f._belief_covariance for f in on])
        if len(on) == 1:
            return states[0]
        precision = torch.stack([torch.inverse(f._belief_covariance + 1e-9) for f in on])
        cov = torch.inverse(torch.sum(precision, dim=0) + 1e-9)
        info = precision.reshape(-1, sd, sd).bmm(states.reshape(-1, sd, 1)).reshape(len(on), N, sd, 1)
        mean = cov.bmm(torch.sum(info, dim=0)).squeeze(-1)
        assert mean.shape == (N, sd) and covs.shape == (len(on), N, sd, sd)
        return mean

    @property
    def state_covariance_estimate(self):
        return self.weighted_covariances

    def initialize_beliefs(self, *, mean, covariance):
        N = mean.shape[0]
        assert mean.shape == (N, self.state_dim)
        assert covariance.shape == (N, self.state_dim, self.state_dim)
        for f in self.filter_models:
            f.initialize_beliefs(mean=mean, covariance=covariance)


class _KFWeightModel(CrossmodalKalmanFilterWeightModel, _ObservationEncoders):
    """ref: crossmodal/door_models/crossmodal_kf.py:101-167 (push twin identical)."""

    def __init__(self, units: int = UNITS, state_dim: int = 2, know_image_blackout=False):
        super().__init__(modality_count=2, state_dim=state_dim)
        self._build_encoders(MODALITIES, units)
        self.weighting_type = "sigmoid"
        self.fusion_layers = nn.Sequential(
            nn.Linear(units * 3, units),
            nn.ReLU(inplace=True),
            resblocks.Linear(units),
            nn.Linear(units, self.modality_count * self.state_dim),
            nn.Sigmoid(),
        )
        self.know_image_blackout = know_image_blackout

    def forward(self, *, observations):
        N = observations["gripper_pos"].shape[0]
        out = self.fusion_layers(self.observation_features(observations))
        assert out.shape == (N, self.modality_count * self.state_dim)
        beta = out.reshape(self.modality_count, N, self.state_dim)  # quirk Q5: NOT a transpose
        return beta / (torch.sum(beta, dim=0) + 1e-9)


class PushCrossmodalKalmanFilterWeightModel(_KFWeightModel):
    pass


class DoorCrossmodalKalmanFilterWeightModel(_KFWeightModel):
    pass


class _TaskCrossmodalKalmanFilter(CrossmodalKalmanFilter):
    """Blackout-aware forward override (ref: crossmodal/door_models/crossmodal_kf.py:43-98)."""

    def forward(self, *, observations, controls):
        if not self.know_image_blackout:
            return super().forward(observations=observations, controls=controls)
        N = controls.shape[0]
        device = controls.device
        black = image_is_blacked_out(observations)
        if torch.sum(black) == 0 or np.sum(self._enabled_models) < len(self._enabled_models):
            return super().forward(observations=observations, controls=controls)
        states, covs = self.calculate_unimodal_states(observations, controls)
        raw = self.crossmodal_weight_model(observations=observations)
        keep = torch.ones((N, 1), device=device)
        keep[black] = 0
        image_floor = torch.zeros((N, 1), device=device)
        image_floor[black] = 1e-9
        force_floor = torch.zeros((N, 1), device=device)
        force_floor[black] = 1.0 - 1e-9
        weights = torch.stack([image_floor + keep * raw[0], force_floor + keep * raw[1]])  # quirk Q7
        mean, cov = self.calculate_weighted_states(weights, states, covs)
        self.weighted_covariances = cov
        return mean


def _kf_family(prefix, dynamics_cls, sensor_cls, kf_cls, weight_cls, sd):
    def two_filters():
        return [
            kf_cls(dynamics_model=dynamics_cls(), virtual_sensor_model=sensor_cls(modalities={"image"})),
            kf_cls(dynamics_model=dynamics_cls(), virtual_sensor_model=sensor_cls(modalities={"pos", "sensors"})),
        ]

    def two_sensors():
        return [sensor_cls(modalities={"image"}), sensor_cls(modalities={"pos", "sensors"})]

    class Crossmodal(_TaskCrossmodalKalmanFilter):
        def __init__(self, know_image_blackout=False):
            super().__init__(filter_models=two_filters(), crossmodal_weight_model=weight_cls(state_dim=sd), state_dim=sd)
            self.know_image_blackout = know_image_blackout

    class Unimodal(UnimodalKalmanFilter):
        def __init__(self):
            super().__init__(filter_models=two_filters(), state_dim=sd)

    class MeasurementCrossmodal(kf_cls):
        def __init__(self):
            super().__init__(
                dynamics_model=dynamics_cls(),
                virtual_sensor_model=CrossmodalVirtualSensorModel(
                    virtual_sensor_model=two_sensors(), crossmodal_weight_model=weight_cls(state_dim=sd), state_dim=sd
                ),
            )

    class MeasurementUnimodal(kf_cls):
        def __init__(self):
            super().__init__(
                dynamics_model=dynamics_cls(),
                virtual_sensor_model=UnimodalVirtualSensorModel(virtual_sensor_model=two_sensors(), state_dim=sd),
            )

    for cls, name in (
        (Crossmodal, "CrossmodalKalmanFilter"),
        (Unimodal, "UnimodalKalmanFilter"),
        (MeasurementCrossmodal, "MeasurementCrossmodalKalmanFilter"),
        (MeasurementUnimodal, "MeasurementUnimodalKalmanFilter"),
    ):
        cls.__name__ = cls.__qualname__ = prefix + name
    return Crossmodal, Unimodal, MeasurementCrossmodal, MeasurementUnimodal


# NB: the reference's *Push* MeasurementCrossmodal/MeasurementUnimodal types are broken as shipped
# (ref: crossmodal/push_models/crossmodal_kf.py:175, unimodal_kf.py:41-46); the port builds the
# obviously intended objects, and no parity is claimed for those two.
(PushCrossmodalKalmanFilter, PushUnimodalKalmanFilter, PushMeasurementCrossmodalKalmanFilter, PushMeasurementUnimodalKalmanFilter) = _kf_family(
    "Push", PushDynamicsModel, PushVirtualSensorModel, PushKalmanFilter, PushCrossmodalKalmanFilterWeightModel, 2
)
(DoorCrossmodalKalmanFilter, DoorUnimodalKalmanFilter, DoorMeasurementCrossmodalKalmanFilter, DoorMeasurementUnimodalKalmanFilter) = _kf_family(
    "Door", DoorDynamicsModel, DoorVirtualSensorModel, DoorKalmanFilter, DoorCrossmodalKalmanFilterWeightModel, 3
)

MODEL_TYPES = {
    cls.__name__: cls
    for cls in (
        PushParticleFilter, PushCrossmodalParticleFilter, PushCrossmodalParticleFilterSeq5, PushUnimodalParticleFilter,
        PushKalmanFilter, PushCrossmodalKalmanFilter, PushUnimodalKalmanFilter,
        DoorParticleFilter, DoorCrossmodalParticleFilter, DoorCrossmodalParticleFilterSeq5, DoorUnimodalParticleFilter,
        DoorKalmanFilter, DoorCrossmodalKalmanFilter, DoorUnimodalKalmanFilter,
        DoorMeasurementCrossmodalKalmanFilter, DoorMeasurementUnimodalKalmanFilter,
    )
}
