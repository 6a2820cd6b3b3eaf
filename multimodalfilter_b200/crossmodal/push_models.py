"""ref: crossmodal/push_models/__init__.py:5-21 (LSTM baseline excluded: not a filtering recursion)."""
from .models import MODEL_TYPES as _ALL
from .models import (  # noqa: F401
    PushCrossmodalKalmanFilter,
    PushCrossmodalKalmanFilterWeightModel,
    PushCrossmodalParticleFilter,
    PushCrossmodalParticleFilterSeq5,
    PushCrossmodalWeightModel,
    PushDynamicsModel,
    PushKalmanFilter,
    PushMeasurementCrossmodalKalmanFilter,
    PushMeasurementModel,
    PushMeasurementUnimodalKalmanFilter,
    PushParticleFilter,
    PushUnimodalKalmanFilter,
    PushUnimodalParticleFilter,
    PushVirtualSensorModel,
)

model_types = _ALL["push"]
