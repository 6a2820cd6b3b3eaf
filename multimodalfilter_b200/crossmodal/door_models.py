"""ref: crossmodal/door_models/__init__.py:5-19 (LSTM baseline excluded: not a filtering recursion)."""
from .models import MODEL_TYPES as _ALL
from .models import (  # noqa: F401
    DoorCrossmodalKalmanFilter,
    DoorCrossmodalKalmanFilterWeightModel,
    DoorCrossmodalParticleFilter,
    DoorCrossmodalParticleFilterSeq5,
    DoorCrossmodalWeightModel,
    DoorDynamicsModel,
    DoorDynamicsModelBrent,
    DoorKalmanFilter,
    DoorMeasurementCrossmodalKalmanFilter,
    DoorMeasurementModel,
    DoorMeasurementUnimodalKalmanFilter,
    DoorParticleFilter,
    DoorUnimodalKalmanFilter,
    DoorUnimodalParticleFilter,
    DoorVirtualSensorModel,
)

model_types = _ALL["door"]
