"""ref: crossmodal/base_models/__init__.py:1-10."""
from .models import (  # noqa: F401
    CrossmodalKalmanFilter,
    CrossmodalKalmanFilterWeightModel,
    CrossmodalParticleFilterMeasurementModel,
    CrossmodalVirtualSensorModel,
    CrossmodalWeightModel,
    UnimodalKalmanFilter,
    UnimodalVirtualSensorModel,
    weighted_average,
)
