"""Type aliases of ``torchfilter.types`` (Appendix A.1)."""
from typing import Any, Dict, NamedTuple, Union

import numpy as np
import torch

NumpyDict = Dict[str, np.ndarray]
TorchDict = Dict[str, torch.Tensor]
StatesNumpy = np.ndarray
StatesTorch = torch.Tensor
ObservationsNumpy = Union[np.ndarray, NumpyDict]
ObservationsTorch = Union[torch.Tensor, TorchDict]
ControlsNumpy = Union[np.ndarray, NumpyDict]
ControlsTorch = Union[torch.Tensor, TorchDict]
ScaleTrilTorch = torch.Tensor
CovarianceTorch = torch.Tensor


class TrajectoryNumpy(NamedTuple):  # positional construction at ref: crossmodal/tasks/_push.py:402-406
    states: Any
    observations: Any
    controls: Any
