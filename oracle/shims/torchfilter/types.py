"""ORACLE: torchfilter.types (Appendix A.1).  Aliases only; TrajectoryNumpy is constructed
positionally at ref: crossmodal/tasks/_push.py:402-406."""
from typing import Any, Dict, NamedTuple, Union

import numpy as np
import torch

NumpyDict = Dict[str, np.ndarray]
TorchDict = Dict[str, torch.Tensor]
NumpyArrayOrDict = Union[np.ndarray, NumpyDict]
TorchTensorOrDict = Union[torch.Tensor, TorchDict]

StatesNumpy = np.ndarray
StatesTorch = torch.Tensor
ObservationsNumpy = NumpyArrayOrDict
ObservationsTorch = TorchTensorOrDict
ControlsNumpy = NumpyArrayOrDict
ControlsTorch = TorchTensorOrDict
ScaleTrilTorch = torch.Tensor
CovarianceTorch = torch.Tensor


class TrajectoryNumpy(NamedTuple):
    states: Any
    observations: Any
    controls: Any
