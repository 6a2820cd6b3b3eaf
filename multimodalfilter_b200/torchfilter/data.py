"""``torchfilter.data``: the datasets the reference's training helpers build
(ref: crossmodal/train_helpers.py:39,63-65,83-87,113,146-148).

Host-side glue only: trajectories are numpy ``TrajectoryNumpy(states (T, sd), observations, controls)`` tuples; every
``__getitem__`` returns numpy leaves that the default collate turns into (batch, ...) tensors."""
from typing import List

import numpy as np
import scipy.stats
import torch.utils.data

from ..fannypack.utils import SliceWrapper
from .types import TrajectoryNumpy


def _leading_length(traj: TrajectoryNumpy) -> int:
    T = len(traj.states)
    assert len(SliceWrapper(traj.observations)) == T and len(SliceWrapper(traj.controls)) == T, \
        "states, observations and controls of a trajectory must share their length"
    return T


class SingleStepDataset(torch.utils.data.Dataset):
    """(state_{t}, state_{t+1}, observation_{t+1}, control_{t+1}) for every transition of every trajectory."""

    def __init__(self, trajectories: List[TrajectoryNumpy]):
        self.samples = []
        for traj in trajectories:
            T = _leading_length(traj)
            obs, ctrl = SliceWrapper(traj.observations), SliceWrapper(traj.controls)
            for t in range(T - 1):
                self.samples.append((traj.states[t], traj.states[t + 1], obs[t + 1], ctrl[t + 1]))

    def __getitem__(self, index):
        return self.samples[index]

    def __len__(self):
        return len(self.samples)


def split_trajectories(trajectories: List[TrajectoryNumpy], subsequence_length: int) -> List[TrajectoryNumpy]:
    """Chop every trajectory into sections of ``subsequence_length`` steps, once from the front and (when the length is
    not a multiple) once from the back, so that every step is covered."""
    out = []
    for traj in trajectories:
        T = _leading_length(traj)
        sections = T // subsequence_length
        if sections == 0:
            continue
        covered = sections * subsequence_length
        obs, ctrl = SliceWrapper(traj.observations), SliceWrapper(traj.controls)
        for offset in sorted({0, T - covered}):
            for s in range(sections):
                lo = offset + s * subsequence_length
                sl = slice(lo, lo + subsequence_length)
                out.append(TrajectoryNumpy(traj.states[sl], obs[sl], ctrl[sl]))
    return out


class SubsequenceDataset(torch.utils.data.Dataset):
    """Fixed-length (states (L, sd), observations (L, ...), controls (L, ...)) windows for BPTT training
    (ref: crossmodal/train_helpers.py:146-148: ``SubsequenceDataset(trajectories=, subsequence_length=)``)."""

    def __init__(self, trajectories: List[TrajectoryNumpy], subsequence_length: int):
        self.subsequences = split_trajectories(trajectories, subsequence_length)

    def __getitem__(self, index):
        t = self.subsequences[index]
        return t.states, t.observations, t.controls

    def __len__(self):
        return len(self.subsequences)


class ParticleFilterMeasurementDataset(torch.utils.data.Dataset):
    """(noisy state, observation, log-likelihood of the noisy state under N(true state, covariance)) triples for
    pre-training a particle-filter measurement model (ref: crossmodal/train_helpers.py:83-87).  Half of the
    ``samples_per_pair`` draws of a (state, observation) pair come from that Gaussian, the other half from a 5x wider
    one, so that the model also sees unlikely particles."""

    def __init__(self, trajectories: List[TrajectoryNumpy], *, covariance: np.ndarray, samples_per_pair: int,
                 seed: int = 0):
        self.covariance = np.asarray(covariance, dtype=np.float64)
        self.samples_per_pair = samples_per_pair
        self.pairs = []
        for traj in trajectories:
            T = _leading_length(traj)
            obs = SliceWrapper(traj.observations)
            self.pairs += [(traj.states[t], obs[t]) for t in range(T)]
        sd = self.covariance.shape[0]
        self._pdf = scipy.stats.multivariate_normal(mean=np.zeros(sd), cov=self.covariance)
        self._seed = seed

    def __getitem__(self, index):
        state, observation = self.pairs[index // self.samples_per_pair]
        rng = np.random.default_rng((self._seed, index))
        sd = self.covariance.shape[0]
        scale = 1.0 if (index % self.samples_per_pair) * 2 < self.samples_per_pair else 5.0
        offset = rng.multivariate_normal(np.zeros(sd), self.covariance * scale)
        noisy = (np.asarray(state, dtype=np.float64) + offset).astype(np.float32)
        log_likelihood = np.float32(self._pdf.logpdf(offset))
        return noisy, observation, log_likelihood

    def __len__(self):
        return len(self.pairs) * self.samples_per_pair
