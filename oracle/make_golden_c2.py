"""ORACLE (test infrastructure): float64 evaluation of BASELINE config C2 (DoorCrossmodalKalmanFilter, 256 trajectories x
100 steps) by the oracle port -> tests/golden/c2_fp64.npz.

Why a float64 fixture: at this size the float32 CPU evaluation of the reference recursion is itself 3.3e-4 away from the
exact-arithmetic result at two ill-conditioned trajectories (180 and 184, around step 64; the errors decay again), i.e.
further than the 1e-4 parity bar.  The parity test therefore checks the CUDA path against this float64 evaluation
everywhere, and against the live float32 oracle wherever float32 and float64 agree.

    python oracle/make_golden_c2.py        (about two minutes on 8 cores)
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from multimodalfilter_b200.synthetic import fill_parameters, synthetic_trajectories  # noqa: E402
from oracle import crossmodal_port as port  # noqa: E402

NAME, SD, N, T, DATA_SEED, PARAM_SEED = "DoorCrossmodalKalmanFilter", 3, 256, 100, 34, 35
COV_STEPS = [0, 1, 2, 9, 19, 29, 39, 49, 59, 63, 64, 65, 69, 79, 89, 99]


def main():
    states, obs, controls = synthetic_trajectories(T + 1, N, SD, seed=DATA_SEED)
    dt = torch.float64
    cov = (torch.eye(SD, dtype=dt) * 0.1)[None].expand(N, SD, SD)
    o = fill_parameters(getattr(port, NAME)(), seed=PARAM_SEED).eval().to(dt)
    est, covs = [], {}
    with torch.no_grad():
        o.initialize_beliefs(mean=states[0].to(dt), covariance=cov)
        for t in range(T):
            est.append(o(observations={k: v[1 + t].to(dt) for k, v in obs.items()}, controls=controls[1 + t].to(dt)))
            if t in COV_STEPS:
                covs[t] = o.weighted_covariances.clone()
    out = {
        "estimates": torch.stack(est).numpy(),
        "cov_steps": np.asarray(COV_STEPS),
        "fused_covariances": torch.stack([covs[t] for t in COV_STEPS]).numpy(),
        "belief_means": torch.stack([f.belief_mean for f in o.filter_models]).numpy(),
        "belief_covariances": torch.stack([f.belief_covariance for f in o.filter_models]).numpy(),
        "seeds": np.asarray([DATA_SEED, PARAM_SEED]),
    }
    path = os.path.join(REPO, "tests", "golden", "c2_fp64.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
