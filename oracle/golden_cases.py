"""ORACLE (test infrastructure): the known-answer cases behind ``tests/golden/reference_modules.npz``.

``run_all(resolve)`` builds each model through ``resolve(class_name)``, fills it with the
name-keyed deterministic parameters, feeds the deterministic synthetic inputs and returns
``{case_key: ndarray}``.  The *same* function is run over
  * the reference's own ``crossmodal`` package (``oracle/make_golden.py``, only where
    ``/root/reference`` is mounted)  -> the committed fixture,
  * the oracle port (tests/test_oracle_golden.py, CPU),
  * the CUDA product (tests/test_gpu_parity.py, ``-m gpu``),
so a fixture mismatch is attributable to exactly one side.
"""
import numpy as np
import torch

from multimodalfilter_b200.synthetic import fill_parameters, synthetic_trajectories


def _cov(N, sd, device):
    return (torch.eye(sd, device=device) * 0.1)[None].expand(N, sd, sd)  # ref: crossmodal/eval_helpers.py:125-127


def _np(x):
    return x.detach().cpu().numpy().copy()


def _prep(model, seed, device, train=False):
    """fill + move + set mode; the reference's ``train()`` overrides return None, so never chain."""
    fill_parameters(model, seed=seed)
    model.to(device)
    model.train(train)
    return model


def _obs_at(obs, t, device):
    return {k: v[t].to(device) for k, v in obs.items()}


def run_all(resolve, device="cpu", include_rng_cases=True, only=None):
    out = {}
    dev = torch.device(device)

    def want(key):
        return only is None or any(key.startswith(p) for p in only)

    with torch.no_grad():
        # ---- R3: dynamics forward (+ autograd jacobian) -------------------------------------------
        for name, sd in (("PushDynamicsModel", 2), ("DoorDynamicsModel", 3), ("DoorDynamicsModelBrent", 3)):
            if not want("dyn/" + name):
                continue
            model = _prep(resolve(name)(), 11, dev)
            states, _, controls = synthetic_trajectories(1, 16, sd, seed=21)
            new_states, trils = model(initial_states=states[0].to(dev), controls=controls[0].to(dev))
            out[f"dyn/{name}/states"] = _np(new_states)
            out[f"dyn/{name}/tril"] = _np(trils[0])
            jac = model.jacobian(initial_states=states[0].to(dev), controls=controls[0].to(dev))
            out[f"dyn/{name}/jacobian"] = _np(jac)

        # ---- R4: per-particle measurement heads ------------------------------------------------
        for name, sd in (("PushMeasurementModel", 2), ("DoorMeasurementModel", 3)):
            for tag, mods in (("image", {"image"}), ("force", {"pos", "sensors"}), ("all", {"image", "pos", "sensors"})):
                key = f"head/{name}/{tag}"
                if not want(key):
                    continue
                model = _prep(resolve(name)(modalities=mods), 12, dev)
                _, obs, _ = synthetic_trajectories(1, 3, sd, seed=22)
                particles = torch.from_numpy(
                    np.random.default_rng(32).standard_normal((3, 5, sd)).astype(np.float32)
                ).to(dev)
                out[key] = _np(model(states=particles, observations=_obs_at(obs, 0, dev)))

        # ---- R5: fused particle log-likelihoods -------------------------------------------------
        for name, sd in (
            ("PushCrossmodalParticleFilter", 2),
            ("DoorCrossmodalParticleFilter", 3),
            ("PushUnimodalParticleFilter", 2),
            ("PushCrossmodalParticleFilterSeq5", 2),
        ):
            if not want("fuse/" + name):
                continue
            filt = _prep(resolve(name)(), 13, dev)
            mm = filt.measurement_model
            _, obs, _ = synthetic_trajectories(1, 4, sd, seed=23, blackout_fraction=0.5)
            particles = torch.from_numpy(
                np.random.default_rng(33).standard_normal((4, 7, sd)).astype(np.float32)
            ).to(dev)
            for flags in ([True, True], [True, False], [False, True]):
                if name.endswith("Seq5") and flags == [True, False]:
                    continue  # -inf everywhere on blacked-out rows (quirk Q3); nothing to compare
                mm.enabled_models = flags
                tag = "".join("1" if f else "0" for f in flags)
                out[f"fuse/{name}/{tag}"] = _np(mm(states=particles, observations=_obs_at(obs, 0, dev)))
            if mm.crossmodal_weight_model is not None:
                out[f"fuse/{name}/log_weights"] = _np(
                    mm.crossmodal_weight_model(observations=_obs_at(obs, 0, dev))
                )

        # ---- R9: virtual sensors ------------------------------------------------------------------
        for name, sd in (("PushVirtualSensorModel", 2), ("DoorVirtualSensorModel", 3)):
            for tag, mods in (("image", {"image"}), ("force", {"pos", "sensors"})):
                key = f"vsensor/{name}/{tag}"
                if not want(key):
                    continue
                model = _prep(resolve(name)(modalities=mods), 14, dev)
                _, obs, _ = synthetic_trajectories(1, 6, sd, seed=24)
                z, tril = model(observations=_obs_at(obs, 0, dev))
                out[key + "/z"] = _np(z)
                out[key + "/tril"] = _np(tril)

        # ---- R8 + R10 + R11 + R12: Kalman filters over a short sequence -------------------------
        kf_cases = (
            ("DoorCrossmodalKalmanFilter", 3, {}, None),
            ("DoorCrossmodalKalmanFilter", 3, {"know_image_blackout": True}, None),
            ("DoorCrossmodalKalmanFilter", 3, {}, [False, True]),
            ("PushCrossmodalKalmanFilter", 2, {}, None),
            ("DoorUnimodalKalmanFilter", 3, {}, None),
            ("DoorKalmanFilter", 3, {}, None),
            ("DoorMeasurementCrossmodalKalmanFilter", 3, {}, None),
        )
        for name, sd, kwargs, flags in kf_cases:
            tag = name + ("+blackout" if kwargs else "") + ("+" + "".join("1" if f else "0" for f in flags) if flags else "")
            if not want("kf/" + tag):
                continue
            filt = _prep(resolve(name)(**kwargs), 15, dev)
            if flags is not None:
                filt.enabled_models = flags
            T, N = 5, 6
            states, obs, controls = synthetic_trajectories(T, N, sd, seed=25, blackout_fraction=0.3 if kwargs else 0.0)
            filt.initialize_beliefs(mean=states[0].to(dev), covariance=_cov(N, sd, dev))
            est = filt.forward_loop(
                observations={k: v[1:].to(dev) for k, v in obs.items()}, controls=controls[1:].to(dev)
            )
            out[f"kf/{tag}/estimates"] = _np(est)
            cov = getattr(filt, "weighted_covariances", None)
            if cov is None:
                cov = getattr(filt, "_belief_covariance", None)
            if cov is not None:
                out[f"kf/{tag}/covariance"] = _np(cov)
            if hasattr(filt, "crossmodal_weight_model"):
                out[f"kf/{tag}/beta"] = _np(filt.crossmodal_weight_model(observations=_obs_at(obs, 1, dev)))

        if want("kf/measurement_init"):
            filt = _prep(resolve("DoorCrossmodalKalmanFilter")(), 16, dev)
            _, obs, _ = synthetic_trajectories(1, 5, 3, seed=26)
            filt.measurement_initialize_beliefs(_obs_at(obs, 0, dev))
            out["kf/measurement_init/mean"] = _np(filt.filter_models[0].belief_mean)
            out["kf/measurement_init/covariance"] = _np(filt.filter_models[0].belief_covariance)

        # ---- R1 + R2 + R6 + R7 through torch's own RNG (only meaningful on the oracle torchfilter) -
        if include_rng_cases:
            for name, sd, train in (
                ("PushCrossmodalParticleFilter", 2, False),
                ("PushCrossmodalParticleFilter", 2, True),
                ("DoorCrossmodalParticleFilter", 3, False),
                ("PushUnimodalParticleFilter", 2, False),
            ):
                tag = f"pf_rng/{name}/{'train' if train else 'eval'}"
                if not want(tag):
                    continue
                filt = _prep(resolve(name)(), 17, dev, train=train)
                filt.num_particles = 30  # quirk Q8: set after .eval()
                T, N = 6, 4
                states, obs, controls = synthetic_trajectories(T, N, sd, seed=27)
                torch.manual_seed(1234)
                filt.initialize_beliefs(mean=states[0].to(dev), covariance=_cov(N, sd, dev))
                est = filt.forward_loop(
                    observations={k: v[1:].to(dev) for k, v in obs.items()}, controls=controls[1:].to(dev)
                )
                out[tag + "/estimates"] = _np(est)
                out[tag + "/particle_states"] = _np(filt.particle_states)
                out[tag + "/particle_log_weights"] = _np(filt.particle_log_weights)
    return out
