"""Pins for the restated torchfilter recursion: the explicit
noise recipes are bit-identical to the torch.distributions calls torchfilter makes, so
"identical uniform draws" is a well-defined thing to hand to the CUDA path."""
import math

import numpy as np
import torch

from oracle import crossmodal_port as port
from oracle.noise import RecordedNoise, TorchRNGNoise
from multimodalfilter_b200.synthetic import fill_parameters, synthetic_trajectories
from torchfilter.filters import multinomial_inverse_cdf


def test_mvn_sample_recipe_bit_exact():
    torch.manual_seed(3)
    mean = torch.randn(5, 3)
    A = torch.randn(5, 3, 3)
    cov = A @ A.transpose(-1, -2) + 0.5 * torch.eye(3)
    torch.manual_seed(9)
    ref = torch.distributions.MultivariateNormal(mean, cov).sample((7,))
    torch.manual_seed(9)
    eps = torch.randn(7, 5, 3)
    mine = mean[None] + (torch.linalg.cholesky(cov)[None] @ eps[..., None]).squeeze(-1)
    assert torch.equal(ref, mine)


def test_mvn_rsample_recipe_bit_exact():
    torch.manual_seed(4)
    loc = torch.randn(40, 2)
    tril = torch.linalg.cholesky(torch.diag(torch.tensor([0.02, 0.02])))[None].expand(40, 2, 2)
    torch.manual_seed(10)
    ref = torch.distributions.MultivariateNormal(loc=loc, scale_tril=tril).rsample()
    torch.manual_seed(10)
    mine = loc + (tril @ torch.randn(40, 2)[..., None]).squeeze(-1)
    assert torch.equal(ref, mine)


def test_categorical_sample_is_sequential_fp32_inverse_cdf():
    """torch.multinomial on CPU == lower bound of float64 uniforms on a sequential fp32 CDF."""
    mismatches = 0
    total = 0
    for seed, (N, M, S) in enumerate([(256, 30, 30), (64, 300, 300), (80, 1000, 1000), (3, 37, 101)]):
        torch.manual_seed(100 + seed)
        logits = torch.randn(N, M) * 3.0
        logits = logits - torch.logsumexp(logits, dim=1, keepdim=True)
        torch.manual_seed(200 + seed)
        ref = torch.distributions.Categorical(logits=logits).sample((S,)).T
        torch.manual_seed(200 + seed)
        u = torch.rand(N * S, dtype=torch.float64).reshape(N, S)
        probs = torch.softmax(logits - torch.logsumexp(logits, dim=-1, keepdim=True), dim=-1)
        mine = multinomial_inverse_cdf(probs, u)
        mismatches += int((ref != mine).sum())
        total += ref.numel()
    assert total > 100000
    assert mismatches == 0, f"{mismatches}/{total}"


def test_fp64_accumulated_cdf_is_not_what_torch_does():
    """Negative control: the sequential-fp32 detail matters (guards against a lazy oracle)."""
    torch.manual_seed(5)
    N, M = 64, 1000
    probs = torch.softmax(torch.randn(N, M) * 3.0, dim=-1)
    torch.manual_seed(6)
    ref = torch.multinomial(probs, M, replacement=True)
    torch.manual_seed(6)
    u = torch.rand(N * M, dtype=torch.float64).reshape(N, M).numpy()
    cdf = np.cumsum(probs.numpy().astype(np.float64), axis=1)
    cdf /= cdf[:, -1:]
    idx = np.stack([np.searchsorted(cdf[n], u[n], side="left") for n in range(N)])
    assert (idx != ref.numpy()).sum() > 0


def test_filter_with_injected_noise_equals_distribution_path():
    """Whole recursion: noise=None (torch.distributions, as torchfilter does) vs explicit draws."""
    T, N, sd = 6, 5, 2
    states, obs, controls = synthetic_trajectories(T, N, sd, seed=41)
    cov = (torch.eye(sd) * 0.1)[None].expand(N, sd, sd)
    results = []
    for noise in (None, TorchRNGNoise()):
        f = fill_parameters(port.PushCrossmodalParticleFilter(), seed=2)
        f.eval()
        f.num_particles = 30
        f.noise = noise
        torch.manual_seed(77)
        with torch.no_grad():
            f.initialize_beliefs(mean=states[0], covariance=cov)
            est = f.forward_loop(observations={k: v[1:] for k, v in obs.items()}, controls=controls[1:])
        results.append((est, f.particle_states, f.particle_log_weights))
    for a, b in zip(*results):
        assert torch.equal(a, b)


def test_recorded_noise_replays_torch_rng_noise():
    T, N, sd, M = 4, 3, 3, 30
    states, obs, controls = synthetic_trajectories(T, N, sd, seed=42)
    cov = (torch.eye(sd) * 0.1)[None].expand(N, sd, sd)

    def run(noise):
        f = fill_parameters(port.DoorCrossmodalParticleFilter(), seed=3)
        f.eval()
        f.num_particles = M
        f.noise = noise
        torch.manual_seed(5)  # after construction: nn.Linear's default init consumes the stream
        with torch.no_grad():
            f.initialize_beliefs(mean=states[0], covariance=cov)
            return f.forward_loop(observations={k: v[1:] for k, v in obs.items()}, controls=controls[1:])

    a = run(TorchRNGNoise())
    torch.manual_seed(5)
    init = torch.randn(M, N, sd)
    eps, us = [], []
    for _ in range(T - 1):
        eps.append(torch.randn(N * M, sd))
        us.append(torch.rand(N * M, dtype=torch.float64).reshape(N, M))
    b = run(RecordedNoise(init_eps=init, process_eps=eps, uniforms=us))
    assert torch.equal(a, b)


def test_systematic_mode_is_low_variance():
    N, M = 4, 64
    probs = torch.softmax(torch.randn(N, M), dim=-1)
    noise = RecordedNoise(uniforms=[torch.full((N,), 0.5, dtype=torch.float64)], mode="systematic")
    idx = noise.resample_indices(probs, M)
    assert idx.shape == (N, M)
    assert (idx[:, 1:] >= idx[:, :-1]).all()  # sorted by construction
    counts = torch.stack([torch.bincount(idx[n], minlength=M) for n in range(N)]).double()
    assert (counts - probs.double() * M).abs().max() <= 1.0 + 1e-6
