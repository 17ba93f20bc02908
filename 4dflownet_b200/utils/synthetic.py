"""Synthetic inputs for benchmarks and profiling runs: the 11-tuple one training batch carries
(PatchHandler3D.py:78-81) with the value ranges of the reference's normalised data -- velocities U(-1, 1),
magnitudes U(0, 0.016) (= 65/4095), HR targets N(0, 0.08^2) inside a ~12 % Bernoulli fluid mask, venc 1.5."""
import numpy as np


def synthetic_batch(B, patch_size, res_increase, seed=0):
    g = np.random.default_rng(seed)
    P, H = patch_size, patch_size * res_increase
    lr = [g.uniform(-1, 1, size=(B, P, P, P, 1)).astype(np.float32) for _ in range(3)]
    mg = [g.uniform(0, 0.016, size=(B, P, P, P, 1)).astype(np.float32) for _ in range(3)]
    mask = (g.uniform(size=(B, H, H, H)) < 0.12).astype(np.float32)
    hr = [(g.standard_normal((B, H, H, H, 1)).astype(np.float32) * 0.08) * mask[..., None] for _ in range(3)]
    venc = np.full((B,), 1.5, dtype=np.float32)
    return (*lr, *mg, *hr, venc, mask)
