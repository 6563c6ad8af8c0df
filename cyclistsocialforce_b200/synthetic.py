"""Seeded synthetic open-plane crowds (SURVEY.md section 8d recipe) for benchmarks and smoke
runs: side L = spacing*sqrt(N); x, y ~ U(0, L); psi ~ U(-pi, pi); v = v0; a queue of
``n_dest`` destinations every ``dest_step`` metres along psi + U(-0.5, 0.5), stop = 0."""
import math

import numpy as np


def synthetic_crowd(n, seed=1, spacing=4.0, n_dest=5, dest_step=60.0, v0=5.0, n_states=5):
    rng = np.random.default_rng(seed)
    L = spacing * math.sqrt(n)
    x = rng.uniform(0, L, n)
    y = rng.uniform(0, L, n)
    psi = rng.uniform(-np.pi, np.pi, n)
    s0 = np.zeros((n, n_states))
    s0[:, 0], s0[:, 1], s0[:, 2], s0[:, 3] = x, y, psi, v0
    a = psi + rng.uniform(-0.5, 0.5, n)
    d = dest_step * np.arange(1, n_dest + 1)
    q = np.zeros((n, n_dest, 3))
    q[:, :, 0] = x[:, None] + d[None, :] * np.cos(a)[:, None]
    q[:, :, 1] = y[:, None] + d[None, :] * np.sin(a)[:, None]
    return s0, q


def queues_with_start(s0, q):
    """Destination queues in the reference convention: entry 0 = start position
    (reference vehicle.py:183-185), then the destinations."""
    n = s0.shape[0]
    start = np.zeros((n, 1, 3))
    start[:, 0, 0], start[:, 0, 1] = s0[:, 0], s0[:, 1]
    return np.concatenate([start, q], axis=1)


def spatial_order(x, y, bits=16):
    """Permutation that sorts positions along a Hilbert curve (same curve as the kernels' keys).
    Sharding a crowd by contiguous ranges of this order is a spatial domain decomposition: every
    rank's agents are compact, so its target blocks cull most source chunks."""
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    lo = min(x.min(), y.min())
    span = max(x.max(), y.max()) - lo
    n = 1 << bits
    xi = np.minimum(((x - lo) / max(span, 1e-300) * (n - 1)).astype(np.int64), n - 1)
    yi = np.minimum(((y - lo) / max(span, 1e-300) * (n - 1)).astype(np.int64), n - 1)
    d = np.zeros_like(xi)
    s = n >> 1
    while s > 0:
        rx = ((xi & s) > 0).astype(np.int64)
        ry = ((yi & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        flip = (ry == 0) & (rx == 1)
        xi = np.where(flip, n - 1 - xi, xi)
        yi = np.where(flip, n - 1 - yi, yi)
        swap = ry == 0
        xi, yi = np.where(swap, yi, xi), np.where(swap, xi, yi)
        s >>= 1
    return np.argsort(d, kind="stable")
