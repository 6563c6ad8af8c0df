"""The five helpers of the reference's ``utils`` module that the stepping path uses, for host code that
talks to this package the way it talks to the reference (same names, argument order and results):
``limitAngle``, ``angleDifference``, ``cart2polar``, ``thresh``, ``limitMagnitude``
(reference src/cyclistsocialforce/utils.py:56-86, :124-227).

On the stepping path itself these run on the device (csrc/csf_common.cuh: ``limit_angle``,
``angle_difference``, ``clampT``; the trig-free pair kernel needs no ``cart2polar``).  The host versions
below are written for numpy broadcasting: scalars and arrays take the same branch-free path.
The reference's SUMO angle converters, plotting and FIFO helpers belong to paths that are out of scope.
"""
import numpy as np

_TWO_PI = 2.0 * np.pi


def _scalar_like(out, *inputs):
    """Python float for all-scalar input (the reference returns scalars there), ndarray otherwise."""
    return out if any(isinstance(a, np.ndarray) for a in inputs) else float(out)


def limitAngle(theta):
    """Angle in (-pi, pi]: one period shift by floor division, then the two half-open ends folded back
    (exactly the reference's arithmetic, so wrapped values agree to the bit)."""
    t = np.asarray(theta, dtype=float)
    t = t - _TWO_PI * np.floor(t / _TWO_PI)
    t = t - _TWO_PI * (t > np.pi) + _TWO_PI * (t < -np.pi)
    return _scalar_like(t, theta)


def angleDifference(a1, a2):
    """Signed shortest rotation that takes ``a1`` to ``a2``: magnitude min(|a1 - a2|, 2 pi - |a1 - a2|),
    negative iff stepping back from a1 lands closer to a2 than stepping forward (a tie counts as
    forward)."""
    p, q = np.asarray(a1, dtype=float), np.asarray(a2, dtype=float)
    gap = np.abs(p - q)
    gap = np.where(gap > np.pi, _TWO_PI - gap, gap)
    back = np.abs(limitAngle(p - gap) - q)
    fwd = np.abs(limitAngle(p + gap) - q)
    return _scalar_like(np.where(back < fwd, -gap, gap), a1, a2)


def cart2polar(x, y):
    """(rho, phi) with phi = +-arccos(x / rho), the sign taken from y (y == 0 counts as +)."""
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    rho = np.sqrt(x * x + y * y)          # (not hypot: the reference's rounding)
    phi = np.where(y < 0, -1.0, 1.0) * np.arccos(x / rho)
    return rho, phi


def thresh(x, minmax):
    """Clamp ``x`` to [minmax[0], minmax[1]]."""
    lo, hi = minmax[0], minmax[1]
    if lo > hi:
        raise AssertionError(f"Minimum must be smaller then the maximum! Instead it was [{lo}, {hi}]")
    return np.clip(x, lo, hi)


def limitMagnitude(x, y, r):
    """Scale the vectors (x[i], y[i]) that are longer than r[i] down to length r[i] -- IN PLACE, as the
    reference does (its callers rely on that); all-zero input is left untouched."""
    norm = np.sqrt(x * x + y * y)
    long_ = norm > r
    if norm.any():
        x[long_] = x[long_] * r[long_] / norm[long_]
        y[long_] = y[long_] * r[long_] / norm[long_]
    return x, y
