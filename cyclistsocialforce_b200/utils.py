"""Host-side angle / vector helpers with the reference's names and semantics
(reference src/cyclistsocialforce/utils.py:56-86, :114-227).  The device versions
live in csrc/csf_common.cuh; these serve user code and host-side set-up."""
import numpy as np


def limitMagnitude(x, y, r):
    """utils.py:56-86 (in place, like the reference)."""
    rin = np.sqrt(x ** 2 + y ** 2)
    ids = rin > r
    if np.any(rin):
        x[ids] = x[ids] * r[ids] / rin[ids]
        y[ids] = y[ids] * r[ids] / rin[ids]
    return x, y


def to_deg(rad):
    return 360 * rad / (2 * np.pi)


def to_rad(deg):
    return 2 * np.pi * deg / 360


def limitAngle(theta):
    """Wrap to (-pi, pi], utils.py:124-139."""
    if isinstance(theta, np.ndarray):
        theta = np.floor(theta / (2 * np.pi)) * (-2 * np.pi) + theta
        theta[theta > np.pi] = (theta - 2 * np.pi)[theta > np.pi]
        theta[theta < -np.pi] = (theta + 2 * np.pi)[theta < -np.pi]
        return theta
    theta = np.floor(theta / (2 * np.pi)) * (-2 * np.pi) + theta
    if theta > np.pi:
        theta = theta - 2 * np.pi
    elif theta < -np.pi:
        theta = theta + 2 * np.pi
    return theta


def expandAngle(theta):
    return 2 * np.pi + theta if theta < 0 else theta


def angleSUMOtoSFM(theta):
    return limitAngle((np.pi / 2) - to_rad(theta))


def angleSFMtoSUMO(theta):
    return to_deg(expandAngle((np.pi / 2) - theta))


def angleDifference(a1, a2):
    """Signed shortest rotation from a1 to a2, utils.py:151-182."""
    if isinstance(a1, np.ndarray):
        da = np.where(a1 > a2, a1 - a2, a2 - a1)
        da = np.where(da > np.pi, (2 * np.pi) - da, da)
        t1 = np.abs(limitAngle(a1 - da) - a2)
        t2 = np.abs(limitAngle(a1 + da) - a2)
        return np.where(t1 < t2, -da, da)
    da = a1 - a2 if a1 > a2 else a2 - a1
    if da > np.pi:
        da = (2 * np.pi) - da
    t1 = abs(limitAngle(a1 - da) - a2)
    t2 = abs(limitAngle(a1 + da) - a2)
    return -da if t1 < t2 else da


def cart2polar(x, y):
    rho = np.sqrt(np.power(x, 2) + np.power(y, 2))
    phi = np.array(np.arccos(x / rho))
    phi[y < 0] = -phi[y < 0]
    return rho, phi


def polar2cart(rho, phi):
    return rho * np.cos(phi), rho * np.sin(phi)


def thresh(x, minmax):
    assert minmax[0] <= minmax[1], (
        f"Minimum must be smaller then the maximum! Instead it was [{minmax[0]}, {minmax[1]}]")
    return np.maximum(np.minimum(x, minmax[1]), minmax[0])
