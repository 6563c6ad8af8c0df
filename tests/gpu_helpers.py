"""Builders for the GPU side of the parity tests."""
import numpy as np
import torch

from cyclistsocialforce_b200 import parameters as P
from cyclistsocialforce_b200.engine import AgentGroup, Engine

PARAMS = dict(twod=P.InvPendulumBicycleParameters, invpendulum=P.InvPendulumBicycleParameters,
              balancingrider=P.BalancingRiderBicycleParameters, planarpoint=P.PlanarPointBicycleParameters,
              bicycle=P.BicycleParameters)


def make_engine(model, s0, vd, dests, dtype=torch.float64, **kw):
    """dests: per agent (Q,2|3) arrays WITHOUT the start entry (like Vehicle.setDestinations)."""
    s0 = np.asarray(s0, float)
    queues = []
    for k in range(s0.shape[0]):
        d = np.asarray(dests[k], float)
        if d.shape[1] == 2:
            d = np.c_[d, np.zeros(len(d))]
        queues.append(np.vstack([[s0[k, 0], s0[k, 1], 0.0], d]))
    g = AgentGroup(model, s0, PARAMS[model](), vd_default=vd, destqueues=queues, dtype=dtype)
    return Engine([g], dtype=dtype, **kw), g
