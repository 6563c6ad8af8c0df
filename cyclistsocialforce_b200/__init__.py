"""cyclistsocialforce_b200 -- B200-native stepping engine behind the
cyclistsocialforce vehicle / intersection / scenario API.

The hot path (``SocialForceIntersection.step()``) runs as hand-written sm_100a
CUDA kernels reached through the C ABI in ``include/csf_b200.h``; the Python
classes here mirror the reference's interface for that path
(``vehicle.TwoDBicycle`` ..., ``intersection.SocialForceIntersection``,
``scenario.Scenario``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
