"""hop_b200 -- the trajectory hot path of Becksteinlab/hop on B200 (sm_100a).

density histogram -> hydration-site map -> hopping trajectory -> site-to-site transitions, behind
hop's own Python classes (DensityCreator, Density, HoppingTrajectory, TransportNetwork).  Python is
host code only; every stage runs as hand-written CUDA kernels from libhopb200.so through a C ABI
(include/hopb200.h).  There is no CPU fallback.
"""
from .constants import SITELABEL  # noqa: F401

__version__ = "0.1.0"
__all__ = ["constants", "density", "sitemap", "trajectory", "graph", "interactive", "universe", "kernels", "pipeline"]
