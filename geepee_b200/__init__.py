"""geepee_b200 -- B200-native (sm_100a CUDA) hot path of thangbui/geepee.

Scope: the per-minibatch AEP / VFE energy-and-gradient evaluation of the sparse-GP
layer family (SGPR, SDGPR, SGPLVM, SGPSSM) behind the reference's Python model API.
See DESIGN.md.  There is no CPU fallback: the ops need libgeepee_b200.so and a GPU.
"""
from .config import JITTER, PROP_MM, PROP_LIN, PROP_MC, MC_NO_SAMPLES, GH_DEGREE  # noqa: F401

__version__ = '0.1.0'
