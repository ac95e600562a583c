"""Constants of the reference (geepee/config.py:11-16) plus the precision switch."""
JITTER = 1e-5
GH_DEGREE = 10
PROP_MM = 'MM'
PROP_LIN = 'LIN'
PROP_MC = 'MC'
MC_NO_SAMPLES = 5

# arithmetic of the per-row kernels: 'fp64' (reference arithmetic, 1e-6 parity) or
# 'fp32' ("fp32-psi mode", 1e-3 parity).  The M x M tail is always fp64.
DEFAULT_PREC = 'fp64'

# replicated M x M tails: capture each phase in a CUDA graph after this many eager calls
# (tailgraph.py); GPB_TAIL_GRAPHS=0 in the environment disables capture
TAIL_GRAPHS = True
TAIL_GRAPH_WARMUP = 2
