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

# Deterministic layer of the single-layer models: the forward saves Kfu[n,MP] and T[n,Do,MP] for the
# backward ((1 + Dout) MP sizeof(prec) bytes per row).  Rows are processed in chunks -- forward,
# likelihood, backward per chunk, additive statistics accumulated -- so that the saved buffers never
# exceed this many bytes: capacity is then bounded by the resident training data (8 (D + Dout) bytes
# per row), not by the saved tiles (82 GB at N = 1e7, M = 512 without chunking).
DET_SAVE_BYTES = 8 << 30
DET_MIN_CHUNK_ROWS = 4096

# fp32-psi mode: deterministic-layer forward on the 5th-generation tensor cores (tcgen05, 3xTF32; csrc/gpb_umma.cuh)
# instead of the SIMT fp32 kernel
DET_FP32_TENSOR_CORES = True

# replicated M x M tails: capture each phase in a CUDA graph after this many eager calls
# (tailgraph.py); GPB_TAIL_GRAPHS=0 in the environment disables capture
TAIL_GRAPHS = True
TAIL_GRAPH_WARMUP = 2
