/*
 * geepee_b200.h -- C ABI of the B200-native hot path of thangbui/geepee.
 *
 * Drop-in boundary.  geepee is a pure-Python/numpy library; the only foreign
 * function interface on its hot path is `weave.inline(...)` in
 * geepee/kernels.py:236-239, which compiles the C++ loop of
 * `compute_psi_weave` (kernels.py:181-240) at run time.  `gpb_psi_stats` is the
 * 1:1 replacement of that routine.  Every other entry point replaces one
 * numpy/scipy contraction group of the SGP_Layer forward / backward
 * (file:line cited per function) so that the N x M (Kfu) and N x M x M (psi2)
 * intermediates the reference materialises are produced tile by tile on chip.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory (HBM) unless
 *     the name starts with `h_`; all arrays are C-contiguous row-major.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     every call is asynchronous and stream ordered, none synchronises.
 *   - `prec`: GPB_F64 ("fp64 mode", reference arithmetic) or GPB_F32 ("fp32-psi
 *     mode": kernel / psi generation and tile contractions in fp32; every
 *     cross-row accumulator and every interface array stays fp64).
 *   - hyper-parameters are passed as the reference stores them: `ls[D]` = log
 *     lengthscales, `sf[1]` = log signal std, `sn[1]` = log noise std
 *     (base_models.py:630-658), on the device.
 *   - return value 0 = ok, otherwise a negative error code; gpb_last_error()
 *     gives the message.  No entry point has a CPU fallback.
 *   - workspaces: `gpb_*_ws_bytes` returns the scratch size an op needs; the
 *     caller owns the buffer (256-byte aligned).
 */
#ifndef GEEPEE_B200_H
#define GEEPEE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPB_F64 0
#define GPB_F32 1

#define GPB_OK 0
#define GPB_ERR_ARG -1   /* unsupported size / null pointer */
#define GPB_ERR_CUDA -2  /* CUDA runtime error (launch, attribute, ...) */
#define GPB_ERR_WS -3    /* workspace too small */

int gpb_version(void);
const char* gpb_last_error(void);
/* number of SMs the launch heuristics size grids for (148 on B200) */
int gpb_sm_count(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
long gpb_launch_count(void);

/* ---- a1: ARD-SE kernel matrix.  kernels.py:10-22 compute_kernel(2*ls, 2*sf, x, z);
 *      with jitter != 0 and n == M it is Kuu of base_models.py:454-464. */
int gpb_kmat(const double* x, const double* z, const double* ls, const double* sf,
             int n, int M, int D, double jitter, double* out /*[n,M]*/, void* stream);

/* ---- a2: psi statistics, materialised.  kernels.py:181-240 compute_psi_weave(2*ls, 2*sf,
 *      mx, vx, z) -> psi1[n,M], psi2[n,M,M].  The reference's weave FFI, one to one. */
int gpb_psi_stats(const double* mx, const double* vx, const double* z, const double* ls,
                  const double* sf, int n, int M, int Q, double* psi1, double* psi2, void* stream);

/* ---- a7 / a7': Gaussian likelihood.  mode 0: lik_layers.py:104-133 compute_log_Z (+ the dv sum of
 *      backprop_grads 154-181); mode 1: lik_layers.py:183-199 compute_log_lik_exp (+ the sum of
 *      backprop_grads_log_lik_exp 217-226).  dm, dv come out multiplied by `scale`.
 *      out2[0] = sum of log-terms (unscaled), out2[1] = mode 0: sum of unscaled dv; mode 1: dsn sum. */
size_t gpb_gauss_lik_ws_bytes(long total);
int gpb_gauss_lik(const double* m, const double* v, const double* y, const double* sn,
                  double alpha, double scale, long total, int mode, double* dm, double* dv,
                  double* out2, void* ws, size_t ws_bytes, void* stream);

/* ---- a3 / a4 / a10: batched inverse + log-determinant of SPD matrices A[batch,M,M] (M <= 512), fp64.
 *      Replaces np.linalg.inv / np.linalg.slogdet of base_models.py:464,471,476 and
 *      aep_models.py:68,78,91,525,533.  One thread-block cluster per matrix, blocked Gauss-Jordan. */
int gpb_spd_inverse(const double* A, int batch, int M, double* Ainv, double* logdet, void* stream);

/* ---- (8f rank 1) Probit likelihood, y in {-1,+1}.  mode 0: lik_layers.py:303-362 Probit_Layer.compute_log_Z
 *      (alpha == 1: closed form; otherwise Gauss-Hermite quadrature with the nodes / weights of
 *      numpy.polynomial.hermite.hermgauss(ngh), ngh <= 64, passed in device arrays);
 *      mode 1: lik_layers.py:418-436 compute_log_lik_exp.  dm, dv come out multiplied by `scale`;
 *      out2[0] = sum of the log terms (unscaled).  Workspace: gpb_gauss_lik_ws_bytes(total). */
int gpb_probit_lik(const double* m, const double* v, const double* y, const double* gh_x,
                   const double* gh_w, int ngh, double alpha, double scale, long total, int mode,
                   double* dm, double* dv, double* out2, void* ws, size_t ws_bytes, void* stream);

/* ---- a14: linear-Gaussian emission, tilted.  lik_layers.py:573-627 Gauss_Emis.compute_emission_tilted:
 *      y ~ N(C x, diag(R)), per row Vy = diag(R/alpha) + C diag(vx) C^T.  R[Do] are VARIANCES.
 *      dmx, dvx[n,Q] come out multiplied by `scale`;
 *      out[2 + Do + Do*Q] = [ sum_n -(y-Cm)^T Vy^-1 (y-Cm)/2 | sum_n log|Vy| |
 *                             sum_n (-Vy^-1[a,a]/2 + w_a^2/2) | dC1 + dC2 of lines 612-615 ] (unscaled).
 *      Do <= 8 and Q <= 8 (GPB_ERR_ARG otherwise). */
size_t gpb_gauss_emis_ws_bytes(int n, int Do, int Q);
int gpb_gauss_emis(const double* mx, const double* vx, const double* y, const double* C, const double* R,
                   double alpha, double scale, int n, int Q, int Do, double* dmx, double* dvx,
                   double* out, void* ws, size_t ws_bytes, void* stream);

/* ---- deterministic-input layer -------------------------------------------------------------- */
/* M padded to the GEMM tile (128, 256 or 512); -1 if M > 512. */
int gpb_det_pad_m(int M);
/* element size of `prec` (8 or 4) */
int gpb_prec_bytes(int prec);
/* zero-padded, `prec`-typed copies of A[Do,M] -> Ap[Do,MP] and B[Do,M,M] -> Bp[Do,MP,MP] */
int gpb_det_pad_operands(int prec, const double* A, const double* B, int M, int Do,
                         void* Ap, void* Bp, void* stream);
/* a5: aep_models.py:142-158 _forward_prop_deterministic_thru_cav (post twin base_models.py:265-284):
 *      mout = kfu A^T, vout = sf2 + sum_ab B[d,a,b] kfu_a kfu_b, kfu generated on chip.
 *      Optionally saves kfu[n,MP] and T[n,Do,MP] = B_d kfu (typed by prec) for the backward. */
int gpb_det_fwd(int prec, const double* x, const double* z, const double* ls, const double* sf,
                const void* Ap, const void* Bp, int n, int M, int D, int Do,
                double* mout, double* vout, void* Ksave, void* Tsave, void* stream);
/* a8 (per-row part): aep_models.py:452-460,490 + kernels.py:381-399 (kfucompDer), vfe twin
 *      vfe_models.py:498-506.  dm, dv are the SCALED upstream gradients.
 *      -> dA[Do,M] = sum_n dm kfu ; dzu[M,D] ; dl[D] (wrt lengthscale) ; dsf2[1] (wrt variance) */
size_t gpb_det_bwd_ws_bytes(int n, int M, int D, int Do);
int gpb_det_bwd(int prec, const double* x, const double* z, const double* ls, const double* sf,
                const void* Ap, const double* dm, const double* dv, const void* Ksave,
                const void* Tsave, int n, int M, int D, int Do, double* dA, double* dzu,
                double* dl, double* dsf2, void* ws, size_t ws_bytes, void* stream);
/* Monte-Carlo propagation: gradient wrt the layer INPUT of the deterministic layer,
 *      aep_models.py:346-350 (backprop_grads_lvm_mc) + kernels.py:393-395 (kfucompDer, grad_x=True):
 *      dx[n,D] = sum_m (dm A + 2 dv T)[n,m] kfu[n,m] (z[m,:] - x[n,:]) / l^2, from the buffers
 *      gpb_det_fwd saved. */
int gpb_det_dx(int prec, const double* x, const double* z, const double* ls, const void* Ap,
               const double* dm, const double* dv, const void* Ksave, const void* Tsave,
               int n, int M, int D, int Do, double* dx, void* stream);
/* a8 (rank update): aep_models.py:493  dB[Do,M,M] = sum_n dv[n,d] kfu kfu^T */
size_t gpb_det_syrk_ws_bytes(int n, int M, int Do);
int gpb_det_syrk(int prec, const void* Ksave, const double* dv, int n, int M, int Do,
                 double* dB, void* ws, size_t ws_bytes, void* stream);

/* ---- moment-matched (uncertain-input) layer ------------------------------------------------- */
/* a6: aep_models.py:183-199 _forward_prop_random_thru_cav_mm (post twin base_models.py:286-307):
 *      mout = psi1 A^T, vout = sf2 + sum_ab B[d,a,b] psi2[n,a,b] - mout^2; psi2 never stored.
 *      Also returns vacc[n,Do] = sum_ab B[d,a,b] psi2[n,a,b] and (if psi1save != NULL) psi1[n,M],
 *      which the backward reuses instead of re-evaluating them. */
size_t gpb_mm_ws_bytes(int n, int M, int Q, int Do, int backward);
int gpb_mm_fwd(int prec, const double* mx, const double* vx, const double* z, const double* ls,
               const double* sf, const double* A, const double* B, int n, int M, int Q, int Do,
               double* mout, double* vout, double* vacc, double* psi1save, void* ws, size_t ws_bytes,
               void* stream);
/* a9 (per-row part): aep_models.py:238-250 + kernels.py:302-309,355-378,402-444
 *      (compute_psi_derivatives), vfe twin vfe_models.py:351-361.
 *      dm, dv scaled upstream gradients; mout, vacc, psi1 from gpb_mm_fwd on the same inputs and B.
 *      -> dA[Do,M] = sum_n dm_all psi1 ; dB[Do,M,M] = sum_n dv psi2 ; dzu[M,Q] ; dl[Q] ; dsf2[1] ;
 *         dvsum[1] = sum dv ; dmx[n,Q], dvx[n,Q] */
int gpb_mm_bwd(int prec, const double* mx, const double* vx, const double* z, const double* ls,
               const double* sf, const double* A, const double* B, const double* dm,
               const double* dv, const double* mout, const double* vacc, const double* psi1,
               int n, int M, int Q, int Do,
               double* dA, double* dB, double* dzu, double* dl, double* dsf2, double* dvsum,
               double* dmx, double* dvx, void* ws, size_t ws_bytes, void* stream);

/* ---- per-kernel device timing for bench.py's roofline (CUDA events on the launching stream).
 *      slots: 0 det_fwd, 1 det_bwd, 2 det_syrk, 3 mm_pairs(fwd), 4 mm_pairs(bwd), 5 mm_rows_bwd,
 *      6 mm_cols_bwd, 7 mm_psi1_fwd.  collect() synchronises the recorded events, returns the summed
 *      milliseconds and launch counts per slot into HOST arrays of 8 and resets them. */
int gpb_profile_enable(int on);
int gpb_profile_collect(double* h_ms, long* h_count);

/* ---- microbenchmarks used by bench.py for the roofline denominators ------------------------- */
/* runs `iters` dependent-chain-free FMAs per thread on a full grid; returns total flops in
 * *h_flops (host pointer).  `sink` is a device buffer of >= 8*gpb_sm_count() doubles.
 * Time it with events on `stream`. */
int gpb_fma_peak(int prec, long iters, double* sink, double* h_flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEEPEE_B200_H */
