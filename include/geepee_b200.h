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
 *      aep_models.py:68,78,91,525,533.  One thread-block cluster (8 CTAs) per matrix, blocked Gauss-Jordan;
 *      M <= 256: the matrix lives in FP64 tensor-core accumulator tiles in registers. */
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
/*      lik_layers.py:600-627 from those sums: fin = [scale*logZ | 0 | scale*dR[Do] (wrt the log-sqrt parameter) |
 *      scale*dC[Do*Q]]; raw = the `out` of gpb_gauss_emis over Nb rows. */
int gpb_gauss_emis_finish(const double* raw, const double* R, double alpha, double scale, long Nb, int Do,
                          int Q, double* fin, void* stream);

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
/* a5 in fp32-psi mode on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 operand split, fp32
 *      accumulators in TMEM, B tiles through the TMA engine).  gpb_det_tc_prep turns the padded fp32 B operand and
 *      the pseudo-inputs into the layouts the kernel consumes (Bu: gpb_det_tc_bu_bytes, Zs: gpb_det_tc_zs_bytes);
 *      gpb_det_fwd_tc has the semantics of gpb_det_fwd(GPB_F32, ...) with mandatory save buffers.
 *      gpb_det_tc_available() is 0 in builds without the tensor-core path (the CPU emulator of the tests). */
int gpb_det_tc_available(void);
size_t gpb_det_tc_bu_bytes(int M, int Do);
size_t gpb_det_tc_zs_bytes(int M, int D);
int gpb_det_tc_prep(const void* Bp, const double* z, const double* ls, int M, int D, int Do, void* Bu, void* Zs,
                    void* stream);
int gpb_det_fwd_tc(const double* x, const double* ls, const double* sf, const void* Zs, const void* Ap,
                   const void* Bu, int n, int M, int D, int Do, double* mout, double* vout, void* Ksave,
                   void* Tsave, void* stream);
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
/* a8 (rank update): aep_models.py:493  dB[Do,M,M] = sum_n dv[n,d] kfu kfu^T
 *      (fp64: FP64 tensor cores; fp32 with M <= 256: tcgen05, 3xTF32, fp64 flush every 2048 rows) */
size_t gpb_det_syrk_ws_bytes(int n, int M, int Do);
int gpb_det_syrk(int prec, const void* Ksave, const double* dv, int n, int M, int Do,
                 double* dB, void* ws, size_t ws_bytes, void* stream);

/* ---- moment-matched (uncertain-input) layer ------------------------------------------------- */
/* a6: aep_models.py:183-199 _forward_prop_random_thru_cav_mm (post twin base_models.py:286-307):
 *      mout = psi1 A^T, vout = sf2 + sum_ab B[d,a,b] psi2[n,a,b] - mout^2; psi2 never stored.
 *      Also returns vacc[n,Do] = sum_ab B[d,a,b] psi2[n,a,b] and (if psi1save != NULL) psi1[n,M],
 *      which the backward reuses instead of re-evaluating them.
 *      (fp32, Do <= 4, Q <= 7: the pair exponents are a 3xTF32 tcgen05 GEMM of row against pair features.) */
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

/* ---- the replicated O(Dout M^3) tail (tail_pre / tail_post) -------------
 *      Everything the reference computes between the parameters and the per-row work
 *      (update_hypers / compute_kuu / update_posterior base_models.py:630-658,454-488, compute_cavity
 *      aep_models.py:513-546, compute_phi 62-114, compute_KL vfe_models.py:309-325) and between
 *      the reduced statistics and the parameter gradients (aep_models.py:252-297,462-511,548-586,
 *      base_models.py:490-516, vfe_models.py:363-394,518-541, d_trace_MKzz_dhypers kernels.py:447-475)
 *      is numpy einsum / linalg there.  Here a phase is a PROGRAM of batched fp64 primitives that the
 *      host side assembles once per layer shape and the library executes in order on `stream`
 *      (`gpb_tail_exec`; the matrix inverses are `gpb_spd_inverse`).  Matrices are row-major with
 *      leading dimension ld[i]; sstride[i] is the batch stride of operand i in elements (0 = shared
 *      by the whole batch); unused operands are NULL.
 *
 *      GPB_TOP_GEMM     dst[b] (m x n) = coef[0] op(src0[b]) (m x k) op(src1[b]) (k x n) + coef[1] src2[b]
 *                       flags bit0 / bit1: src0 / src1 stored transposed.  FP64 tensor cores (DMMA).
 *      GPB_TOP_LINCOMB  dst[b][i][j] = sum_{s<4} coef[s] src_s[b][i][j] (flags bit s: [j][i])
 *                       + coef[4] src4[b][i] src5[b][j] + coef[5] (i == j);  flags bit 8: summed over b;
 *                       flags bit 9: src2, src3 are a second outer pair (coef[2] src2[b][i] src3[b][j])
 *      GPB_TOP_MATVEC   dst[b][i] = coef[0] (op(src0[b]) src1[b])_i + coef[1] (op(src2[b]) src3[b])_i
 *                       + coef[2] src4[b][i] + coef[3] src5[b][i]   (m outputs, k inner; flags bit0/1: transposed)
 *      GPB_TOP_DOTS     dst[0] = (flags bit0 ? dst[0] : 0) + coef[6] + sum_{t<3} coef[t] sum_{e<ld[2t]}
 *                       src_{2t}[e] (src_{2t+1} ? src_{2t+1}[e] : 1)
 *      GPB_TOP_UNPACK_R dst[b] (m x m) = upper-triangular R from src0[b][m(m+1)/2], diagonal exponentiated
 *                       (base_models.py:645-653)
 *      GPB_TOP_PACK_R   dst[b][m(m+1)/2] = coef[0] triu(src0[b]) with the diagonal times diag(src1[b])
 *                       (base_models.py:505-514)
 *      GPB_TOP_KHYPER   kernels.py:447-475 folded with aep_models.py:455-460,497-504: src = {Mm, Kuu, zu,
 *                       ls, sf, stats = [dzu0[m*k] | dl[k] | dsf2 | dvsum]}; dst = coef[1] [dsf | dls[k] | dzu[m*k]]
 *                       followed by scratch: dst holds gpb_tail_khyper_out_len(m, k) doubles;
 *                       coef[0] = jitter; m = M, k = D <= 32 (two launches)
 *      GPB_TOP_SUM      dst[0] = sum of src0[0 .. sstride[0]) (any length; src1 = scratch of >= 1024 doubles) */
#define GPB_TOP_GEMM 1
#define GPB_TOP_LINCOMB 2
#define GPB_TOP_MATVEC 3
#define GPB_TOP_DOTS 4
#define GPB_TOP_UNPACK_R 5
#define GPB_TOP_PACK_R 6
#define GPB_TOP_KHYPER 7
#define GPB_TOP_SUM 8
typedef struct GpbTailOp {
    int kind, flags;
    int batch, m, n, k;
    const double* src[6];
    long sstride[6];
    int ld[6];
    double coef[8];
    double* dst;
    long dstride;
    int ldd;
} GpbTailOp;
long gpb_tail_khyper_out_len(int M, int D);
/* run `n_ops` tail primitives in order on `stream` (h_ops: HOST array; one kernel launch per op,
 * two for GPB_TOP_SUM) */
int gpb_tail_exec(const GpbTailOp* h_ops, int n_ops, void* stream);
/* dst = scale * concat(src_0[0..count_0), src_1[0..count_1), ...): the flat energy + gradient vector that
 * goes back to the optimiser with one copy (utils.py:68-90 flatten order is the caller's).
 * h_srcs / h_counts: HOST arrays of n device pointers / element counts. */
int gpb_tail_gather(int n, const double* const* h_srcs, const long* h_counts, double scale,
                    double* dst, void* stream);

/* dst_i[0..count_i) = src_i[0..count_i) for i < n in one launch (per 24 entries): refreshes the static
 * input buffers of a captured tail phase.  h_*: HOST arrays. */
int gpb_tail_copy(int n, const double* const* h_srcs, double* const* h_dsts, const long* h_counts,
                  void* stream);

/* Device-side address of pinned (mapped) host memory: lets gpb_tail_copy read a small per-step upload (the parameter
 * vector the optimiser hands to objective_function, utils.py:42-53) straight from host memory instead of queueing it on
 * the H2D copy engine behind a large input copy.  Error if the memory is not mapped. */
int gpb_host_device_ptr(const void* host_ptr, void** dev_ptr);

/* ---- a12 / a13: elementwise latent-variable algebra (one thread per (row, latent dim)) ----------
 *      SGPLVM: get_cavity_x aep_models.py:840-861, compute_phi_x 863-867, compute_cav_grad_x 817-838,
 *      get_posterior_x base_models.py:765-775, compute_posterior_grad_x 913-929; VFE twin vfe_models.py:749-845.
 *      x1, x2: the raw [N,Q] parameters on the device; sel: device row indices of the minibatch (NULL: rows
 *      lo .. lo+n-1); mode 0 = AEP (prior1/prior2 = prior natural parameters; output = cavity moments),
 *      mode 1 = VFE (prior1/prior2 = prior mean / variance; output = posterior moments). */
size_t gpb_latent_ws_bytes(long total);
int gpb_lvm_x_fwd(int mode, int nat, const double* x1, const double* x2, const long* sel, long lo, int n, int Q,
                  double prior1, double prior2, double alpha, double* m /*[n,Q]*/, double* v /*[n,Q]*/, void* stream);
/* backward: dmx, dvx[n,Q] from the layer -> gx1, gx2[N,Q] (rows outside the selection zeroed) and
 * sums[2] = {phi_x(cavity), phi_x(posterior)} (AEP) / {KL(q(x)||p(x)), 0} (VFE).  s_cav, s_post: the
 * scale factors of the two log-partition terms (AEP) / s_cav = N/n (VFE). */
int gpb_lvm_x_bwd(int mode, int nat, const double* x1, const double* x2, const long* sel, long lo, int n, long N,
                  int Q, double prior1, double prior2, double alpha, double s_cav, double s_post,
                  const double* dmx, const double* dvx, double* gx1, double* gx2, double* sums, void* ws,
                  size_t ws_bytes, void* stream);
/*      SGPSSM (natural parameters): compute_cavity_x aep_models.py:1376-1387 */
int gpb_ssm_cavity(const double* xf1, const double* xf2, long T, int Q, double prior1, double prior2,
                   double alpha, double* cav_m, double* cav_v, void* stream);
/*      compute_transition_tilted aep_models.py:1317-1348 (2-D branch) on `total` = rows*Q elements:
 *      mt, vt = cavity of state t+1, mp, vp = propagated moments.  -> dm_layer = -dmt, dvt (scaled by s_dyn),
 *      sums[2] = {sum log Z terms, sum dvt} */
int gpb_ssm_transition(const double* mt, const double* vt, const double* mp, const double* vp, const double* sn,
                       long total, double alpha, double s_dyn, double* dm_layer, double* dvt, double* sums,
                       void* ws, size_t ws_bytes, void* stream);
/*      compute_logZ_grad_x aep_models.py:1234-1285: the three gradient sources of every latent state chained
 *      to its cavity naturals -> l1, l2[T,Q].  prev: (-dmt, dvt) of the transitions INTO rows first..; next /
 *      up: the layers' input gradients (row stride ld >= Q).  NULL dm = no such source. */
int gpb_ssm_sources(const double* xf1, const double* xf2, long T, int Q, double prior1, double prior2, double alpha,
                    const double* prev_dm, const double* prev_dv, long prev_first, long prev_count, int prev_ld,
                    const double* next_dm, const double* next_dv, long next_first, long next_count, int next_ld,
                    const double* up_dm, const double* up_dv, long up_first, long up_count, int up_ld,
                    double* l1, double* l2, void* stream);
/*      compute_posterior_grad_x / compute_cavity_grad_x / compute_phi_{posterior,cavity}_x
 *      aep_models.py:1208-1232,1287-1315,1389-1437 over ALL T rows -> gx1, gx2[T,Q], sums[2] = {phi_post, phi_cav} */
int gpb_ssm_xfinal(const double* xf1, const double* xf2, long T, int Q, double prior1, double prior2, double alpha,
                   const double* l1, const double* l2, double* gx1, double* gx2, double* sums, void* ws,
                   size_t ws_bytes, void* stream);

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
