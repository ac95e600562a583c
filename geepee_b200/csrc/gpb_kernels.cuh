// gpb_kernels.cuh -- device kernels of the geepee_b200 hot path (sm_100a).
//
// Every kernel is a template on the arithmetic type T (double = "fp64 mode",
// float = "fp32-psi mode"); all cross-row accumulators and all interfaces are
// fp64.  The source is also compiled for the host by tests/emu (GPB_CPU_EMU),
// see gpb_rt.cuh.  Reference formulas are cited per kernel (paths relative to
// /root/reference/geepee/).
//
// Notation: n rows, M pseudo-points (MP = M padded to 128/256/512 for the GEMM
// kernels), D/Q input dims, Do output dims, P = M(M+1)/2 unordered pairs.
#pragma once
#include "gpb_rt.cuh"

namespace gpb {

constexpr int kThreads = 256;
constexpr double kTwoPi = 6.283185307179586476925286766559;

template <typename T> struct V16;
template <> struct V16<double> { typedef double2 type; static constexpr int N = 2; };
template <> struct V16<float> { typedef float4 type; static constexpr int N = 4; };

template <typename T>
union VecU {
    typename V16<T>::type v;
    T e[V16<T>::N];
};

// -------------------------------------------------------------------------
// deterministic two-stage reduction helper: out[i] (+)= sum_g part[g*len + i]
// -------------------------------------------------------------------------
GPB_KERNEL void reduce_partials_kernel(const double* __restrict__ part, int G, long gstride,
                                       long len, double* __restrict__ out, int accumulate) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < len;
         i += (long)gridDim.x * blockDim.x) {
        double s = 0;
        for (int g = 0; g < G; g++) s += part[(long)g * gstride + i];
        out[i] = accumulate ? out[i] + s : s;
    }
}
// same result layout, many partials per output (G >= 64): one WARP per output element, the lanes
// share the G partials (fixed assignment and a fixed shuffle tree: deterministic)
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) reduce_partials_warp_kernel(const double* __restrict__ part, int G,
                                                                  long gstride, long len,
                                                                  double* __restrict__ out, int accumulate) {
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long i = wid; i < len; i += nw) {
        double s = 0;
        for (int g = lane; g < G; g += 32) s += part[(long)g * gstride + i];
        s = warp_sum(s);
        if (lane == 0) out[i] = accumulate ? out[i] + s : s;
    }
}

// sum of one double per thread over the block -> returned to thread 0 (others get junk)
GPB_DEVICE double block_sum(double v, double* scratch /* >= 8 doubles of smem */) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    sync_threads();
    if (lane == 0) scratch[w] = v;
    sync_threads();
    double s = 0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += scratch[i];
    return s;
}

// -------------------------------------------------------------------------
// a1. ARD-SE kernel matrix, kernels.py:10-22 (+ JITTER on the diagonal for Kuu,
//     base_models.py:461-463).  fp64, one thread per entry.
// -------------------------------------------------------------------------
GPB_KERNEL void kmat_kernel(const double* __restrict__ x, const double* __restrict__ z,
                            const double* __restrict__ ls, const double* __restrict__ sf,
                            int n, int M, int D, double jitter, double* __restrict__ out) {
    const double sf2 = exp(2.0 * sf[0]);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)n * M;
         i += (long)gridDim.x * blockDim.x) {
        int r = (int)(i / M), m = (int)(i % M);
        double r2 = 0;
        for (int q = 0; q < D; q++) {
            double d = x[(long)r * D + q] - z[(long)m * D + q];
            r2 += d * d * exp(-2.0 * ls[q]);
        }
        double k = sf2 * exp(-0.5 * r2);
        if (jitter != 0.0 && r == m) k += jitter;
        out[i] = k;
    }
}

// -------------------------------------------------------------------------
// a2. psi1[n,M], psi2[n,M,M] materialised -- the drop-in twin of the reference's
//     one native routine, kernels.py:181-240 (compute_psi_weave).  Same log-domain
//     expression as the weave body (lines 214-227).  Only the layer-level API and
//     the parity tests use it; the training path never writes psi2 to HBM.
// -------------------------------------------------------------------------
GPB_KERNEL void psi_stats_kernel(const double* __restrict__ mx, const double* __restrict__ vx,
                                 const double* __restrict__ z, const double* __restrict__ ls,
                                 const double* __restrict__ sf, int n, int M, int Q,
                                 double* __restrict__ psi1, double* __restrict__ psi2) {
    const double sf2 = exp(2.0 * sf[0]);
    const long total = (long)n * M * M;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long)gridDim.x * blockDim.x) {
        int b = (int)(i % M);
        int a = (int)((i / M) % M);
        long r = i / ((long)M * M);
        double lp2 = 0, lp1 = 0;
        for (int q = 0; q < Q; q++) {
            double lq = exp(2.0 * ls[q]);
            double vq = vx[r * Q + q], mq = mx[r * Q + q];
            double z1 = z[(long)a * Q + q], z2 = z[(long)b * Q + q];
            double muzhat = mq - (z1 + z2) / 2.0;
            double dz = z1 - z2;
            lp2 += -dz * dz / (4.0 * lq) - muzhat * muzhat / (2.0 * vq + lq) +
                   0.5 * log(lq / (lq + 2.0 * vq));
            if (b == 0) {
                double muz = mq - z1;
                lp1 += -muz * muz / 2.0 / (vq + lq) + 0.5 * log(lq / (lq + vq));
            }
        }
        psi2[i] = sf2 * sf2 * exp(lp2);
        if (b == 0) psi1[r * M + a] = sf2 * exp(lp1);
    }
}

// -------------------------------------------------------------------------
// a7 / a7'. Gaussian likelihood: lik_layers.py:104-133 (mode 0: AEP log Z tilted)
//     and 183-199,217-226 (mode 1: VFE expected log-lik).  Elementwise over
//     [n,Do]; writes dm, dv ALREADY multiplied by `scale` (what the layers'
//     backward consumes) and per-block partials {sum logZ-terms, sum for dsn}.
//     mode 0: part1 = sum of UNscaled dv (lik_layers.py:171-181)
//     mode 1: part1 = sum(-1 + (y-m)^2/sn2 + v/sn2)
// -------------------------------------------------------------------------
GPB_KERNEL void gauss_lik_kernel(const double* __restrict__ m, const double* __restrict__ v,
                                 const double* __restrict__ y, const double* __restrict__ sn,
                                 double alpha, double scale, long total, int mode,
                                 double* __restrict__ dm, double* __restrict__ dv,
                                 double* __restrict__ part /* [gridDim.x][2] */) {
    GPB_SHARED double scratch[16];
    const double sn2 = exp(2.0 * sn[0]);
    double s0 = 0, s1 = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long)gridDim.x * blockDim.x) {
        double mi = m[i], yi = y[i];
        if (mode == 0) {
            double vi = v[i] + sn2 / alpha;
            double e = yi - mi;
            s0 += -0.5 * (log(kTwoPi * vi) + e * e / vi) +
                  (0.5 * log(kTwoPi * sn2 / alpha) - 0.5 * alpha * log(kTwoPi * sn2));
            double dmi = e / vi;
            double dvi = -0.5 / vi + 0.5 * e * e / (vi * vi);
            s1 += dvi;
            dm[i] = scale * dmi;
            dv[i] = scale * dvi;
        } else {
            double vi = v[i];
            double t = yi * yi - 2 * yi * mi + mi * mi + vi;
            s0 += -0.5 * log(kTwoPi * sn2) - 0.5 / sn2 * t;
            s1 += -1.0 + t / sn2;
            dm[i] = scale * (yi - mi) / sn2;
            dv[i] = scale * (-0.5 / sn2);
        }
    }
    double r0 = block_sum(s0, scratch);
    double r1 = block_sum(s1, scratch + 8);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 2 + 0] = r0;
        part[blockIdx.x * 2 + 1] = r1;
    }
}

// -------------------------------------------------------------------------
// (8f rank 1) Probit likelihood for binary y in {-1,+1}: lik_layers.py:303-362 (mode 0: AEP
//     log Z tilted -- closed form for alpha == 1, Gauss-Hermite quadrature of degree ngh
//     otherwise, with the reference's eps guards) and 418-436 (mode 1: VFE expected log-lik).
//     Elementwise over [n,Do]; writes scale*dm, scale*dv and per-block sums of the log terms.
// -------------------------------------------------------------------------
GPB_KERNEL void probit_lik_kernel(const double* __restrict__ m, const double* __restrict__ v,
                                  const double* __restrict__ y, const double* __restrict__ gh_x,
                                  const double* __restrict__ gh_w, int ngh, double alpha, double scale,
                                  long total, int mode, double* __restrict__ dm, double* __restrict__ dv,
                                  double* __restrict__ part /* [gridDim.x][2] */) {
    GPB_SHARED double scratch[16];
    GPB_SHARED double sx[64], sw[64];
    for (int i = threadIdx.x; i < ngh && i < 64; i += blockDim.x) { sx[i] = gh_x[i]; sw[i] = gh_w[i]; }
    sync_threads();
    const double kSqrt2 = 1.4142135623730951, kPi = 3.14159265358979323846;
    double s0 = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long)gridDim.x * blockDim.x) {
        const double mi = m[i], vi = v[i], yi = y[i];
        double dmi, dvi;
        if (mode == 0 && alpha == 1.0) {
            const double t = yi * mi / sqrt(1.0 + vi);
            const double Z = 0.5 * (1.0 + erf(t / kSqrt2));
            const double eps = 1e-16;
            s0 += log(Z + eps);
            const double dt = 1.0 / (Z + eps) / sqrt(2.0 * kPi) * exp(-t * t / 2.0);
            dmi = dt * yi / sqrt(1.0 + vi);
            dvi = dt * (-0.5 * yi * mi / ((1.0 + vi) * sqrt(1.0 + vi)));
        } else if (mode == 0) {
            const double eps = 1e-8, sd = sqrt(2.0 * vi);
            double Zt = 0, sa = 0, sax = 0;
            for (int k = 0; k < ngh; k++) {
                const double ts = sx[k] * sd + mi;
                const double pdf = 0.5 * (1.0 + erf(yi * ts / kSqrt2)) + eps;
                Zt += pow(pdf, alpha) * sw[k];
                const double a = pow(pdf, alpha - 1.0) * exp(-ts * ts / 2.0);
                sa += sw[k] * a;
                sax += sw[k] * (a * sx[k]);
            }
            Zt /= sqrt(kPi);
            s0 += log(Zt);
            const double dZdm = sa * yi * alpha / kPi / kSqrt2;
            const double dZdv = sax * yi * alpha / kPi / kSqrt2 / sd;
            dmi = dZdm / Zt + eps;
            dvi = dZdv / Zt + eps;
        } else {
            const double sd = sqrt(2.0 * vi), isp = 1.0 / sqrt(kPi);
            double a = 0, b = 0;
            for (int k = 0; k < ngh; k++) {
                const double u = (sx[k] * sd + mi) * yi;
                const double cdf = 0.5 * erfc(-u / kSqrt2);
                const double w = sw[k] * isp;
                s0 += w * log(cdf);
                const double g = yi * w * exp(-u * u / 2.0) / sqrt(2.0 * kPi) / cdf;
                a += g;
                b += g * 0.5 * sx[k] * sqrt(2.0 / vi);
            }
            dmi = a;
            dvi = b;
        }
        dm[i] = scale * dmi;
        dv[i] = scale * dvi;
    }
    const double r0 = block_sum(s0, scratch);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 2 + 0] = r0;
        part[blockIdx.x * 2 + 1] = 0.0;
    }
}

// -------------------------------------------------------------------------
// a14. Linear-Gaussian emission, tilted (AEP): lik_layers.py:573-627
//     y ~ N(C x, diag(R)):  per row  Vy = diag(R/alpha) + C diag(vx) C^T  (Do x Do, SPD),
//     w = Vy^-1 (y - C mx),  log|Vy|,  Vy^-1;  dSig = -Vy^-1/2 + w w^T/2.
//     Thread per row; Cholesky, triangular inverse and Vy^-1 fully unrolled in registers
//     (DO, QT compile-time, zero / identity padded); writes scale*dmx, scale*dvx and per-block
//     partials  [ sum quad | sum log|Vy| | dRacc[DO] | dC[DO][QT] ]  with
//       dRacc[a] = sum_n (-Vy^-1[a,a]/2 + w_a^2/2),
//       dC[a,q]  = sum_n ( w_a mx_q + 2 vx_q sum_b dSig[a,b] C[b,q] ).
//     Replaces a batched cuSOLVER potrf/potrs/potri over n tiny matrices (whose per-call
//     cudaMalloc/cudaFree cost 220 ms per step at T = 1e6 in the first version of this round).
// -------------------------------------------------------------------------
template <int DO, int QT>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(128) gauss_emis_kernel(
    const double* __restrict__ mx, const double* __restrict__ vx, const double* __restrict__ y,
    const double* __restrict__ C, const double* __restrict__ R, double alpha, double scale, int n,
    int Q, int Do, double* __restrict__ dmx, double* __restrict__ dvx, double* __restrict__ part) {
    constexpr int NV = 2 + DO + DO * QT;
    GPB_SHARED double sC[DO * QT], sR[DO];
    GPB_SHARED double s_red[4 * NV];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < DO * QT; i += blockDim.x) {
        const int a = i / QT, q = i - a * QT;
        sC[i] = (a < Do && q < Q) ? C[a * Q + q] : 0.0;
    }
    if (tid < DO) sR[tid] = tid < Do ? R[tid] / alpha : 1.0;    // padded outputs: Vy = 1, y = 0
    sync_threads();
    double acc[NV];
    GPB_UNROLL
    for (int i = 0; i < NV; i++) acc[i] = 0;
    for (long row = (long)blockIdx.x * blockDim.x + tid; row < n; row += (long)gridDim.x * blockDim.x) {
        double m[QT], v[QT], L[DO][DO], yd[DO];
        GPB_UNROLL
        for (int q = 0; q < QT; q++) {
            m[q] = q < Q ? mx[row * Q + q] : 0.0;
            v[q] = q < Q ? vx[row * Q + q] : 0.0;
        }
        GPB_UNROLL
        for (int a = 0; a < DO; a++) {
            double s = a < Do ? y[row * Do + a] : 0.0;
            GPB_UNROLL
            for (int q = 0; q < QT; q++) s -= sC[a * QT + q] * m[q];
            yd[a] = s;
            GPB_UNROLL
            for (int b = 0; b <= a; b++) {
                double t = a == b ? sR[a] : 0.0;
                GPB_UNROLL
                for (int q = 0; q < QT; q++) t += sC[a * QT + q] * v[q] * sC[b * QT + q];
                L[a][b] = t;
            }
        }
        // Cholesky (lower, in place), log-determinant
        double ld = 0;
        GPB_UNROLL
        for (int j = 0; j < DO; j++) {
            double s = L[j][j];
            GPB_UNROLL
            for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
            const double d = sqrt(s), inv = 1.0 / d;
            L[j][j] = inv;                      // keep 1/L_jj on the diagonal
            ld += log(s);                       // = 2 log L_jj
            GPB_UNROLL
            for (int i = j + 1; i < DO; i++) {
                double t = L[i][j];
                GPB_UNROLL
                for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
                L[i][j] = t * inv;
            }
        }
        // Linv (lower) in place: column by column
        double Li[DO][DO];
        GPB_UNROLL
        for (int j = 0; j < DO; j++) {
            Li[j][j] = L[j][j];
            GPB_UNROLL
            for (int i = j + 1; i < DO; i++) {
                double t = 0;
                GPB_UNROLL
                for (int k = j; k < i; k++) t -= L[i][k] * Li[k][j];
                Li[i][j] = t * L[i][i];
            }
        }
        // Vinv = Linv^T Linv (symmetric), w = Vinv yd
        double Vi[DO][DO], wv[DO];
        GPB_UNROLL
        for (int a = 0; a < DO; a++)
            GPB_UNROLL
            for (int b = 0; b <= a; b++) {
                double t = 0;
                GPB_UNROLL
                for (int k = a; k < DO; k++) t += Li[k][a] * Li[k][b];
                Vi[a][b] = t;
                Vi[b][a] = t;
            }
        double quad = 0;
        GPB_UNROLL
        for (int a = 0; a < DO; a++) {
            double t = 0;
            GPB_UNROLL
            for (int b = 0; b < DO; b++) t += Vi[a][b] * yd[b];
            wv[a] = t;
            quad += t * yd[a];
        }
        acc[0] += -0.5 * quad;
        acc[1] += ld;
        // dSig C  and the per-row input gradients
        double dm_[QT], dv_[QT];
        GPB_UNROLL
        for (int q = 0; q < QT; q++) { dm_[q] = 0; dv_[q] = 0; }
        GPB_UNROLL
        for (int a = 0; a < DO; a++) {
            acc[2 + a] += -0.5 * Vi[a][a] + 0.5 * wv[a] * wv[a];
            GPB_UNROLL
            for (int q = 0; q < QT; q++) {
                double sc = 0;                  // (dSig C)[a][q]
                GPB_UNROLL
                for (int b = 0; b < DO; b++) sc += (-0.5 * Vi[a][b] + 0.5 * wv[a] * wv[b]) * sC[b * QT + q];
                acc[2 + DO + a * QT + q] += wv[a] * m[q] + 2.0 * v[q] * sc;
                dm_[q] += wv[a] * sC[a * QT + q];
                dv_[q] += sc * sC[a * QT + q];
            }
        }
        GPB_UNROLL
        for (int q = 0; q < QT; q++)
            if (q < Q) {
                dmx[row * Q + q] = scale * dm_[q];
                dvx[row * Q + q] = scale * dv_[q];
            }
    }
    // block reduction of the NV sums (4 warps)
    GPB_UNROLL
    for (int i = 0; i < NV; i++) {
        const double r = warp_sum(acc[i]);
        if (lane == 0) s_red[warp * NV + i] = r;
    }
    sync_threads();
    for (int i = tid; i < NV; i += blockDim.x) {
        double r = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += s_red[w * NV + i];
        part[(long)blockIdx.x * NV + i] = r;
    }
}

// padded record [2 | DO | DO*QT] -> compact [2 | Do | Do*Q]
// lik_layers.py:600-627: the sums of gauss_emis -> scale*logZ, scale*dR (wrt log-sqrt R), scale*dC
//   raw = [quad | log|Vy| sum | dRacc[Do] | dC[Do*Q]], R = variances;  fin = [scale logZ | 0 | dR[Do] | dC[Do*Q]]
GPB_KERNEL void gauss_emis_finish_kernel(const double* __restrict__ raw, const double* __restrict__ R, double alpha,
                                         double scale, double Nb, int Do, int Q, double* __restrict__ fin) {
    if (threadIdx.x == 0) {
        double slr = 0.0, slra = 0.0;
        for (int d = 0; d < Do; d++) {
            slr += log(R[d]);
            slra += log(R[d] / alpha);
        }
        const double vlog = -0.5 * (raw[1] - Nb * slra);
        fin[0] = scale * (-Nb * Do * 0.5 * alpha * log(2.0 * 3.14159265358979323846) - 0.5 * Nb * alpha * slr
                          + vlog + raw[0]);
        fin[1] = 0.0;
    }
    for (int i = threadIdx.x; i < Do; i += blockDim.x)
        fin[2 + i] = scale * ((raw[2 + i] / alpha + 0.5 * Nb * (1.0 - alpha) / R[i]) * 2.0 * R[i]);
    for (int i = threadIdx.x; i < Do * Q; i += blockDim.x) fin[2 + Do + i] = scale * raw[2 + Do + i];
}
GPB_KERNEL void gauss_emis_compact_kernel(const double* __restrict__ full, int DO, int QT, int Do, int Q,
                                          double* __restrict__ out) {
    for (int i = threadIdx.x; i < 2 + Do + Do * Q; i += blockDim.x) {
        double v;
        if (i < 2) v = full[i];
        else if (i < 2 + Do) v = full[2 + (i - 2)];
        else {
            const int k = i - 2 - Do, a = k / Q, q = k - a * Q;
            v = full[2 + DO + a * QT + q];
        }
        out[i] = v;
    }
}

// =========================================================================
// Deterministic-input layer (a5, a8)
// =========================================================================
// Tile geometry of the fused Kfu-generation + Kfu.B GEMM kernel.  256 threads =
// 8 warps arranged CW column-warps x RW row-warps; each warp owns 16 rows x 128
// columns of the [TN x MP] output tile T = Kfu_tile . B_d, each thread 16 rows x 4
// columns (64 accumulators).  The A operand (Kfu tile, generated on the fly from
// x and zu) stays resident in shared memory; B_d is streamed through a
// double-buffered cp.async ring in chunks of KB rows.
template <typename T, int MP>
struct DetCfg {
    static constexpr int CW = MP / 128;
    static constexpr int RW = 8 / CW;
    static constexpr int TN = 16 * RW;
    static constexpr int KB = 32768 / (MP * (int)sizeof(T));
    static constexpr int VEC = V16<T>::N;
    static constexpr int NJ = 4 / VEC;  // 16-byte column groups per thread
    static constexpr int DPMAX = 32;
    static constexpr size_t kt_bytes = (size_t)TN * MP * sizeof(T);
    static constexpr size_t bs_bytes = 2 * (size_t)KB * MP * sizeof(T);
    static constexpr size_t red_bytes = (size_t)CW * TN * 2 * sizeof(double);
    static constexpr size_t smem_bytes = kt_bytes + bs_bytes + red_bytes + DPMAX * sizeof(T);
};

template <typename T>
struct DetFwdArgs {
    const double* x;   // [n, D]
    const double* z;   // [M, D]
    const double* ls;  // [D]   log lengthscales
    const double* sf;  // [1]   log signal std
    const T* Ap;       // [Do, MP]      zero padded (Ahat or A)
    const T* Bp;       // [Do, MP, MP]  zero padded (Bhat_det or B_det)
    int n, M, D, Do;
    double* mout;      // [n, Do]
    double* vout;      // [n, Do]
    T* Ksave;          // [n, MP] or null
    T* Tsave;          // [n, Do, MP] or null
};

// Kfu tile generation, kernels.py:10-22: thread per column, rows looped.
template <typename T, int DP>
GPB_DEVICE void gen_k_tile(T* Kt, int MP, const T* xs, const double* __restrict__ z,
                           const T* ils2, T sf2, int M, int D, int TN, int rows_valid,
                           T* Ksave_tile) {
    for (int m = threadIdx.x; m < MP; m += blockDim.x) {
        T zr[DP];
        GPB_UNROLL
        for (int q = 0; q < DP; q++) zr[q] = (m < M && q < D) ? (T)z[(long)m * D + q] : (T)0;
        for (int r = 0; r < TN; r++) {
            T r2 = 0;
            GPB_UNROLL
            for (int q = 0; q < DP; q++) {
                T d = xs[r * DP + q] - zr[q];
                r2 += d * d * ils2[q];
            }
            T k = (m < M && r < rows_valid) ? sf2 * fast_exp((T)(-0.5) * r2) : (T)0;
            Kt[r * MP + m] = k;
            if (Ksave_tile != nullptr && r < rows_valid) Ksave_tile[(long)r * MP + m] = k;
        }
    }
}

// a5: aep_models.py:142-158 / base_models.py:265-284.
//   mout[n,d] = sum_m kfu[n,m] A[d,m];  vout[n,d] = sf2 + sum_ab B[d,a,b] kfu[n,a] kfu[n,b]
// also emits T[n,d,:] = B_d kfu[n,:] and kfu itself for the backward kernels.
template <typename T, int MP>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) det_fwd_kernel(DetFwdArgs<T> a) {
    typedef DetCfg<T, MP> C;
    typedef typename V16<T>::type VT;
    constexpr int TN = C::TN, KB = C::KB, VEC = C::VEC, NJ = C::NJ, CW = C::CW;
    GPB_DYN_SMEM(smem);
    T* Kt = (T*)smem;
    T* Bs = (T*)(smem + C::kt_bytes);
    T* xs = Bs;  // x tile aliases the B ring (only live during Kfu generation)
    double* red = (double*)(smem + C::kt_bytes + C::bs_bytes);
    T* ils2 = (T*)(smem + C::kt_bytes + C::bs_bytes + C::red_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cw = warp % CW, rw = warp / CW;
    const int D = a.D, M = a.M, Do = a.Do, n = a.n;
    const int DP = D <= 4 ? 4 : (D <= 8 ? 8 : (D <= 16 ? 16 : 32));
    const T sf2 = (T)exp(2.0 * a.sf[0]);
    if (tid < C::DPMAX) ils2[tid] = tid < D ? (T)exp(-2.0 * a.ls[tid]) : (T)0;

    const int ntiles = (n + TN - 1) / TN;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * TN;
        const int rows_valid = (n - row0) < TN ? (n - row0) : TN;
        sync_threads();  // previous tile fully consumed (Kt, Bs/xs, red)
        for (int i = tid; i < TN * DP; i += kThreads) {
            int r = i / DP, q = i - r * DP;
            xs[i] = (r < rows_valid && q < D) ? (T)a.x[(long)(row0 + r) * D + q] : (T)0;
        }
        sync_threads();
        T* ks = a.Ksave ? a.Ksave + (long)row0 * MP : nullptr;
        if (DP == 4) gen_k_tile<T, 4>(Kt, MP, xs, a.z, ils2, sf2, M, D, TN, rows_valid, ks);
        else if (DP == 8) gen_k_tile<T, 8>(Kt, MP, xs, a.z, ils2, sf2, M, D, TN, rows_valid, ks);
        else if (DP == 16) gen_k_tile<T, 16>(Kt, MP, xs, a.z, ils2, sf2, M, D, TN, rows_valid, ks);
        else gen_k_tile<T, 32>(Kt, MP, xs, a.z, ils2, sf2, M, D, TN, rows_valid, ks);
        sync_threads();

        for (int d = 0; d < Do; d++) {
            const T* Bd = a.Bp + (long)d * MP * MP;
            T acc[16][4];
            GPB_UNROLL
            for (int r = 0; r < 16; r++)
                GPB_UNROLL
                for (int c = 0; c < 4; c++) acc[r][c] = 0;

            constexpr int nchunks = MP / KB;
            constexpr int chunk_vecs = KB * MP / VEC;  // 16B vectors per chunk
            // prologue: chunk 0 -> buffer 0
            for (int i = tid; i < chunk_vecs; i += kThreads)
                cp_async16(Bs + (long)i * VEC, Bd + (long)i * VEC);
            cp_async_commit();
            for (int c = 0; c < nchunks; c++) {
                if (c + 1 < nchunks) {
                    T* dst = Bs + (long)((c + 1) & 1) * KB * MP;
                    const T* src = Bd + (long)(c + 1) * KB * MP;
                    for (int i = tid; i < chunk_vecs; i += kThreads)
                        cp_async16(dst + (long)i * VEC, src + (long)i * VEC);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                sync_threads();
                const T* Bc = Bs + (long)(c & 1) * KB * MP;
                const T* Ka = Kt + (long)(rw * 16) * MP + c * KB;
                GPB_UNROLL_N(2)
                for (int k0 = 0; k0 < KB; k0 += VEC) {
                    // B operand: VEC k-rows x 4 columns of this thread
                    T b[VEC][4];
                    GPB_UNROLL
                    for (int kk = 0; kk < VEC; kk++)
                        GPB_UNROLL
                        for (int j = 0; j < NJ; j++) {
                            VecU<T> u;
                            u.v = *(const VT*)(Bc + (long)(k0 + kk) * MP + cw * 128 +
                                               j * (32 * VEC) + lane * VEC);
                            GPB_UNROLL
                            for (int e = 0; e < VEC; e++) b[kk][j * VEC + e] = u.e[e];
                        }
                    // 8 rows at a time, k outermost inside the group: an accumulator is touched again
                    // only after 32 other FMAs (the fp64 pipe stalls on back-to-back dependent FMAs)
                    GPB_UNROLL
                    for (int rg = 0; rg < 16; rg += 8) {
                        VecU<T> av[8];  // VEC consecutive k of each row (warp-uniform addresses)
                        GPB_UNROLL
                        for (int r = 0; r < 8; r++) av[r].v = *(const VT*)(Ka + (long)(rg + r) * MP + k0);
                        GPB_UNROLL
                        for (int kk = 0; kk < VEC; kk++)
                            GPB_UNROLL
                            for (int r = 0; r < 8; r++)
                                GPB_UNROLL
                                for (int cc = 0; cc < 4; cc++) acc[rg + r][cc] += av[r].e[kk] * b[kk][cc];
                    }
                }
                sync_threads();
            }

            // epilogue: row reductions  sum_m K[r,m] T[r,m]  and  sum_m K[r,m] A[d,m]
            GPB_UNROLL
            for (int r = 0; r < 16; r++) {
                const int row = rw * 16 + r;
                double pv = 0, pm = 0;
                GPB_UNROLL
                for (int j = 0; j < NJ; j++) {
                    const int col = cw * 128 + j * (32 * VEC) + lane * VEC;
                    VecU<T> kv, av;
                    kv.v = *(const VT*)(Kt + (long)row * MP + col);
                    av.v = *(const VT*)(a.Ap + (long)d * MP + col);
                    GPB_UNROLL
                    for (int e = 0; e < VEC; e++) {
                        pv += (double)acc[r][j * VEC + e] * (double)kv.e[e];
                        pm += (double)kv.e[e] * (double)av.e[e];
                    }
                    if (a.Tsave != nullptr && row < rows_valid) {
                        VecU<T> tv;
                        GPB_UNROLL
                        for (int e = 0; e < VEC; e++) tv.e[e] = acc[r][j * VEC + e];
                        *(VT*)(a.Tsave + ((long)(row0 + row) * Do + d) * MP + col) = tv.v;
                    }
                }
                pv = warp_sum(pv);
                pm = warp_sum(pm);
                if (lane == 0) {
                    red[(cw * TN + row) * 2 + 0] = pv;
                    red[(cw * TN + row) * 2 + 1] = pm;
                }
            }
            sync_threads();
            if (tid < rows_valid) {
                double pv = 0, pm = 0;
                for (int w = 0; w < CW; w++) {
                    pv += red[(w * TN + tid) * 2 + 0];
                    pm += red[(w * TN + tid) * 2 + 1];
                }
                a.vout[(long)(row0 + tid) * Do + d] = (double)sf2 + pv;
                a.mout[(long)(row0 + tid) * Do + d] = pm;
            }
        }
    }
}

// -------------------------------------------------------------------------
// fp64 tensor-core variant of a5 (DMMA.8x8x4, the native FP64 MMA of sm_100a).
// Measured on this pool's B200 (tools/probe): DMMA sustains 37 TFLOP/s = the nominal FP64 peak with 8
// resident warps, whereas DFMA streams with three distinct register operands top out at 65-78 %
// of it (register-operand bandwidth) -- the SIMT kernel above reaches 16.4 TFLOP/s.
//   * tile: TN rows x MP columns per CTA pass; 8 warps as CW column-warps x RW row-warps, each
//     warp 32 rows x 64 columns = 4 x 8 C fragments (64 accumulators per lane);
//   * A operand: the Kfu tile, generated on chip into shared memory (row stride MP+4 doubles =
//     4 mod 16, so the 32 fragment loads of a warp hit 32 different banks); the exponent is
//     formed in expanded form  S(-|x'|^2/2 + 2 sf) + S(-|z'|^2/2) + sum_q (S x'_q) z'_q  with
//     x' = x/l, z' = z/l (D fma per element) and exponentiated with the table-based exp
//     (64 entries x 16 replicas, degree-5 polynomial, see ExpDom);
//   * B operand: B_d streamed through a double-buffered cp.async ring of KB rows, same padding;
//   * epilogue on the C fragments: mout, vout row sums (quad shuffle + cross-warp smem), T store.
template <int MP>
struct DetMmaCfg {
    static constexpr int CW = MP / 64;            // 2, 4, 8 column warps
    static constexpr int RW = 8 / CW;             // 4, 2, 1 row warps
    static constexpr int TN = 32 * RW;            // 128, 64, 32 rows per tile
    static constexpr int LD = MP + 4;             // padded row stride in doubles
    static constexpr int KB = 4096 / MP;          // B rows per ring stage: 32, 16, 8
    static constexpr int DPMAX = 32;
    static constexpr int ETAB = 64 * 16;          // replicated 2^(j/64) table
    static constexpr size_t kt_bytes = (size_t)TN * LD * 8;
    static constexpr size_t bs_bytes = 2 * (size_t)KB * LD * 8;     // also holds the x tile
    static constexpr size_t red_bytes = (size_t)CW * TN * 2 * 8;
    static constexpr size_t misc_bytes = (ETAB + TN + DPMAX) * 8;
    static constexpr size_t smem_bytes = kt_bytes + bs_bytes + red_bytes + misc_bytes;
};

// exp(xs / S) for xs = S x, S = 64/ln2: xs = 64 k + j + r -> 2^k 2^(j/64) e^(r ln2/64); 9 fp64 ops.
GPB_DEVICE double exp_dom64(double xs, const double* __restrict__ tab /* [64][16] */, int lane16) {
    constexpr double h = 0.693147180559945309417232 / 64.0;
    constexpr double c1 = h, c2 = h * h / 2, c3 = h * h * h / 6, c4 = h * h * h * h / 24,
                     c5 = h * h * h * h * h / 120;
    const double magic = 6755399441055744.0;
    double kd = xs + magic;
#ifndef GPB_CPU_EMU
    const int n = __double2loint(kd);
#else
    int64_t bits;
    memcpy(&bits, &kd, 8);
    const int n = (int)(int32_t)(bits & 0xffffffff);
#endif
    kd -= magic;
    const double r = xs - kd;
    const double t = tab[((n & 63) << 4) + lane16];
    double q = c5 * r + c4;
    q = q * r + c3;
    q = q * r + c2;
    q = q * r + c1;
    const double p = (t * r) * q + t;
    int k = n >> 6;
    k = k < -1021 ? -1021 : k;
#ifndef GPB_CPU_EMU
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
    return ldexp(p, k);
#endif
}

template <int MP, int DP>
GPB_DEVICE void gen_k_tile_mma(double* Kt, const double* xs, const double* an, const double* __restrict__ z,
                               const double* ils, const double* tab, int M, int D, int TN, int LD,
                               int rows_valid, double* Ksave_tile) {
    const int lane16 = threadIdx.x & 15;
    for (int m = threadIdx.x; m < MP; m += blockDim.x) {
        double zr[DP], bm = 0;
        GPB_UNROLL
        for (int q = 0; q < DP; q++) {
            zr[q] = (m < M && q < D) ? z[(long)m * D + q] * ils[q] : 0.0;
            bm -= zr[q] * zr[q];
        }
        bm *= 0.5 * (64.0 / 0.693147180559945309417232);
        GPB_UNROLL_N(4)
        for (int r = 0; r < TN; r++) {
            double e = an[r] + bm;
            GPB_UNROLL
            for (int q = 0; q < DP; q++) e += xs[r * DP + q] * zr[q];
            const double k = (m < M && r < rows_valid) ? exp_dom64(e, tab, lane16) : 0.0;
            Kt[r * LD + m] = k;
            if (Ksave_tile != nullptr && r < rows_valid) Ksave_tile[(long)r * MP + m] = k;
        }
    }
}

template <int MP>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) det_fwd_mma_kernel(DetFwdArgs<double> a) {
    typedef DetMmaCfg<MP> C;
    constexpr int TN = C::TN, KB = C::KB, LD = C::LD, CW = C::CW;
    constexpr double kS = 64.0 / 0.693147180559945309417232;
    GPB_DYN_SMEM(smem);
    double* Kt = (double*)smem;
    double* Bs = (double*)(smem + C::kt_bytes);
    double* xs = Bs;   // x tile aliases the B ring (only live during Kfu generation)
    double* red = (double*)(smem + C::kt_bytes + C::bs_bytes);
    double* tab = (double*)(smem + C::kt_bytes + C::bs_bytes + C::red_bytes);
    double* an = tab + C::ETAB;
    double* ils = an + TN;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int cw = warp % CW, rw = warp / CW;
    const int D = a.D, M = a.M, Do = a.Do, n = a.n;
    const int DP = D <= 4 ? 4 : (D <= 8 ? 8 : (D <= 16 ? 16 : 32));
    const double sf2 = exp(2.0 * a.sf[0]);
    if (tid < C::DPMAX) ils[tid] = tid < D ? exp(-a.ls[tid]) : 0.0;
    for (int i = tid; i < C::ETAB; i += kThreads) tab[i] = exp2((double)(i >> 4) * (1.0 / 64.0));

    const int ntiles = (n + TN - 1) / TN;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * TN;
        const int rows_valid = (n - row0) < TN ? (n - row0) : TN;
        sync_threads();  // previous tile fully consumed (Kt, Bs/xs, red); ils/tab visible
        for (int i = tid; i < TN * DP; i += kThreads) {
            int r = i / DP, q = i - r * DP;
            xs[i] = (r < rows_valid && q < D) ? a.x[(long)(row0 + r) * D + q] * ils[q] : 0.0;
        }
        sync_threads();
        if (tid < TN) {   // an = S (2 sf - |x'|^2 / 2); then scale the row by S
            double s = 0;
            for (int q = 0; q < DP; q++) s += xs[tid * DP + q] * xs[tid * DP + q];
            an[tid] = kS * (2.0 * a.sf[0] - 0.5 * s);
        }
        sync_threads();
        for (int i = tid; i < TN * DP; i += kThreads) xs[i] *= kS;
        sync_threads();
        double* ks = a.Ksave ? a.Ksave + (long)row0 * MP : nullptr;
        if (DP == 4) gen_k_tile_mma<MP, 4>(Kt, xs, an, a.z, ils, tab, M, D, TN, LD, rows_valid, ks);
        else if (DP == 8) gen_k_tile_mma<MP, 8>(Kt, xs, an, a.z, ils, tab, M, D, TN, LD, rows_valid, ks);
        else if (DP == 16) gen_k_tile_mma<MP, 16>(Kt, xs, an, a.z, ils, tab, M, D, TN, LD, rows_valid, ks);
        else gen_k_tile_mma<MP, 32>(Kt, xs, an, a.z, ils, tab, M, D, TN, LD, rows_valid, ks);
        sync_threads();

        for (int d = 0; d < Do; d++) {
            const double* Bd = a.Bp + (long)d * MP * MP;
            double acc[4][8][2];
            GPB_UNROLL
            for (int i = 0; i < 4; i++)
                GPB_UNROLL
                for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0;

            constexpr int nchunks = MP / KB;
            constexpr int row_vecs = MP / 2;              // 16-byte vectors per B row
            constexpr int chunk_vecs = KB * row_vecs;
            auto issue = [&](int c) {
                double* dst = Bs + (long)(c & 1) * KB * LD;
                const double* src = Bd + (long)c * KB * MP;
                for (int i = tid; i < chunk_vecs; i += kThreads) {
                    const int r = i / row_vecs, cv = i - r * row_vecs;
                    cp_async16(dst + (long)r * LD + cv * 2, src + (long)r * MP + cv * 2);
                }
                cp_async_commit();
            };
            issue(0);
            for (int c = 0; c < nchunks; c++) {
                if (c + 1 < nchunks) {
                    issue(c + 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                sync_threads();
                const double* Bc = Bs + (long)(c & 1) * KB * LD + cw * 64 + g;
                const double* Ka = Kt + (long)(rw * 32 + g) * LD + c * KB + t;
                GPB_UNROLL_N(2)
                for (int ks4 = 0; ks4 < KB; ks4 += 4) {
                    double af[4], bf[8];
                    GPB_UNROLL
                    for (int i = 0; i < 4; i++) af[i] = Ka[(long)(i * 8) * LD + ks4];
                    GPB_UNROLL
                    for (int j = 0; j < 8; j++) bf[j] = Bc[(long)(ks4 + t) * LD + j * 8];
                    GPB_UNROLL
                    for (int i = 0; i < 4; i++)
                        GPB_UNROLL
                        for (int j = 0; j < 8; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
                sync_threads();
            }

            // epilogue: this lane holds T[row][col], T[row][col+1] for row = rw*32 + i*8 + g,
            // col = cw*64 + j*8 + 2t
            GPB_UNROLL
            for (int i = 0; i < 4; i++) {
                const int row = rw * 32 + i * 8 + g;
                double pv = 0, pm = 0;
                GPB_UNROLL
                for (int j = 0; j < 8; j++) {
                    const int col = cw * 64 + j * 8 + 2 * t;
                    const double2 kv = *(const double2*)(Kt + (long)row * LD + col);
                    const double2 av = *(const double2*)(a.Ap + (long)d * MP + col);
                    pv += acc[i][j][0] * kv.x + acc[i][j][1] * kv.y;
                    pm += kv.x * av.x + kv.y * av.y;
                    if (a.Tsave != nullptr && row < rows_valid)
                        *(double2*)(a.Tsave + ((long)(row0 + row) * Do + d) * MP + col) =
                            make_double2(acc[i][j][0], acc[i][j][1]);
                }
                pv += shfl_xor(pv, 1);
                pm += shfl_xor(pm, 1);
                pv += shfl_xor(pv, 2);
                pm += shfl_xor(pm, 2);
                if (t == 0) {
                    red[(cw * TN + row) * 2 + 0] = pv;
                    red[(cw * TN + row) * 2 + 1] = pm;
                }
            }
            sync_threads();
            if (tid < rows_valid) {
                double pv = 0, pm = 0;
                for (int w = 0; w < CW; w++) {
                    pv += red[(w * TN + tid) * 2 + 0];
                    pm += red[(w * TN + tid) * 2 + 1];
                }
                a.vout[(long)(row0 + tid) * Do + d] = sf2 + pv;
                a.mout[(long)(row0 + tid) * Do + d] = pm;
            }
        }
    }
}

// a8 (row-streaming part), aep_models.py:452-460,490 + kernels.py:381-399 (kfucompDer):
//   L[n,m] = (sum_d dm[n,d] A[d,m] + 2 dv[n,d] T[n,d,m]) kfu[n,m]
//   dsf2 += sum L / sf2 ; dZ[m,q] -= L (z_mq - x_nq)/l_q^2 ; dl_q += L (z_mq-x_nq)^2/l_q^3
//   dA[d,m] += dm[n,d] kfu[n,m]
// Thread per pseudo-point column, rows streamed from the saved Kfu / T buffers, U rows in
// flight per thread (all loads issued before the first use) so the kernel runs at HBM speed.
// grid = (row chunks, MP/CWB); each (block, row-group) writes one partial record:
//   [ cs(MP) | dz(MP*D) | dl(MP*D) | dA(Do*MP) ]   (dl kept per column, summed later)
// DP = input dims per pass (exact for the common D, extra passes if D > 16);
// DOB = output dims unrolled at compile time (1, 2, 4) or 0 = runtime loop (any Do).
template <typename T, int DP, int DOB>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) det_bwd_kernel(
    const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ ls,
    const T* __restrict__ Ap, const double* __restrict__ dm, const double* __restrict__ dv,
    const T* __restrict__ Ksave, const T* __restrict__ Tsave, int n, int M, int MP, int D, int Do,
    int rows_per_block, double* __restrict__ part, long rec_len) {
    constexpr int TR = 64;
    constexpr int U = 4;
    constexpr int DOS = DOB > 0 ? DOB : 1;
    GPB_SHARED T xs[TR * DP];
    GPB_SHARED double dms[TR * 8], dvs[TR * 8];
    const int tid = threadIdx.x;
    const int CWB = MP < 256 ? MP : 256;
    const int RY = kThreads / CWB;
    const int cx = tid % CWB, ry = tid / CWB;
    const int c = blockIdx.y * CWB + cx;
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = (r_begin + rows_per_block) < n ? (r_begin + rows_per_block) : n;
    double* rec = part + ((long)(blockIdx.x * RY + ry)) * rec_len;

    // ---- pass family 1: L-dependent sums, DP input dims at a time -------------------
    for (int q0 = 0; q0 < D; q0 += DP) {
        T zr[DP];
        double dz[DP], dl[DP];
        double cs = 0;
        T ap[DOS];
        GPB_UNROLL
        for (int q = 0; q < DP; q++) {
            zr[q] = (c < M && q0 + q < D) ? (T)z[(long)c * D + q0 + q] : (T)0;
            dz[q] = 0;
            dl[q] = 0;
        }
        GPB_UNROLL
        for (int d = 0; d < DOS; d++) ap[d] = (DOB > 0 && d < Do) ? Ap[(long)d * MP + c] : (T)0;
        for (int t0 = r_begin; t0 < r_end; t0 += TR) {
            const int tv = (r_end - t0) < TR ? (r_end - t0) : TR;
            sync_threads();
            for (int i = tid; i < TR * DP; i += kThreads) {
                int r = i / DP, q = i - r * DP;
                xs[i] = (r < tv && q0 + q < D) ? (T)x[(long)(t0 + r) * D + q0 + q] : (T)0;
            }
            if (DOB > 0)
                for (int i = tid; i < TR * DOS; i += kThreads) {
                    int r = i / DOS, d = i - r * DOS;
                    bool ok = r < tv && d < Do;
                    dms[r * 8 + d] = ok ? dm[(long)(t0 + r) * Do + d] : 0.0;
                    dvs[r * 8 + d] = ok ? 2.0 * dv[(long)(t0 + r) * Do + d] : 0.0;
                }
            sync_threads();
            for (int rb = ry * U; rb < tv; rb += RY * U) {
                // issue every global load of U rows first
                T kk[U], tt[U][DOS];
                GPB_UNROLL
                for (int u = 0; u < U; u++) {
                    const bool ok = (rb + u) < tv;
                    const long row = t0 + (ok ? rb + u : rb);
                    kk[u] = ok ? Ksave[row * MP + c] : (T)0;
                    if (DOB > 0) {
                        GPB_UNROLL
                        for (int d = 0; d < DOS; d++)
                            tt[u][d] = (d < Do) ? Tsave[(row * Do + d) * MP + c] : (T)0;
                    }
                }
                GPB_UNROLL
                for (int u = 0; u < U; u++) {
                    const int r = (rb + u) < tv ? rb + u : rb;   // clamped rows carry k = 0
                    double g = 0;
                    if (DOB > 0) {
                        GPB_UNROLL
                        for (int d = 0; d < DOS; d++)
                            g += dms[r * 8 + d] * (double)ap[d] + dvs[r * 8 + d] * (double)tt[u][d];
                    } else {
                        const long row = t0 + r;
                        for (int d = 0; d < Do; d++)
                            g += dm[row * Do + d] * (double)Ap[(long)d * MP + c] +
                                 2.0 * dv[row * Do + d] * (double)Tsave[(row * Do + d) * MP + c];
                    }
                    const double L = g * (double)kk[u];
                    cs += L;
                    GPB_UNROLL
                    for (int q = 0; q < DP; q++) {
                        double diff = (double)zr[q] - (double)xs[r * DP + q];
                        double t = L * diff;
                        dz[q] += t;
                        dl[q] += t * diff;
                    }
                }
            }
        }
        if (q0 == 0) rec[c] = cs;
        GPB_UNROLL
        for (int q = 0; q < DP; q++)
            if (q0 + q < D) {
                double il2 = exp(-2.0 * ls[q0 + q]);
                rec[(long)MP + (long)c * D + q0 + q] = -dz[q] * il2;
                rec[(long)MP + (long)MP * D + (long)c * D + q0 + q] = dl[q] * il2 * exp(-ls[q0 + q]);
            }
    }
    // ---- pass family 2: dA[d, c] = sum_n dm[n,d] kfu[n,c], 8 output dims at a time ---
    for (int d0 = 0; d0 < Do; d0 += 8) {
        const int dn = (Do - d0) < 8 ? (Do - d0) : 8;
        double dA[8];
        GPB_UNROLL
        for (int i = 0; i < 8; i++) dA[i] = 0;
        for (int t0 = r_begin; t0 < r_end; t0 += TR) {
            const int tv = (r_end - t0) < TR ? (r_end - t0) : TR;
            sync_threads();
            for (int i = tid; i < TR * 8; i += kThreads) {
                int r = i / 8, d = i - r * 8;
                dms[i] = (r < tv && d < dn) ? dm[(long)(t0 + r) * Do + d0 + d] : 0.0;
            }
            sync_threads();
            for (int rb = ry * U; rb < tv; rb += RY * U) {
                T kk[U];
                GPB_UNROLL
                for (int u = 0; u < U; u++)
                    kk[u] = (rb + u) < tv ? Ksave[(long)(t0 + rb + u) * MP + c] : (T)0;
                GPB_UNROLL
                for (int u = 0; u < U; u++) {
                    const int r = (rb + u) < tv ? rb + u : rb;
                    GPB_UNROLL
                    for (int i = 0; i < 8; i++) dA[i] += dms[r * 8 + i] * (double)kk[u];
                }
            }
        }
        for (int i = 0; i < dn; i++) rec[(long)MP + 2L * MP * D + (long)(d0 + i) * MP + c] = dA[i];
    }
}

// Input gradient of the deterministic layer (Monte-Carlo propagation: the layer is fed samples of an
// uncertain input and the gradient travels back to the sample, aep_models.py:346-350 +
// kernels.py:393-395 kfucompDer(grad_x=True)):
//   L[n,m]  = (sum_d dm[n,d] A[d,m] + 2 dv[n,d] T[n,d,m]) kfu[n,m]
//   dx[n,q] = sum_m L[n,m] (z[m,q] - x[n,q]) / l_q^2
// One warp per row streams the saved Kfu / T rows (coalesced), z sits in shared memory; the D + 1
// row sums are reduced with shuffles.  HBM bound: (1 + Do) MP sizeof(T) bytes per row.
template <typename T>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) det_dx_kernel(
    const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ ls,
    const T* __restrict__ Ap, const double* __restrict__ dm, const double* __restrict__ dv,
    const T* __restrict__ Ksave, const T* __restrict__ Tsave, int n, int M, int MP, int D, int Do,
    double* __restrict__ dx) {
    GPB_DYN_SMEM(dsm);
    double* zs = (double*)dsm;              // [M, D]
    double* il2 = zs + (long)M * D;         // [D]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < M * D; i += kThreads) zs[i] = z[i];
    for (int i = tid; i < D; i += kThreads) il2[i] = exp(-2.0 * ls[i]);
    sync_threads();
    for (long row = (long)blockIdx.x * 8 + warp; row < n; row += (long)gridDim.x * 8) {
        for (int q0 = 0; q0 < D; q0 += 8) {
            double acc[8], accL = 0;
            GPB_UNROLL
            for (int q = 0; q < 8; q++) acc[q] = 0;
            for (int m = lane; m < M; m += 32) {
                double g = 0;
                for (int d = 0; d < Do; d++)
                    g += dm[row * Do + d] * (double)Ap[(long)d * MP + m] +
                         2.0 * dv[row * Do + d] * (double)Tsave[(row * Do + d) * MP + m];
                const double L = g * (double)Ksave[row * MP + m];
                accL += L;
                GPB_UNROLL
                for (int q = 0; q < 8; q++)
                    if (q0 + q < D) acc[q] += L * zs[(long)m * D + q0 + q];
            }
            accL = warp_sum(accL);
            GPB_UNROLL
            for (int q = 0; q < 8; q++) acc[q] = warp_sum(acc[q]);
            if (lane == 0) {
                GPB_UNROLL
                for (int q = 0; q < 8; q++)
                    if (q0 + q < D) dx[row * D + q0 + q] = (acc[q] - x[row * D + q0 + q] * accL) * il2[q0 + q];
            }
        }
    }
}

// Ring version of the row-streaming backward (Do <= 4, D <= 16): the saved Kfu / T rows and the
// per-row x, dm, dv records of TRS rows travel through a cp.async ring (2-3 stages, no register
// staging), so the global loads of the next chunks are in flight while a chunk is reduced --
// the register-batched kernel above exposes the full load latency once per 4 rows (ncu: 50 %
// long-scoreboard stalls, 27 % of HBM peak).  dA is accumulated in the same sweep (the kernel
// above re-reads Kfu for it).  Same partial-record layout.
template <typename T, int DOS>
struct DetBwdRing {
    static constexpr int TRS = 8;                                   // rows per stage
    static constexpr int STAGES = DOS == 4 ? 2 : 3;
    static constexpr int CW = 256;                                  // columns per block (<= MP)
    static constexpr size_t tile_bytes = (size_t)TRS * CW * sizeof(T) * (1 + DOS);
    static constexpr size_t rec_doubles = TRS * (16 + 2 * DOS);     // x (DP <= 16), dm, 2 dv
    static constexpr size_t stage_bytes = tile_bytes + rec_doubles * sizeof(double);
    static constexpr size_t smem_bytes = STAGES * stage_bytes;
};

template <typename T, int DP, int DOS, int CWB>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) det_bwd_ring_kernel(
    const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ ls,
    const T* __restrict__ Ap, const double* __restrict__ dm, const double* __restrict__ dv,
    const T* __restrict__ Ksave, const T* __restrict__ Tsave, int n, int M, int MP, int D, int Do,
    int rows_per_block, double* __restrict__ part, long rec_len) {
    typedef DetBwdRing<T, DOS> C;
    constexpr int TRS = C::TRS, STAGES = C::STAGES, VEC = V16<T>::N;
    GPB_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    constexpr int RY = kThreads / CWB;            // CWB = min(MP, 256) columns per block
    const int cx = tid % CWB, ry = tid / CWB;
    const int cbase = blockIdx.y * CWB;
    const int c = cbase + cx;
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = (r_begin + rows_per_block) < n ? (r_begin + rows_per_block) : n;
    const int nchunk = (r_end - r_begin + TRS - 1) / TRS;
    double* rec = part + ((long)(blockIdx.x * RY + ry)) * rec_len;

    auto stage_K = [&](int st) { return (T*)(smem + (size_t)st * C::stage_bytes); };
    auto stage_R = [&](int st) { return (double*)(smem + (size_t)st * C::stage_bytes + C::tile_bytes); };
    auto issue = [&](int chunk) {
        if (chunk < nchunk) {
            const int st = chunk % STAGES;
            T* Ks = stage_K(st);
            double* Rs = stage_R(st);          // [x: TRS*DP | dm: TRS*DOS | 2dv (raw dv here): TRS*DOS]
            const int t0 = r_begin + chunk * TRS;
            constexpr int vec_per_row = CWB / VEC;
            constexpr int nvec = TRS * vec_per_row * (1 + DOS);
            static_assert(nvec % kThreads == 0, "tile copies must divide evenly");
            GPB_UNROLL
            for (int it = 0; it < nvec / kThreads; it++) {
                const int v = tid + it * kThreads;
                const int which = v / (TRS * vec_per_row);          // 0: K, 1..DOS: T_d
                const int rem = v - which * (TRS * vec_per_row);
                const int r = rem / vec_per_row, cv = rem - r * vec_per_row;
                const long row = (long)t0 + r;
                const bool ok = row < r_end && (which == 0 || which - 1 < Do);
                const long rowc = row < r_end ? row : (long)r_begin;
                const T* src = which == 0 ? Ksave + rowc * MP + cbase + cv * VEC
                                          : Tsave + (rowc * Do + (which - 1 < Do ? which - 1 : 0)) * MP + cbase + cv * VEC;
                cp_async16_zfill(Ks + (long)which * TRS * CWB + (long)r * CWB + cv * VEC, src, ok);
            }
            for (int i = tid; i < TRS * DP; i += kThreads) {
                const int r = i / DP, q = i - r * DP;
                const long row = (long)t0 + r;
                const bool ok = row < r_end && q < D;
                cp_async8_zfill(Rs + i, x + (ok ? row * D + q : 0), ok);
            }
            for (int i = tid; i < 2 * TRS * DOS; i += kThreads) {
                const int w = i / (TRS * DOS), rem = i - w * (TRS * DOS);
                const int r = rem / DOS, d = rem - r * DOS;
                const long row = (long)t0 + r;
                const bool ok = row < r_end && d < Do;
                const double* src = (w == 0 ? dm : dv) + (ok ? row * Do + d : 0);
                cp_async8_zfill(Rs + TRS * DP + i, src, ok);
            }
        }
        cp_async_commit();
    };

    T zr[DP];
    double dz[DP], dl[DP], dA[DOS], cs = 0;
    T ap[DOS];
    GPB_UNROLL
    for (int q = 0; q < DP; q++) {
        zr[q] = (c < M && q < D) ? (T)z[(long)c * D + q] : (T)0;
        dz[q] = 0;
        dl[q] = 0;
    }
    GPB_UNROLL
    for (int d = 0; d < DOS; d++) {
        ap[d] = d < Do ? Ap[(long)d * MP + c] : (T)0;
        dA[d] = 0;
    }
    GPB_UNROLL
    for (int s0 = 0; s0 < STAGES - 1; s0++) issue(s0);
    for (int ch = 0; ch < nchunk; ch++) {
        if (STAGES == 3) cp_async_wait<1>(); else cp_async_wait<0>();
        sync_threads();                       // chunk ch landed everywhere; stage (ch-1) is free
        issue(ch + STAGES - 1);
        const int st = ch % STAGES;
        const T* Ks = stage_K(st);
        const double* Rs = stage_R(st);
        for (int r = ry; r < TRS; r += RY) {
            const double k = (double)Ks[r * CWB + cx];
            double g = 0;
            GPB_UNROLL
            for (int d = 0; d < DOS; d++) {
                const double dmd = Rs[TRS * DP + r * DOS + d];
                const double dvd = Rs[TRS * DP + TRS * DOS + r * DOS + d];
                g += dmd * (double)ap[d] + 2.0 * dvd * (double)Ks[(long)(1 + d) * TRS * CWB + r * CWB + cx];
                dA[d] += dmd * k;
            }
            const double L = g * k;
            cs += L;
            GPB_UNROLL
            for (int q = 0; q < DP; q++) {
                const double diff = (double)zr[q] - Rs[r * DP + q];
                const double t = L * diff;
                dz[q] += t;
                dl[q] += t * diff;
            }
        }
    }
    rec[c] = cs;
    GPB_UNROLL
    for (int q = 0; q < DP; q++)
        if (q < D) {
            const double il2 = exp(-2.0 * ls[q]);
            rec[(long)MP + (long)c * D + q] = -dz[q] * il2;
            rec[(long)MP + (long)MP * D + (long)c * D + q] = dl[q] * il2 * exp(-ls[q]);
        }
    GPB_UNROLL
    for (int d = 0; d < DOS; d++)
        if (d < Do) rec[(long)MP + 2L * MP * D + (long)d * MP + c] = dA[d];
}

// a8 (rank-update part), aep_models.py:493: dB[d] = sum_n dv[n,d] kfu[n,:] kfu[n,:]^T.
// Upper block-triangle of 128x128 output blocks; split over rows; partial records
//   part[((split*Do + d)*NBU + ub)*128*128 + i*128 + j]
// Each thread owns a 4 x 16 register tile (64 accumulators).  The two 128-column operand
// panels of RK saved-Kfu rows are brought in by a 3-stage cp.async ring (no register
// staging, one barrier per chunk); the row weight dv[n,d] is applied to the 4 A values
// of a row as they are read (4 multiplies per 64 FMAs).
template <typename T>
struct SyrkCfg {
    static constexpr int RK = 64 / (int)sizeof(T);   // rows per chunk: 8 KB per operand panel
    static constexpr int STAGES = 3;
    static constexpr int VEC = V16<T>::N;
    static constexpr size_t stage_bytes = 2 * (size_t)RK * 128 * sizeof(T);
    static constexpr size_t smem_bytes = STAGES * stage_bytes + STAGES * RK * sizeof(double);
};

template <typename T>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) det_syrk_kernel(const T* __restrict__ Ksave,
                                                       const double* __restrict__ dv, int n, int MP,
                                                       int Do, int rows_per_split,
                                                       double* __restrict__ part) {
    typedef typename V16<T>::type VT;
    typedef SyrkCfg<T> C;
    constexpr int VEC = C::VEC, RK = C::RK, STAGES = C::STAGES;
    constexpr int NA = 4 / VEC > 0 ? 4 / VEC : 1;    // 16B vectors of the 4 A values (fp64: 2, fp32: 1)
    constexpr int NB = 16 / VEC;                      // 16B vectors of the 16 B values
    constexpr int CPT = 2 * RK * 128 / VEC / kThreads;  // cp.async vectors per thread per chunk
    GPB_DYN_SMEM(smem);
    double* s_dv = (double*)(smem + STAGES * C::stage_bytes);
    const int tid = threadIdx.x, ty = tid >> 3, tx = tid & 7;   // 32 row groups x 8 column groups
    const int nb = MP / 128;
    int ub = blockIdx.x, bi = 0;
    while (ub >= nb - bi) { ub -= nb - bi; bi++; }
    const int bj = bi + ub;
    const int d = blockIdx.z;
    const int r_begin = blockIdx.y * rows_per_split;
    const int r_end = (r_begin + rows_per_split) < n ? (r_begin + rows_per_split) : n;
    const int nchunk = (r_end - r_begin + RK - 1) / RK;

    T acc[4][16];
    GPB_UNROLL
    for (int i = 0; i < 4; i++)
        GPB_UNROLL
        for (int j = 0; j < 16; j++) acc[i][j] = 0;

    auto issue = [&](int chunk) {
        if (chunk < nchunk) {
            const int st = chunk % STAGES;
            T* base = (T*)(smem + (size_t)st * C::stage_bytes);
            const int t0 = r_begin + chunk * RK;
            GPB_UNROLL
            for (int i = 0; i < CPT; i++) {
                int v = tid + kThreads * i;
                int which = v / (RK * 128 / VEC);
                int rem = v - which * (RK * 128 / VEC);
                int r = rem / (128 / VEC), cv = rem - r * (128 / VEC);
                long row = (long)t0 + r;
                bool ok = row < r_end;
                const T* src = Ksave + (ok ? row : (long)r_begin) * MP + (which ? bj : bi) * 128 + cv * VEC;
                cp_async16_zfill(base + (long)which * RK * 128 + (long)rem * VEC, src, ok);
            }
            if (tid < RK) {
                long row = (long)t0 + tid;
                s_dv[st * RK + tid] = row < r_end ? dv[row * Do + d] : 0.0;
            }
        }
        cp_async_commit();   // always commit (possibly empty) so the group counting stays uniform
    };

    issue(0);
    issue(1);
    for (int c = 0; c < nchunk; c++) {
        cp_async_wait<STAGES - 2>();   // chunk c has landed (this thread's copies)
        sync_threads();                // ... everybody's copies; stage (c+2)%3 is free again
        issue(c + 2);
        const int st = c % STAGES;
        const T* As = (const T*)(smem + (size_t)st * C::stage_bytes);
        const T* Bs = As + RK * 128;
        GPB_UNROLL
        for (int r = 0; r < RK; r++) {
            const T w = (T)s_dv[st * RK + r];
            T av[4], bv[16];
            if (VEC == 2) {
                GPB_UNROLL
                for (int j = 0; j < 2; j++) {
                    VecU<T> u;
                    u.v = *(const VT*)(As + r * 128 + j * 64 + ty * 2);
                    av[j * 2 + 0] = u.e[0] * w;
                    av[j * 2 + 1] = u.e[1 % VEC] * w;
                }
            } else {
                VecU<T> u;
                u.v = *(const VT*)(As + r * 128 + ty * 4);
                GPB_UNROLL
                for (int e = 0; e < 4; e++) av[e] = u.e[e % VEC] * w;
            }
            GPB_UNROLL
            for (int j = 0; j < NB; j++) {
                VecU<T> u;
                u.v = *(const VT*)(Bs + r * 128 + j * (8 * VEC) + tx * VEC);
                GPB_UNROLL
                for (int e = 0; e < VEC; e++) bv[j * VEC + e] = u.e[e];
            }
            GPB_UNROLL
            for (int i = 0; i < 4; i++)
                GPB_UNROLL
                for (int j = 0; j < 16; j++) acc[i][j] += av[i] * bv[j];
        }
    }
    const int nbu = nb * (nb + 1) / 2;
    double* out = part + (((long)blockIdx.y * Do + d) * nbu + blockIdx.x) * (128 * 128);
    GPB_UNROLL
    for (int i = 0; i < 4; i++) {
        const int oi = VEC == 2 ? ((i / 2) * 64 + ty * 2 + (i % 2)) : (ty * 4 + i);
        GPB_UNROLL
        for (int j = 0; j < 16; j++) {
            const int oj = (j / VEC) * (8 * VEC) + tx * VEC + (j % VEC);
            out[oi * 128 + oj] = (double)acc[i][j];
        }
    }
}

// fp64 tensor-core variant of the rank update (DMMA.8x8x4): the reduction index of the MMA is
// the data row, A = (dv o K)^T and B = K come from the same staged [RK x 128] panels (row stride
// 132 doubles = 4 mod 16: conflict-free fragment loads).  8 warps = 4 row-warps x 2 column-warps,
// each 32 x 64 of the 128 x 128 output block (4 x 8 C fragments); 3-stage cp.async ring.
struct SyrkMmaCfg {
    static constexpr int RK = 16;        // data rows per stage
    static constexpr int LD = 132;       // padded panel row stride (doubles)
    static constexpr int STAGES = 3;
    static constexpr size_t stage_bytes = 2 * (size_t)RK * LD * 8;
    static constexpr size_t smem_bytes = STAGES * stage_bytes + STAGES * RK * sizeof(double);
};

GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) det_syrk_mma_kernel(const double* __restrict__ Ksave,
                                                           const double* __restrict__ dv, int n, int MP,
                                                           int Do, int rows_per_split,
                                                           double* __restrict__ part) {
    typedef SyrkMmaCfg C;
    constexpr int RK = C::RK, LD = C::LD, STAGES = C::STAGES;
    constexpr int CPT = 2 * RK * 64 / kThreads;       // 16-byte copies per thread per stage
    GPB_DYN_SMEM(smem);
    double* s_dv = (double*)(smem + STAGES * C::stage_bytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int rw = warp >> 1, cw = warp & 1;
    const int nb = MP / 128;
    int ub = blockIdx.x, bi = 0;
    while (ub >= nb - bi) { ub -= nb - bi; bi++; }
    const int bj = bi + ub;
    const int d = blockIdx.z;
    const int r_begin = blockIdx.y * rows_per_split;
    const int r_end = (r_begin + rows_per_split) < n ? (r_begin + rows_per_split) : n;
    const int nchunk = (r_end - r_begin + RK - 1) / RK;

    double acc[4][8][2];
    GPB_UNROLL
    for (int i = 0; i < 4; i++)
        GPB_UNROLL
        for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0;

    auto issue = [&](int chunk) {
        if (chunk < nchunk) {
            const int st = chunk % STAGES;
            double* base = (double*)(smem + (size_t)st * C::stage_bytes);
            const int t0 = r_begin + chunk * RK;
            GPB_UNROLL
            for (int i = 0; i < CPT; i++) {
                const int v = tid + kThreads * i;
                const int which = v / (RK * 64);
                const int rem = v - which * (RK * 64);
                const int r = rem / 64, cv = rem - r * 64;
                const long row = (long)t0 + r;
                const bool ok = row < r_end;
                const double* src = Ksave + (ok ? row : (long)r_begin) * MP + (which ? bj : bi) * 128 + cv * 2;
                cp_async16_zfill(base + (long)which * RK * LD + (long)r * LD + cv * 2, src, ok);
            }
            if (tid < RK) {
                const long row = (long)t0 + tid;
                s_dv[st * RK + tid] = row < r_end ? dv[row * Do + d] : 0.0;
            }
        }
        cp_async_commit();   // always commit (possibly empty) so the group counting stays uniform
    };

    // diagonal blocks are symmetric: the warp tiles that lie entirely below the diagonal
    // (rows >= 64, columns < 64) are never read by det_syrk_finish_kernel and skip their MMAs,
    // which frees a quarter of the FP64 tensor pipe for the other six warps
    const bool skip = (bi == bj) && rw >= 2 && cw == 0;
    issue(0);
    issue(1);
    for (int c = 0; c < nchunk; c++) {
        cp_async_wait<STAGES - 2>();
        sync_threads();
        issue(c + 2);
        const int st = c % STAGES;
        const double* As = (const double*)(smem + (size_t)st * C::stage_bytes) + rw * 32 + g;
        const double* Bs = (const double*)(smem + (size_t)st * C::stage_bytes) + RK * LD + cw * 64 + g;
        if (skip) continue;
        GPB_UNROLL
        for (int k4 = 0; k4 < RK; k4 += 4) {
            const double w = s_dv[st * RK + k4 + t];
            double af[4], bf[8];
            GPB_UNROLL
            for (int i = 0; i < 4; i++) af[i] = As[(k4 + t) * LD + i * 8] * w;
            GPB_UNROLL
            for (int j = 0; j < 8; j++) bf[j] = Bs[(k4 + t) * LD + j * 8];
            GPB_UNROLL
            for (int i = 0; i < 4; i++)
                GPB_UNROLL
                for (int j = 0; j < 8; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    const int nbu = nb * (nb + 1) / 2;
    double* out = part + (((long)blockIdx.y * Do + d) * nbu + blockIdx.x) * (128 * 128);
    GPB_UNROLL
    for (int i = 0; i < 4; i++) {
        const int oi = rw * 32 + i * 8 + g;
        GPB_UNROLL
        for (int j = 0; j < 8; j++) {
            const int oj = cw * 64 + j * 8 + 2 * t;
            *(double2*)(out + oi * 128 + oj) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

// sum the row-splits of det_syrk and expand the block upper-triangle to dB[Do,M,M]
GPB_KERNEL void det_syrk_finish_kernel(const double* __restrict__ part, int nsplit, int MP, int M,
                                       int Do, double* __restrict__ dB, int tr /* blocks stored transposed */) {
    const int nb = MP / 128, nbu = nb * (nb + 1) / 2;
    const long total = (long)Do * M * M;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        int j = (int)(idx % M), i = (int)((idx / M) % M), d = (int)(idx / ((long)M * M));
        int bi = i / 128, bj = j / 128, ii = i % 128, jj = j % 128;
        // only the upper block triangle exists, and inside a diagonal block only its upper triangle
        // is guaranteed (the fp64 tensor kernel skips the warp tiles below it)
        if (bi > bj || (bi == bj && ii > jj)) { int t = bi; bi = bj; bj = t; t = ii; ii = jj; jj = t; }
        int ub = 0;
        for (int b = 0; b < bi; b++) ub += nb - b;
        ub += bj - bi;
        double s = 0;
        for (int sp = 0; sp < nsplit; sp++)
            s += part[(((long)sp * Do + d) * nbu + ub) * (128 * 128) + (tr ? jj * 128 + ii : ii * 128 + jj)];
        dB[idx] = s;
    }
}

// =========================================================================
// Moment-matched layer (a2 fused with a6 / a9)
// =========================================================================
// Pair table (n-independent): for every unordered pair p=(a>=b)
//   zh[q][p]  = (z_a + z_b)/2
//   ep[p]     = sf2^2 * exp(-sum_q (z_a-z_b)^2 / (4 l_q^2))          (kernels.py:222-226)
//   bs[d][p]  = ep[p] * ( B[d,a,b] + B[d,b,a]  (a != b)  |  B[d,a,a] )
// The n-independent factor ep is folded into the contraction weights, so the row loop only
// forms psi2' = psi2 / ep (one fewer add per pair and row) and the pair sums are rescaled once.
template <typename T>
GPB_KERNEL void mm_pair_table_kernel(const double* __restrict__ z, const double* __restrict__ ls,
                                     const double* __restrict__ sf, const double* __restrict__ B,
                                     int M, int Q, int Qt, int Do, long P, long PP,
                                     T* __restrict__ zh, T* __restrict__ ep, T* __restrict__ bs) {
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < PP;
         p += (long)gridDim.x * blockDim.x) {
        for (int q = Q; q < Qt; q++) zh[(long)q * PP + p] = 0;  // template-padded input dims
        if (p >= P) {  // padding pairs: every weight is 0
            for (int q = 0; q < Q; q++) zh[(long)q * PP + p] = 0;
            ep[p] = 0;
            for (int d = 0; d < Do; d++) bs[(long)d * PP + p] = 0;
            continue;
        }
        long a = (long)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= p) a++;
        while (a * (a + 1) / 2 > p) a--;
        long b = p - a * (a + 1) / 2;
        double e = 0;
        for (int q = 0; q < Q; q++) {
            double za = z[a * Q + q], zb = z[b * Q + q];
            zh[(long)q * PP + p] = (T)(0.5 * (za + zb));
            double dz = za - zb;
            e += dz * dz / (4.0 * exp(2.0 * ls[q]));
        }
        const double epv = exp(4.0 * sf[0] - e);   // sf2^2 exp(-sum_q dz^2/(4 l^2))
        ep[p] = (T)epv;
        for (int d = 0; d < Do; d++) {
            const double* Bd = B + (long)d * M * M;
            bs[(long)d * PP + p] = (T)(epv * (a == b ? Bd[a * M + a] : Bd[a * M + b] + Bd[b * M + a]));
        }
    }
}

template <typename T, int Q, int DOC>
struct MMCfg {
    // pairs per thread, sized to keep the per-thread pair state in registers (fp32: twice as many,
    // the fp32 path is issue bound and the per-row overhead is amortised over the pairs)
    static constexpr int RPA = (2 * Q + 2 * DOC + 2) <= GPB_MM_RP4_MAX ? 4 : ((2 * Q + 2 * DOC + 2) <= 26 ? 2 : 1);
    static constexpr int RP64 = (GPB_MM_RP64 > 0 && GPB_MM_RP64 < RPA) ? GPB_MM_RP64 : RPA;
    static constexpr int RP = sizeof(T) == 4 ? 2 * RP64 : RP64;
    static constexpr int TR = 32;             // rows per staged tile = lanes per warp
    static constexpr int PC = kThreads * RP;  // pairs per block
};

template <typename T>
struct MMArgs {
    const double* mx;  // [n, Q]
    const double* vx;  // [n, Q]
    const double* ls;  // [Q]
    const T* zh;       // [Q, PP]
    const T* ep;       // [PP]
    const T* bs;       // [Do, PP]
    const double* dv;  // [n, Do]   (backward) scaled dlogZ/dv
    int n, Qa, Do, d0; // Qa: actual input dims (<= template Q); d0: first output dim of this d-chunk
    long PP;
    int rows_per_split;
    double* rowacc;    // fwd: [n, Do] += sum_p bs[d,p] psi2[n,p] ; bwd: [n, 2Q] += {U_q, V_q}
    double* pairpart;  // bwd: [nsplit][DOC+1+Q][PP]: {dBp_d, S0, S1_q}
    int full_coef;     // bwd: 1 -> coefficient sum over ALL Do from bs (generic path when Do > DOC)
    int lam_pass;      // bwd: 1 -> this pass also produces the Lambda-dependent sums
};

// exp() in a pre-scaled domain.  The pair kernel forms xs = S * x directly (S is folded into the
// per-row constants), so that no multiply is spent on the argument reduction:
//   fp64: S = 256/ln2.  xs = 256 k + j + r, |r| <= 1/2  ->  exp(x) = 2^k * 2^(j/256) * e^(r ln2/256).
//         2^(j/256) comes from a shared-memory table that is replicated 16 times (32 KB): lane l
//         reads copy l mod 16, so the data-dependent lookups of a half warp always fall into 16
//         different 8-byte banks (the un-replicated 2048-entry table of v9 spent 6 wavefronts per
//         lookup on bank conflicts and made the kernel shared-memory bound; ncu, profiles/).
//         e^(r h) (h = ln2/256, |r h| <= 1.36e-3) is a Taylor polynomial of degree GPB_EXP_DEG
//         (default 3, truncation 1.4e-13) folded with the table value: t + (t r)(c1 + r (c2 + r c3)).
//         7 fp64 instructions (3 add, 1 mul, 3 fma) instead of ~25 for libm's exp; the 2^k scaling
//         is an integer add into the exponent field with k clamped at -1021 (deep underflow returns
//         ~1e-308 instead of 0).
//   fp32: S = log2(e); one SFU instruction (ex2.approx).
// exp_dom_n evaluates N independent arguments in lock step (stage by stage), which is what lets
// the fp64 pipe overlap the 8-deep dependency chains of different pairs / rows.
template <typename T> struct ExpDom;
template <> struct ExpDom<double> {
    static constexpr int ENT = 256, REP = GPB_EXP_REP;
    static constexpr int TAB = ENT * REP;                 // doubles of dynamic shared memory
    static constexpr double S = 256.0 / 0.693147180559945309417232;
};
template <> struct ExpDom<float> {
    static constexpr int ENT = 0, REP = 0, TAB = 0;
    static constexpr double S = 1.4426950408889634074;
};
// GPB_EXP_DEG: degree of the e^(r h) polynomial (relative truncation error 3.8e-17 / 1.4e-13 /
// 4.2e-10 for 4 / 3 / 2; one fp64 instruction per degree).  Default 3: its 1.4e-13 is the size of
// the rounding already accepted in the expanded-form exponents and 7 orders below the 1e-6 parity
// bar; measured on the B200 (tools/kbench.py): pair kernels -4.3 % (fwd) / -2.9 % (bwd) vs degree 4.
// Measured and rejected: rounding through the conversion unit (cvt.rni.s32.f64 + cvt.rn.f64.s32
// instead of the two magic-number DADDs) -- no gain, the fp64 conversions share the fp64 pipe.
#ifndef GPB_EXP_DEG
#define GPB_EXP_DEG 3
#endif
template <int N>
GPB_DEVICE void exp_dom_n(double (&x)[N], const double* __restrict__ tab, int lane16) {
    constexpr double h = 0.693147180559945309417232 / 256.0;
    constexpr double c1 = h, c2 = h * h / 2.0, c3 = h * h * h / 6.0, c4 = h * h * h * h / 24.0;
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
    double kd[N], t[N], q[N];
    int n[N];
    GPB_UNROLL
    for (int i = 0; i < N; i++) kd[i] = x[i] + magic;
    GPB_UNROLL
    for (int i = 0; i < N; i++) {
#ifndef GPB_CPU_EMU
        n[i] = __double2loint(kd[i]);
#else
        int64_t bits;
        memcpy(&bits, &kd[i], 8);
        n[i] = (int)(int32_t)(bits & 0xffffffff);
#endif
        t[i] = tab[(n[i] & 255) * ExpDom<double>::REP + lane16];
    }
    GPB_UNROLL
    for (int i = 0; i < N; i++) kd[i] -= magic;
    GPB_UNROLL
    for (int i = 0; i < N; i++) x[i] -= kd[i];             // r: exact, |r| <= 1/2
#if GPB_EXP_DEG >= 4
    GPB_UNROLL
    for (int i = 0; i < N; i++) q[i] = c4 * x[i] + c3;
    GPB_UNROLL
    for (int i = 0; i < N; i++) q[i] = q[i] * x[i] + c2;
#elif GPB_EXP_DEG == 3
    GPB_UNROLL
    for (int i = 0; i < N; i++) q[i] = c3 * x[i] + c2;
#endif
#if GPB_EXP_DEG >= 3
    GPB_UNROLL
    for (int i = 0; i < N; i++) q[i] = q[i] * x[i] + c1;
#else
    GPB_UNROLL
    for (int i = 0; i < N; i++) q[i] = c2 * x[i] + c1;
#endif
    GPB_UNROLL
    for (int i = 0; i < N; i++) kd[i] = t[i] * x[i];
    GPB_UNROLL
    for (int i = 0; i < N; i++) {
        const double p = kd[i] * q[i] + t[i];
        int k = n[i] >> 8;
        k = k < -1021 ? -1021 : k;
#ifndef GPB_CPU_EMU
        x[i] = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
        x[i] = ldexp(p, k);
#endif
    }
}
// Bit-field variant (GPB_EXP_BITS, pair kernels): the caller adds ExpBits::OFF to the row constant of
// the exponent, so that every argument that matters lies in ONE binade, xs' = xs + OFF in [2^16, 2^17):
//   xs' = 2^16 + 256 k' + j + f,   k' = k + 252 in [0, 256),  j in [0, 256),  f in [0, 1)
// and the high word of the double holds k' (mantissa bits 19:12), j (11:4) and the top bits of f.
// The argument reduction is then integer work on that word instead of two magic-number DADDs:
//   clamp      : high word = max(high word, hi(2^16))  (deep underflow -> 2^-252, sign bit included)
//   r = f - 1/2: xs' - (xs' with the mantissa below j cleared and the 1/2 bit set)      -- ONE DADD, exact
//   table      : 2^((j + 1/2)/256 - 252), replicated like the table of exp_dom_n
//   e^(r h)    : 1 + r (c1 + r (c2 + r c3))  -- every FMA has a constant operand (full issue rate)
//   2^k'       : integer add of k' << 20 into the exponent field of the product
// 5 fp64 instructions (1 add, 3 fma, 1 mul) instead of 7.  The price is the resolution of the sum:
// ulp(xs') = 2^-36, i.e. 2^-37 * ln2/256 = 2e-14 relative per rounding of the 2Q-term exponent, the
// size of the polynomial's own 1.4e-13 truncation.
struct ExpBits {
    static constexpr double OFF = 65536.0 + 252.0 * 256.0;
    static constexpr int HI_MIN = 0x40F00000;     // high word of 2^16
};
GPB_DEVICE double exp_bits_table(int j) { return exp2(((double)j + 0.5) * (1.0 / 256.0) - 252.0); }
GPB_DEVICE int exp_hi(double x) {
#ifndef GPB_CPU_EMU
    return __double2hiint(x);
#else
    int64_t b;
    memcpy(&b, &x, 8);
    return (int)(int32_t)(b >> 32);
#endif
}
GPB_DEVICE double exp_sethi(double x, int hi) {
#ifndef GPB_CPU_EMU
    return __hiloint2double(hi, __double2loint(x));
#else
    int64_t b;
    memcpy(&b, &x, 8);
    b = (int64_t)(((uint64_t)(uint32_t)hi << 32) | ((uint64_t)b & 0xffffffffull));
    memcpy(&x, &b, 8);
    return x;
#endif
}
// The two (a & imm) | reg combinations are written as explicit LOP3s: two immediates cannot be
// encoded and the compiler otherwise spends two instructions on each (the integer pipe is as narrow as
// the fp64 pipe on this chip); `half_bit` is the constant 8 kept in a register by the caller.
template <int N>
GPB_DEVICE void exp_dom_bits_n(double (&x)[N], const double* __restrict__ tab, int lane16, int half_bit) {
    constexpr double h = 0.693147180559945309417232 / 256.0;
    constexpr double c1 = h, c2 = h * h / 2.0, c3 = h * h * h / 6.0;
    int hi[N];
    double t[N], q[N];
    GPB_UNROLL
    for (int i = 0; i < N; i++) {
        hi[i] = exp_hi(x[i]);
        hi[i] = hi[i] < ExpBits::HI_MIN ? ExpBits::HI_MIN : hi[i];
    }
    GPB_UNROLL
    for (int i = 0; i < N; i++) {
#ifndef GPB_CPU_EMU
        int idx;       // (j << 4) | lane16: element index into the [256][16] table
        asm("lop3.b32 %0, %1, 0xff0, %2, 0xEA;" : "=r"(idx) : "r"(hi[i]), "r"(lane16));
        t[i] = tab[idx];
#else
        t[i] = tab[(hi[i] & 0xff0) | lane16];
#endif
    }
    GPB_UNROLL
    for (int i = 0; i < N; i++) {
#ifndef GPB_CPU_EMU
        int th;        // high word of xs' with the mantissa below j cleared and the 1/2 bit set
        asm("lop3.b32 %0, %1, 0xFFFFFFF0, %2, 0xEA;" : "=r"(th) : "r"(hi[i]), "r"(half_bit));
#else
        const int th = (hi[i] & (int)0xFFFFFFF0) | half_bit;
#endif
        x[i] = exp_sethi(x[i], hi[i]) - exp_sethi(0.0, th);   // r in [-1/2, 1/2), exact
    }
    GPB_UNROLL
    for (int i = 0; i < N; i++) q[i] = c3 * x[i] + c2;
    GPB_UNROLL
    for (int i = 0; i < N; i++) q[i] = q[i] * x[i] + c1;
    GPB_UNROLL
    for (int i = 0; i < N; i++) q[i] = q[i] * x[i] + 1.0;
    GPB_UNROLL
    for (int i = 0; i < N; i++) {
        const double p = t[i] * q[i];
#ifndef GPB_CPU_EMU
        int nh;        // exponent field += k' (one multiply-add; the compiler's own form takes three)
        asm("mad.lo.s32 %0, %1, 256, %2;" : "=r"(nh) : "r"(hi[i] & 0xFF000), "r"(exp_hi(p)));
        x[i] = exp_sethi(p, nh);
#else
        x[i] = exp_sethi(p, exp_hi(p) + ((hi[i] & 0xFF000) << 8));
#endif
    }
}
template <int N>
GPB_DEVICE void exp_dom_n(float (&x)[N], const double*, int) {
    GPB_UNROLL
    for (int i = 0; i < N; i++) {
#ifndef GPB_CPU_EMU
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
        x[i] = y;
#else
        x[i] = exp2f(x[i]);
#endif
    }
}

template <bool BITS, int N>
GPB_DEVICE void pair_exp(double (&x)[N], const double* tab, int lane16, int half_bit) {
    if (BITS) exp_dom_bits_n<N>(x, tab, lane16, half_bit);
    else exp_dom_n<N>(x, tab, lane16);
}
template <bool BITS, int N>
GPB_DEVICE void pair_exp(float (&x)[N], const double* tab, int lane16, int) { exp_dom_n<N>(x, tab, lane16); }

// Geometry of the per-warp transposition buffer of the pair kernel: every lane stores the NS row
// sums of its pairs for HT consecutive rows (one 16-byte padded line of 32 lane records per row);
// then lane L adds up row L % HT over a slice of HT lanes.  One store + one load per value
// instead of the 5-level select/shuffle cascade of v4-v10 (which cost ~8 non-fp64 instructions per
// pair and row -- the fp64 pipe only stays busy if at most one other instruction issues per fp64
// instruction, see profiles/).  HT is sized so that the 8 warps use <= 64 KB.
template <typename T, int NS>
struct RowXpose {
    // a row line holds the NS values of the 32 lanes in 16-byte vectors, vector-major:
    //   [vector c][lane] -- stores of a quarter warp cover 128 contiguous bytes, and so do the
    //   reads (8 lanes read the same source lane of 8 different rows; row stride = 16 mod 128)
    static constexpr int VW = 16 / (int)sizeof(T);         // values per 16-byte vector
    static constexpr int NV = (NS + VW - 1) / VW;          // vectors per lane record
    static constexpr int kRecB = NV * 16;                  // bytes of one lane record
    static constexpr int kRowB = 32 * kRecB + 16;          // padded row stride
    static constexpr int kRaw = 65536 / (8 * kRowB);
    static constexpr int HT = kRaw >= 32 ? 32 : (kRaw >= 16 ? 16 : (kRaw >= 8 ? 8 : (kRaw >= 4 ? 4 : 2)));
    static constexpr int kWarpB = HT * kRowB;
    static constexpr int kBytes = 8 * kWarpB;
};

// a2+a6 (forward) / a2+a9 (backward) over unordered pairs.  Each thread owns RP pairs for
// the whole kernel (their constants and accumulators live in registers); rows are staged
// 32 at a time in shared memory as one 16-byte aligned record per row and broadcast.
// psi2'[n,p] = cn[n] exp(-sum_q (mu_nq - zh_pq)^2 c2_nq)  (= psi2 / ep[p]) is formed in registers
// and consumed immediately: the N x M x M tensor never exists in memory.
//   forward : rowacc[n,d]  += sum_p bs[d,p] psi2'[n,p]                  (aep_models.py:196-198)
//   backward: Lam[n,p] = (sum_d dv[n,d] bs[d,p]) psi2'[n,p]             (aep_models.py:243, kernels.py:415-419)
//             rowacc[n,:]  += {sum_p Lam zh_q, sum_p Lam zh_q^2}
//               (sum_p Lam itself equals sum_d dv[n,d] * forward rowacc[n,d]: not recomputed)
//             pair sums     : dBp[d,p] = ep[p] sum_n dv[n,d] psi2'[n,p]  (aep_models.py:240)
//                             S0[p] = sum_n Lam = sum_d bs[d,p] sum_n dv[n,d] psi2'  (from the dBp sums)
//                             S1[p,q] = sum_n Lam c2_nq (mu_nq - zh_pq)
// The exponent is formed in the ExpDom<T>-scaled domain.  fp64 forward uses the expanded form
// xs = a0_n + sum_q (b_nq zh_pq + c_nq zh_pq^2) (2Q fma); the backward needs c2 (mu - zh) anyway.
// Two rows are processed per loop trip when Q <= 4 (independent dependency chains for the fp64 pipe).
// The per-row sums are reduced over the warp through a shared-memory transposition (RowXpose), over
// the 8 warps through shared memory, and over the pair chunks (blocks) by one fp64 atomic per value.
// GEN = true: generic multi-pass path for Do > DOC (runtime full_coef / lam_pass flags);
// GEN = false (Do <= DOC): single pass, Lambda sums always on, coefficient from registers.
// (the body is a device function so that two kernels with different register budgets can share it)
template <typename T, int Q, int DOC, bool BWD, bool GEN, int NRF>
GPB_DEVICE void mm_pairs_body(MMArgs<T> a) {
    typedef MMCfg<T, Q, DOC> C;
    constexpr int RP = C::RP, TR = C::TR;
    constexpr int NS = BWD ? 2 * Q : DOC;
    constexpr bool kGen = BWD && GEN;
    constexpr bool kExpand = !BWD && sizeof(T) == 8;
    constexpr bool kBits = GPB_EXP_BITS != 0 && sizeof(T) == 8;     // bit-field argument reduction
    constexpr int VW = 16 / (int)sizeof(T);
    constexpr int RLraw = 1 + 2 * Q + (BWD ? DOC : 0);
    constexpr int RL = (RLraw + VW - 1) / VW * VW;       // row record length (16-byte multiple)
    constexpr int kTab = ExpDom<T>::TAB;
    // rows per loop trip (wide inputs: one, register budget)
    constexpr int NR = Q <= 4 ? ((!BWD && sizeof(T) == 8 && Q <= 2 && DOC <= 2) ? NRF : 2) : 1;
    constexpr double kS = ExpDom<T>::S;
    // Row tiles are double buffered (tile t+1 is staged while tile t is consumed) and so is the
    // cross-warp staging of the row sums, which leaves ONE barrier per tile.  For wide inputs the
    // staging would not fit the static 48 KB; then every warp issues its own atomics.
    constexpr int kBufs = kGen ? 1 : 2;   // the generic path restages in place
    constexpr size_t kFixed = kBufs * TR * RL * sizeof(T) + (kGen ? TR * 64 * 8 : 8) + Q * 8;   // static
    constexpr bool WARP_ATOMICS = kFixed + 2 * 8 * TR * NS * sizeof(T) > 46 * 1024;
    GPB_SHARED GPB_ALIGN16 T s_rec[kBufs][TR * RL];
    GPB_SHARED double s_dvall[kGen ? TR * 64 : 1];  // generic path (single buffered): all Do (<= 64) per row
    GPB_SHARED GPB_ALIGN16 T s_red[2][WARP_ATOMICS ? 1 : 8 * TR * NS];
    GPB_SHARED double s_l2[Q];
    typedef RowXpose<T, NS> X;
    constexpr int HT = X::HT;
    GPB_DYN_SMEM(dsm);      // [ replicated exp table (fp64: 32 KB) | 8 x per-warp transposition buffers ]
    double* s_tab = (double*)dsm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long pbase = (long)blockIdx.x * C::PC;
    const long PP = a.PP;
    const int Do = a.Do;

    T zh[RP][Q], zh2[RP][Q], bs[RP][DOC];
    T accB[RP][DOC], accS0[RP], accS1[RP][Q];
    GPB_UNROLL
    for (int j = 0; j < RP; j++) {
        const long p = pbase + j * kThreads + tid;
        GPB_UNROLL
        for (int q = 0; q < Q; q++) {
            zh[j][q] = a.zh[(long)q * PP + p];
            zh2[j][q] = zh[j][q] * zh[j][q];
        }
        GPB_UNROLL
        for (int d = 0; d < DOC; d++) {
            bs[j][d] = (a.d0 + d < Do) ? a.bs[(long)(a.d0 + d) * PP + p] : (T)0;
            accB[j][d] = 0;
        }
        accS0[j] = 0;
        GPB_UNROLL
        for (int q = 0; q < Q; q++) accS1[j][q] = 0;
    }
    if (tid < Q) s_l2[tid] = tid < a.Qa ? exp(2.0 * a.ls[tid]) : 1.0;
    if (kTab > 0)
        for (int i = tid; i < kTab; i += kThreads)
            s_tab[i] = kBits ? exp_bits_table(i / (ExpDom<T>::REP > 0 ? ExpDom<T>::REP : 1))
                             : exp2((double)(i / ExpDom<T>::REP) * (1.0 / (ExpDom<T>::ENT > 0 ? ExpDom<T>::ENT : 1)));
    const int lane16 = lane & ((ExpDom<T>::REP > 0 ? ExpDom<T>::REP : 1) - 1);
    const int half_bit = 8 + (a.n < 0);      // = 8, opaque to constant folding (LOP3 operand, exp_dom_bits_n)
    unsigned char* xw = dsm + kTab * sizeof(double) + warp * X::kWarpB;   // this warp's buffer

    const int r_begin = blockIdx.y * a.rows_per_split;
    const int r_end = (r_begin + a.rows_per_split) < a.n ? (r_begin + a.rows_per_split) : a.n;

    // stage one tile: 4 lanes of every warp take one row each (spreads the fp64 div/sqrt/log
    // evenly over the warps).  c2 = 1/(2S + l^2); lcn = sum_q log sqrt(l^2 c2)  (kernels.py:188-190)
    // record: [0] = kS lcn, then (mu_q, kS c2_q), then dv_d -- or, expanded forward:
    //         [0] = kS (lcn - sum c2 mu^2), then (2 kS c2_q mu_q, -kS c2_q)
    auto stage = [&](int buf, int t0) {
        if (lane < 4) {
            const int row = warp * 4 + lane;
            const bool ok = (t0 + row) < r_end;
            T* rec = &s_rec[kGen ? 0 : buf][row * RL];
            double lcn = 0.0, a0 = 0.0;
            for (int q = 0; q < Q; q++) {
                double mu = 0, c2 = 0;
                if (ok && q < a.Qa) {
                    mu = a.mx[(long)(t0 + row) * a.Qa + q];
                    const double lq = s_l2[q];
                    c2 = 1.0 / (2.0 * a.vx[(long)(t0 + row) * a.Qa + q] + lq);
                    lcn += 0.5 * log(lq * c2);
                }
                const double c2s = c2 * kS;
                if (kExpand) {
                    rec[1 + q] = (T)(2.0 * c2s * mu);
                    rec[1 + Q + q] = (T)(-c2s);
                    a0 -= c2s * mu * mu;
                } else {
                    rec[1 + q] = (T)mu;
                    rec[1 + Q + q] = (T)c2s;
                }
            }
            // rows past the end: psi2' = exp(-1e5 + ...) ~ 0 (1e-308 in fp64) and their dv is 0
            rec[0] = (T)((ok ? lcn : -1.0e5) * kS + a0 + (kBits ? ExpBits::OFF : 0.0));
            if (BWD) {
                for (int d = 0; d < DOC; d++)
                    rec[1 + 2 * Q + d] =
                        (ok && a.d0 + d < Do) ? (T)a.dv[(long)(t0 + row) * Do + a.d0 + d] : (T)0;
                if (kGen && a.full_coef)
                    for (int d = 0; d < Do; d++)
                        s_dvall[row * 64 + d] = ok ? a.dv[(long)(t0 + row) * Do + d] : 0.0;
            }
        }
    };

    sync_threads();                 // s_l2 visible
    if (!kGen && r_begin < r_end) stage(0, r_begin);
    sync_threads();
    int buf = 0;
    for (int t0 = r_begin; t0 < r_end; t0 += TR, buf ^= 1) {
        const int tv = (r_end - t0) < TR ? (r_end - t0) : TR;
        if (kGen) {                 // generic path: s_dvall is single buffered -> stage in place
            sync_threads();
            stage(buf, t0);
            sync_threads();
        } else if (t0 + TR < r_end) {
            stage(buf ^ 1, t0 + TR);
        }
        // NR rows against this thread's RP pairs; v[i] = this thread's part of the sums of row r0+i.
        // All NR*RP exponents are formed first, then exponentiated in lock step, then consumed.
        auto rows_body = [&](int r0, T (&v)[NR][NS]) {
            constexpr int RCN = DOC > 4 ? 1 + 2 * Q : RL;   // wide layers read dv from smem in place
            T rc[NR][RCN], x[NR * RP], t[NR][RP][Q];
            GPB_UNROLL
            for (int i = 0; i < NR; i++) {
                GPB_UNROLL
                for (int k = 0; k < RCN; k++) rc[i][k] = s_rec[kGen ? 0 : buf][(r0 + i) * RL + k];
                GPB_UNROLL
                for (int s = 0; s < NS; s++) v[i][s] = 0;
            }
            GPB_UNROLL
            for (int i = 0; i < NR; i++) {
                GPB_UNROLL
                for (int j = 0; j < RP; j++) {
                    T xx = rc[i][0];
                    if (kExpand) {
                        GPB_UNROLL
                        for (int q = 0; q < Q; q++) {
                            xx += rc[i][1 + q] * zh[j][q];
                            xx += rc[i][1 + Q + q] * zh2[j][q];
                            t[i][j][q] = 0;
                        }
                    } else {
                        GPB_UNROLL
                        for (int q = 0; q < Q; q++) {
                            const T diff = rc[i][1 + q] - zh[j][q];
                            t[i][j][q] = diff * rc[i][1 + Q + q];
                            xx -= t[i][j][q] * diff;
                        }
                    }
                    x[i * RP + j] = xx;
                }
            }
            pair_exp<kBits>(x, s_tab, lane16, half_bit);
            if (BWD && !GEN && DOC <= 4) {
                // Register-operand bandwidth: an fp64 FMA with three distinct register operands
                // issues at ~2/3 rate on this chip, one that shares an operand with its predecessor
                // at full rate (tools/probe).  The accumulations are therefore emitted in groups
                // whose consecutive FMAs share an operand and never hit the same accumulator.
                T lam[NR][RP];
                GPB_UNROLL
                for (int i = 0; i < NR; i++)
                    GPB_UNROLL
                    for (int j = 0; j < RP; j++) {
                        T coef = 0;
                        GPB_UNROLL
                        for (int d = 0; d < DOC; d++) coef += rc[i][(1 + 2 * Q + d) < RCN ? (1 + 2 * Q + d) : 0] * bs[j][d];
                        lam[i][j] = coef * x[i * RP + j];
                    }
                GPB_UNROLL
                for (int d = 0; d < DOC; d++)
                    GPB_UNROLL
                    for (int i = 0; i < NR; i++)
                        GPB_UNROLL
                        for (int j = 0; j < RP; j++)      // shares dv[i][d]
                            accB[j][d] += rc[i][(1 + 2 * Q + d) < RCN ? (1 + 2 * Q + d) : 0] * x[i * RP + j];
                GPB_UNROLL
                for (int i = 0; i < NR; i++)
                    GPB_UNROLL
                    for (int j = 0; j < RP; j++)
                        GPB_UNROLL
                        for (int q = 0; q < Q; q++)       // shares lam[i][j]
                            accS1[j][q] += lam[i][j] * t[i][j][q];
                GPB_UNROLL
                for (int j = 0; j < RP; j++)
                    GPB_UNROLL
                    for (int q = 0; q < Q; q++) {
                        GPB_UNROLL
                        for (int i = 0; i < NR; i++) v[i][q] += lam[i][j] * zh[j][q];        // shares zh
                        GPB_UNROLL
                        for (int i = 0; i < NR; i++) v[i][Q + q] += lam[i][j] * zh2[j][q];   // shares zh2
                    }
                return;
            }
            GPB_UNROLL
            for (int i = 0; i < NR; i++) {
                GPB_UNROLL
                for (int j = 0; j < RP; j++) {
                    const T psi2 = x[i * RP + j];
                    if (!BWD) {
                        GPB_UNROLL
                        for (int d = 0; d < DOC; d++) v[i][d] += bs[j][d] * psi2;
                    } else {
                        T coef = 0;
                        GPB_UNROLL
                        for (int d = 0; d < DOC; d++) {
                            const T dvd = DOC > 4 ? s_rec[kGen ? 0 : buf][(r0 + i) * RL + 1 + 2 * Q + d]
                                                  : rc[i][(1 + 2 * Q + d) < RCN ? (1 + 2 * Q + d) : 0];
                            accB[j][d] += dvd * psi2;
                            coef += dvd * bs[j][d];
                        }
                        if (!GEN || a.lam_pass) {
                            if (kGen && a.full_coef) {
                                // coef holds the first DOC output dims (the Lambda pass is the pass
                                // with d0 == 0); the remaining ones come from the L1-resident table
                                const long p = pbase + j * kThreads + tid;
                                for (int d = DOC; d < Do; d++)
                                    coef += (T)s_dvall[(r0 + i) * 64 + d] * a.bs[(long)d * PP + p];
                            }
                            const T lam = coef * psi2;
                            if (kGen) accS0[j] += lam;
                            GPB_UNROLL
                            for (int q = 0; q < Q; q++) {
                                accS1[j][q] += lam * t[i][j][q];
                                v[i][q] += lam * zh[j][q];
                                v[i][Q + q] += lam * zh2[j][q];
                            }
                        }
                    }
                }
            }
        };
        T* red = s_red[buf];
        const bool emit = !BWD || !GEN || a.lam_pass;
        GPB_UNROLL_N(1)
        for (int r0 = 0; r0 < TR; r0 += NR) {
            T v[NR][NS];
            rows_body(r0, v);
            GPB_UNROLL
            for (int i = 0; i < NR; i++) {
                unsigned char* dst = xw + ((r0 + i) & (HT - 1)) * X::kRowB + lane * 16;
                GPB_UNROLL
                for (int c = 0; c < X::NV; c++) {
                    VecU<T> u;
                    GPB_UNROLL
                    for (int e = 0; e < X::VW; e++) u.e[e] = (c * X::VW + e < NS) ? v[i][c * X::VW + e] : (T)0;
                    *(typename V16<T>::type*)(dst + c * 512) = u.v;
                }
            }
            if (((r0 + NR) & (HT - 1)) == 0) {
                // HT rows complete: lane L adds row L % HT over lanes [slice*HT, slice*HT + HT)
                sync_warp();
                const int xr = lane & (HT - 1), xs = lane / HT;
                const unsigned char* src = xw + xr * X::kRowB + (xs * HT) * 16;
                T acc[X::NV * X::VW];
                GPB_UNROLL
                for (int s = 0; s < X::NV * X::VW; s++) acc[s] = 0;
                GPB_UNROLL_N(4)
                for (int k = 0; k < HT; k++) {
                    GPB_UNROLL
                    for (int c = 0; c < X::NV; c++) {
                        VecU<T> u;
                        u.v = *(const typename V16<T>::type*)(src + c * 512 + k * 16);
                        GPB_UNROLL
                        for (int e = 0; e < X::VW; e++) acc[c * X::VW + e] += u.e[e];
                    }
                }
                GPB_UNROLL
                for (int m = HT; m < 32; m <<= 1) {
                    GPB_UNROLL
                    for (int s = 0; s < NS; s++) acc[s] += shfl_xor(acc[s], m);
                }
                const int row = r0 + NR - HT + xr;       // row of the 32-row tile
                if (WARP_ATOMICS) {
                    if (emit && lane < HT && row < tv) {
                        GPB_UNROLL
                        for (int s = 0; s < NS; s++) {
                            if (BWD)
                                atomic_add(a.rowacc + (long)(t0 + row) * NS + s, (double)acc[s]);
                            else if (a.d0 + s < Do)
                                atomic_add(a.rowacc + (long)(t0 + row) * Do + a.d0 + s, (double)acc[s]);
                        }
                    }
                } else if (lane < HT) {
                    GPB_UNROLL
                    for (int s = 0; s < NS; s++) red[(warp * TR + row) * NS + s] = acc[s];
                }
                sync_warp();                             // buffer free for the next HT rows
            }
        }
        // rows of this tile summed per warp: add the 8 warps, then one atomic per value
        if (WARP_ATOMICS) {
            sync_threads();         // next tile staged; this tile's buffers free
        } else {
            sync_threads();         // the only barrier per tile (also publishes the staged next tile)
            if (emit)
                for (int i = tid; i < TR * NS; i += kThreads) {
                    const int r = i / NS, s = i - r * NS;
                    if (r < tv) {
                        double acc = 0;
                        for (int w = 0; w < 8; w++) acc += (double)red[(w * TR + r) * NS + s];
                        if (BWD)
                            atomic_add(a.rowacc + (long)(t0 + r) * NS + s, acc);
                        else if (a.d0 + s < Do)
                            atomic_add(a.rowacc + (long)(t0 + r) * Do + a.d0 + s, acc);
                    }
                }
        }
    }
    if (BWD) {
        double* rec = a.pairpart + (long)blockIdx.y * (DOC + 1 + Q) * PP;
        GPB_UNROLL
        for (int j = 0; j < RP; j++) {
            const long p = pbase + j * kThreads + tid;
            const double epv = (double)a.ep[p];
            double s0 = (double)accS0[j];
            GPB_UNROLL
            for (int d = 0; d < DOC; d++) {
                rec[(long)d * PP + p] = epv * (double)accB[j][d];
                if (!kGen) s0 += (double)bs[j][d] * (double)accB[j][d];
            }
            rec[(long)DOC * PP + p] = s0;
            GPB_UNROLL
            for (int q = 0; q < Q; q++) rec[(long)(DOC + 1 + q) * PP + p] = (double)accS1[j][q] * (1.0 / kS);
        }
    }
}

template <typename T, int Q, int DOC, bool BWD, bool GEN>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) mm_pairs_kernel(MMArgs<T> a) {
    mm_pairs_body<T, Q, DOC, BWD, GEN, 2>(a);
}
// fp64 forward with <= 4 output dims: ptxas' own heuristic stops at 94 registers although two CTAs per
// SM leave 128; asking for exactly two resident CTAs lets it keep more of the lock-step exp chains in
// registers (measured -4 % on the B200).  The same bound makes the backward kernel no faster and
// spills in fp32 instantiations, hence a separate entry point.
template <int Q, int DOC>
GPB_KERNEL void GPB_LAUNCH_BOUNDS2(256, 2) mm_pairs_fwd64_kernel(MMArgs<double> a) {
    mm_pairs_body<double, Q, DOC, false, false, GPB_MM_NR_FWD>(a);
}

// Forward contraction for WIDE output layers (Do > 4, e.g. SGPLVM with Do = 50), roles swapped:
// a thread owns a ROW (its Do partial sums live in registers, no cross-lane reduction at all) and
// the CTA's chunk of PCW pairs -- zh, zh^2 and the Do weights of every pair -- is staged once in
// shared memory and broadcast.  psi2' is evaluated once per (row, pair) instead of once per
// 4-output pass of the pair-owner kernel (13 passes at Do = 50).
//   rowacc[n,d] += sum_{p in chunk} bs[d,p] psi2'[n,p]        (aep_models.py:196-198)
template <typename T, int Q, int DOW>
struct MMWideCfg {
    static constexpr int PCW = sizeof(T) == 8 ? (DOW >= 64 ? 192 : 256) : 256;   // pairs per CTA
    static constexpr int REC = 2 * Q + DOW;                                        // per-pair record
    static constexpr size_t smem_bytes = sizeof(double) * ExpDom<T>::TAB + sizeof(T) * (size_t)PCW * REC;
};

template <typename T, int Q, int DOW>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) mm_fwd_wide_kernel(MMArgs<T> a) {
    typedef MMWideCfg<T, Q, DOW> C;
    constexpr int PCW = C::PCW, REC = C::REC;
    constexpr double kS = ExpDom<T>::S;
    constexpr int kTab = ExpDom<T>::TAB;
    GPB_DYN_SMEM(dsm);
    double* s_tab = (double*)dsm;
    T* s_pair = (T*)(dsm + kTab * sizeof(double));          // [PCW][REC]: zh[Q] | zh2[Q] | bs[DOW]
    GPB_SHARED double s_l2[Q];
    const int tid = threadIdx.x, lane16 = tid & ((ExpDom<T>::REP > 0 ? ExpDom<T>::REP : 1) - 1);
    const long pbase = (long)blockIdx.x * PCW;
    const long PP = a.PP;
    const int Do = a.Do;
    if (tid < Q) s_l2[tid] = tid < a.Qa ? exp(2.0 * a.ls[tid]) : 1.0;
    if (kTab > 0)
        for (int i = tid; i < kTab; i += kThreads)
            s_tab[i] = exp2((double)(i / ExpDom<T>::REP) * (1.0 / (ExpDom<T>::ENT > 0 ? ExpDom<T>::ENT : 1)));
    for (int i = tid; i < PCW * REC; i += kThreads) {
        const int p = i / REC, k = i - p * REC;
        const long pg = pbase + p;
        T v = 0;
        if (pg < PP) {
            if (k < Q) v = a.zh[(long)k * PP + pg];
            else if (k < 2 * Q) { const T z = a.zh[(long)(k - Q) * PP + pg]; v = z * z; }
            else if (k - 2 * Q < Do) v = a.bs[(long)(k - 2 * Q) * PP + pg];
        }
        s_pair[i] = v;
    }
    sync_threads();
    const int r_begin = blockIdx.y * a.rows_per_split;
    const int r_end = (r_begin + a.rows_per_split) < a.n ? (r_begin + a.rows_per_split) : a.n;
    for (int row = r_begin + tid; row < r_end; row += kThreads) {
        // row constants of the expanded exponent: xs = a0 + sum_q (b_q zh_q + c_q zh_q^2)
        T b[Q], c[Q];
        double lcn = 0, a0d = 0;
        GPB_UNROLL
        for (int q = 0; q < Q; q++) {
            double mu = 0, c2 = 0;
            if (q < a.Qa) {
                mu = a.mx[(long)row * a.Qa + q];
                const double lq = s_l2[q];
                c2 = 1.0 / (2.0 * a.vx[(long)row * a.Qa + q] + lq);
                lcn += 0.5 * log(lq * c2);
            }
            const double c2s = c2 * kS;
            b[q] = (T)(2.0 * c2s * mu);
            c[q] = (T)(-c2s);
            a0d -= c2s * mu * mu;
        }
        const T a0 = (T)(lcn * kS + a0d);
        T acc[DOW];
        GPB_UNROLL
        for (int d = 0; d < DOW; d++) acc[d] = 0;
        GPB_UNROLL_N(1)
        for (int p = 0; p < PCW; p += 2) {
            T x[2];
            GPB_UNROLL
            for (int u = 0; u < 2; u++) {
                const T* rec = s_pair + (p + u) * REC;
                T xx = a0;
                GPB_UNROLL
                for (int q = 0; q < Q; q++) {
                    xx += b[q] * rec[q];
                    xx += c[q] * rec[Q + q];
                }
                x[u] = xx;
            }
            exp_dom_n<2>(x, s_tab, lane16);
            GPB_UNROLL
            for (int u = 0; u < 2; u++) {
                const T* rec = s_pair + (p + u) * REC + 2 * Q;
                GPB_UNROLL
                for (int d = 0; d < DOW; d++) acc[d] += rec[d] * x[u];
            }
        }
        GPB_UNROLL
        for (int d = 0; d < DOW; d++)
            if (d < Do) atomic_add(a.rowacc + (long)row * Do + d, (double)acc[d]);
    }
}

// psi1 forward + moment matching epilogue (kernels.py:214-218,233; aep_models.py:195-198):
//   psi1[n,m] = sf2 prod_q sqrt(l_q^2 c1_nq) exp(-1/2 sum_q (mu_nq - z_mq)^2 c1_nq),  c1 = 1/(S + l^2)
//   mout[n,d] = sum_m A[d,m] psi1[n,m] ;  vout[n,d] = sf2 + vacc[n,d] - mout^2
// Row tiles of 32.  Phase A: thread per pseudo-point column (z_m in registers) fills the
// psi1 tile in shared memory and -- coalesced -- the optional psi1 save buffer that both
// backward kernels stream instead of re-evaluating the exponentials.  Phase B: each warp takes
// 4 rows, lanes stride over the columns, warp reduction per (row, d).
// Backward contraction for WIDE output layers (Do > 4, fp64) on the FP64 tensor cores.
// With Do in the tens, everything the pair backward does after generating psi2' is GEMM-shaped
// (rows r of a 32-row tile, pairs p of the CTA's 64-pair chunk, output dims d, 2Q moment columns s):
//   G1  Lraw[r,p] = sum_d dv[r,d] bs[d,p]          -> Lam = Lraw * psi2'   (aep_models.py:243, kernels.py:415-419)
//   G2  dBp[d,p] += sum_r dv[r,d] psi2'[r,p]                               (aep_models.py:240)
//   G3  W[p,s]   += sum_r Lam[r,p] R[r,s],  R = [c2 mu | c2]   -> S1[p,q] = W[p,q] - zh[p,q] W[p,Q+q]
//   G4  V[r,s]    = sum_p Lam[r,p] Z[p,s],  Z = [zh | zh^2]    -> rowacc[r,:] (one atomic per value)
// so a thread never holds Do accumulators: psi2' is evaluated ONCE per (row, pair) by the SIMT phase
// (expanded-form exponent + table exp, one pair column and 8 rows per thread) into shared memory
// and the four products run as DMMA.8x8x4 tiles (fragments: a = A[g][t], b = B[t][g], c = C[g][2t..]).
// All shared-memory strides are 4 mod 16 doubles, so every fragment load is conflict free.
// dBp / W accumulate in registers over the CTA's rows; S0 = sum_d bs dBp is derived at the end.
// Replaces two passes of the 32-outputs-per-thread SIMT kernel for SGPLVM-like layers (Do = 50).
template <int Q>
struct MMWideMma {
    static constexpr int PCW = 64, TR = 32, LDP = PCW + 4, LDV = 68, LDZ = 20;
    static constexpr int QB = (2 * Q + 7) / 8;          // 8-column blocks of the moment columns
    static constexpr int RL = (1 + 2 * Q + 1) / 2 * 2;  // row-constant record (even length)
    static constexpr size_t n_tab = ExpDom<double>::TAB;
    static constexpr size_t n_bs = 64 * LDP, n_dv = TR * LDV, n_psi = TR * LDP, n_lam = TR * LDP;
    static constexpr size_t n_Z = PCW * LDZ, n_R = TR * LDZ, n_rc = TR * RL;
    static constexpr size_t n_stage = n_dv + n_R + n_rc;     // per row tile, double buffered
    static constexpr size_t smem_bytes = 8 * (n_tab + n_bs + n_psi + n_lam + n_Z + 2 * n_stage);
    static_assert(2 * Q <= 16 && Q <= 8, "moment columns must fit two 8-column blocks");
    static_assert(n_psi + n_lam >= 64 * LDP, "the dBp staging aliases psi | lam");
};

// NW = warps per CTA (8 or 16).  With 16 warps the tiles of every product are spread over twice as
// many warps (4 per scheduler instead of 2 hide the DMMA / shared-memory latencies between the
// barriers); G3 and G4 then run side by side on the two warp halves.
template <int Q, int NW>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(NW * 32) mm_bwd_wide_mma_kernel(MMArgs<double> a, int DOP8) {
    typedef MMWideMma<Q> C;
    constexpr int PCW = C::PCW, TR = C::TR, LDP = C::LDP, LDV = C::LDV, LDZ = C::LDZ, QB = C::QB, RL = C::RL;
    constexpr int NT = NW * 32;
    constexpr int RPT = TR * PCW / NT, RSTEP = NT / PCW;     // SIMT phase: rows per thread / row stride
    constexpr int DVT = TR * 64 / NT;                        // dv values staged per thread
    constexpr int NJ = 32 / NW;                              // G1: column blocks per warp
    constexpr int DSTEP = NW / 8, NI = 8 / DSTEP;            // G2: d-blocks interleaved over warp halves
    constexpr double kS = ExpDom<double>::S;
    GPB_DYN_SMEM(dsm);
    double* s_tab = (double*)dsm;
    double* s_bs = s_tab + C::n_tab;     // [DOP8][LDP]   bs[d, p]
    double* s_psi = s_bs + C::n_bs;      // [TR][LDP]     psi2'[r, p]
    double* s_lam = s_psi + C::n_psi;    // [TR][LDP]     Lam[r, p]
    double* s_Z = s_lam + C::n_lam;      // [PCW][LDZ]    zh | zh^2
    double* s_stage = s_Z + C::n_Z;      // 2 x { dv[TR][LDV] | R[TR][LDZ] = c2 mu | c2 | rc[TR][RL] }
    GPB_SHARED double s_l2[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int lane16 = lane & (ExpDom<double>::REP - 1);
    const long pbase = (long)blockIdx.x * PCW;
    const long PP = a.PP;
    const int Do = a.Do, DB = DOP8 / 8, KB1 = DOP8 / 4;

    for (int i = tid; i < (int)C::n_tab; i += NT)
        s_tab[i] = exp2((double)(i / ExpDom<double>::REP) * (1.0 / ExpDom<double>::ENT));
    for (int i = tid; i < 64 * PCW; i += NT) {
        const int d = i / PCW, p = i - d * PCW;
        s_bs[d * LDP + p] = d < Do ? a.bs[(long)d * PP + pbase + p] : 0.0;
    }
    for (int i = tid; i < PCW * 16; i += NT) {
        const int p = i >> 4, sc = i & 15;
        double v = 0.0;
        if (sc < Q) v = a.zh[(long)sc * PP + pbase + p];
        else if (sc < 2 * Q) { const double zz = a.zh[(long)(sc - Q) * PP + pbase + p]; v = zz * zz; }
        s_Z[p * LDZ + sc] = v;
    }
    if (tid < 8) s_l2[tid] = tid < a.Qa ? exp(2.0 * a.ls[tid]) : 1.0;
    // SIMT phase: this thread's pair column and its rows ra, ra + RSTEP, ...
    const int pa = tid & (PCW - 1), ra = tid >> 6;
    double zh[Q], zh2[Q];
    GPB_UNROLL
    for (int q = 0; q < Q; q++) {
        zh[q] = a.zh[(long)q * PP + pbase + pa];
        zh2[q] = zh[q] * zh[q];
    }
    double accB[NI][2], accW[QB][2], accW2[QB][2];
    GPB_UNROLL
    for (int i = 0; i < NI; i++) accB[i][0] = accB[i][1] = 0.0;
    GPB_UNROLL
    for (int j = 0; j < QB; j++) accW[j][0] = accW[j][1] = accW2[j][0] = accW2[j][1] = 0.0;
    const int cbw = warp & 7, dpar = warp >> 3;              // G2 / G3: this warp's pair block
    const bool do_g3 = NW == 8 || warp >= 8, do_g4 = NW == 8 || warp < 8;

    const int r_begin = blockIdx.y * a.rows_per_split;
    const int r_end = (r_begin + a.rows_per_split) < a.n ? (r_begin + a.rows_per_split) : a.n;

    // Stage one row tile in two halves so that the global-memory latency hides under the tensor
    // phases: stage_load issues the loads of the NEXT tile (mu, vx of one (row, q) and DVT dv values
    // per thread) into registers, stage_store turns them into shared-memory operands two phases
    // later.  8 lanes per row (lane q of the group owns input dim q), the sums over q by shuffles.
    // rc = expanded-form constants of the exponent in the scaled domain:
    //   [0] = kS (lcn - sum c2 mu^2), [1+q] = 2 kS c2 mu, [1+Q+q] = -kS c2      (kernels.py:188-190)
    double g_mu = 0.0, g_vx = 0.0, g_dv[DVT];
    auto stage_load = [&](int t0) {
        const int tv = (r_end - t0) < TR ? (r_end - t0) : TR;
        const int row = tid >> 3, q = tid & 7;
        const bool ok = row < tv && q < a.Qa;      // (row >= TR for the upper half of a 16-warp CTA)
        g_mu = ok ? a.mx[(long)(t0 + row) * a.Qa + q] : 0.0;
        g_vx = ok ? a.vx[(long)(t0 + row) * a.Qa + q] : 0.0;
        GPB_UNROLL
        for (int i = 0; i < DVT; i++) {
            const int idx = tid + i * NT, r = idx >> 6, d = idx & 63;
            g_dv[i] = (r < tv && d < Do) ? a.dv[(long)(t0 + r) * Do + d] : 0.0;
        }
    };
    auto stage_store = [&](int buf, int t0) {
        double* s_dv = s_stage + (size_t)buf * C::n_stage;
        double* s_R = s_dv + C::n_dv;
        double* s_rc = s_R + C::n_R;
        const int tv = (r_end - t0) < TR ? (r_end - t0) : TR;
        if (tid < TR * 8) {      // whole warps (TR * 8 = 256)
            const int row = tid >> 3, q = tid & 7;
            const bool ok = row < tv;
            double c2 = 0.0, pr = 1.0;
            if (ok && q < a.Qa) {
                const double lq = s_l2[q];
                c2 = 1.0 / (2.0 * g_vx + lq);
                pr = lq * c2;
            }
            const double mu = g_mu, c2s = c2 * kS;
            double a0 = -c2s * mu * mu;
            GPB_UNROLL
            for (int m = 1; m < 8; m <<= 1) {
                a0 += shfl_xor(a0, m);
                pr *= shfl_xor(pr, m);
            }
            if (q < Q) {
                s_rc[row * RL + 1 + q] = 2.0 * c2s * mu;
                s_rc[row * RL + 1 + Q + q] = -c2s;
                s_R[row * LDZ + q] = c2 * mu;
                s_R[row * LDZ + Q + q] = c2;
            }
            if (2 * Q + q < 16) s_R[row * LDZ + 2 * Q + q] = 0.0;
            if (2 * Q + 8 + q < 16) s_R[row * LDZ + 2 * Q + 8 + q] = 0.0;
            // rows past the end: psi2' ~ 0 and their dv is 0
            if (q == 0) s_rc[row * RL] = (ok ? 0.5 * log(pr) : -1.0e5) * kS + a0;
        }
        GPB_UNROLL
        for (int i = 0; i < DVT; i++) {
            const int idx = tid + i * NT, r = idx >> 6, d = idx & 63;
            s_dv[r * LDV + d] = g_dv[i];
        }
    };

    sync_threads();
    if (r_begin < r_end) {
        stage_load(r_begin);
        stage_store(0, r_begin);
    }
    sync_threads();
    int buf = 0;
    for (int t0 = r_begin; t0 < r_end; t0 += TR, buf ^= 1) {
        const int tv = (r_end - t0) < TR ? (r_end - t0) : TR;
        const double* s_dv = s_stage + (size_t)buf * C::n_stage;
        const double* s_R = s_dv + C::n_dv;
        const double* s_rc = s_R + C::n_R;
        const bool more = t0 + TR < r_end;
        if (more) stage_load(t0 + TR);      // loads in flight during the SIMT phase and G1
        // ---- SIMT phase: psi2'[r, p] ---------------------------------------------------------
        {
            double x[RPT];
            GPB_UNROLL
            for (int i = 0; i < RPT; i++) {
                const double* rc = s_rc + (ra + RSTEP * i) * RL;
                double xx = rc[0];
                GPB_UNROLL
                for (int q = 0; q < Q; q++) {
                    xx += rc[1 + q] * zh[q];
                    xx += rc[1 + Q + q] * zh2[q];
                }
                x[i] = xx;
            }
            exp_dom_n<RPT>(x, s_tab, lane16);
            GPB_UNROLL
            for (int i = 0; i < RPT; i++) s_psi[(ra + RSTEP * i) * LDP + pa] = x[i];
        }
        sync_threads();
        // ---- G1: Lam = (dv . bs) * psi2' -----------------------------------------------------
        {
            const int rb = warp / (8 / NJ), cb0 = (warp % (8 / NJ)) * NJ;
            double c[NJ][2];
            GPB_UNROLL
            for (int j = 0; j < NJ; j++) c[j][0] = c[j][1] = 0.0;
            for (int k = 0; k < KB1; k++) {
                const double av = s_dv[(8 * rb + g) * LDV + 4 * k + t];
                GPB_UNROLL
                for (int j = 0; j < NJ; j++)
                    dmma(c[j][0], c[j][1], av, s_bs[(4 * k + t) * LDP + 8 * (cb0 + j) + g]);
            }
            GPB_UNROLL
            for (int j = 0; j < NJ; j++) {
                const int idx = (8 * rb + g) * LDP + 8 * (cb0 + j) + 2 * t;
                s_lam[idx] = c[j][0] * s_psi[idx];
                s_lam[idx + 1] = c[j][1] * s_psi[idx + 1];
            }
        }
        sync_threads();
        // next tile's constants / dv go to the other buffer while the products below run
        if (more) stage_store(buf ^ 1, t0 + TR);
        // ---- G2: dBp[d, p] += dv^T psi2' (pair columns 8 cbw .. +7, d blocks dpar, dpar + DSTEP, ..) ----
        GPB_UNROLL
        for (int k = 0; k < TR / 4; k++) {
            const double bv = s_psi[(4 * k + t) * LDP + 8 * cbw + g];
            GPB_UNROLL
            for (int i = 0; i < NI; i++)
                if (dpar + DSTEP * i < DB)
                    dmma(accB[i][0], accB[i][1], s_dv[(4 * k + t) * LDV + 8 * (dpar + DSTEP * i) + g], bv);
        }
        // ---- G3: W[p, s] += Lam^T R (even / odd k in separate accumulators: shorter chains) ----
        if (do_g3) {
            GPB_UNROLL
            for (int k = 0; k < TR / 4; k += 2) {
                const double av0 = s_lam[(4 * k + t) * LDP + 8 * cbw + g];
                const double av1 = s_lam[(4 * k + 4 + t) * LDP + 8 * cbw + g];
                GPB_UNROLL
                for (int j = 0; j < QB; j++) {
                    dmma(accW[j][0], accW[j][1], av0, s_R[(4 * k + t) * LDZ + 8 * j + g]);
                    dmma(accW2[j][0], accW2[j][1], av1, s_R[(4 * k + 4 + t) * LDZ + 8 * j + g]);
                }
            }
        }
        // ---- G4: V[r, s] = Lam Z -> row sums ---------------------------------------------------
        if (do_g4) {
            const int rb = cbw >> 1, cb = cbw & 1;
            if (cb < QB) {
                double vv[4][2];     // four independent chains over k
                GPB_UNROLL
                for (int c4 = 0; c4 < 4; c4++) vv[c4][0] = vv[c4][1] = 0.0;
                GPB_UNROLL
                for (int k = 0; k < PCW / 4; k += 4) {
                    GPB_UNROLL
                    for (int c4 = 0; c4 < 4; c4++)
                        dmma(vv[c4][0], vv[c4][1], s_lam[(8 * rb + g) * LDP + 4 * (k + c4) + t],
                             s_Z[(4 * (k + c4) + t) * LDZ + 8 * cb + g]);
                }
                const double v0 = (vv[0][0] + vv[1][0]) + (vv[2][0] + vv[3][0]);
                const double v1 = (vv[0][1] + vv[1][1]) + (vv[2][1] + vv[3][1]);
                const int r = 8 * rb + g, sc = 8 * cb + 2 * t;
                if (r < tv) {
                    if (sc < 2 * Q) atomic_add(a.rowacc + (long)(t0 + r) * (2 * Q) + sc, v0);
                    if (sc + 1 < 2 * Q) atomic_add(a.rowacc + (long)(t0 + r) * (2 * Q) + sc + 1, v1);
                }
            }
        }
        sync_threads();     // psi / lam free, next tile staged
    }
    // ---- pair sums of this (chunk, row split): {dBp_d, S0, S1_q} ------------------------------
    double* U = s_psi;      // [DOP8][LDP], aliases psi | lam
    double* WS = s_Z;       // [PCW][LDZ]
    GPB_UNROLL
    for (int i = 0; i < NI; i++)
        if (dpar + DSTEP * i < DB) {
            const int d = 8 * (dpar + DSTEP * i) + g;
            U[d * LDP + 8 * cbw + 2 * t] = accB[i][0];
            U[d * LDP + 8 * cbw + 2 * t + 1] = accB[i][1];
        }
    if (do_g3) {
        GPB_UNROLL
        for (int j = 0; j < QB; j++) {
            WS[(8 * cbw + g) * LDZ + 8 * j + 2 * t] = accW[j][0] + accW2[j][0];
            WS[(8 * cbw + g) * LDZ + 8 * j + 2 * t + 1] = accW[j][1] + accW2[j][1];
        }
    }
    sync_threads();
    double* rec = a.pairpart + (long)blockIdx.y * (Do + 1 + Q) * PP;
    for (int i = tid; i < Do * PCW; i += NT) {
        const int d = i / PCW, p = i - d * PCW;
        rec[(long)d * PP + pbase + p] = a.ep[pbase + p] * U[d * LDP + p];
    }
    if (tid < PCW) {
        double s0 = 0.0;
        for (int d = 0; d < Do; d++) s0 += s_bs[d * LDP + tid] * U[d * LDP + tid];
        rec[(long)Do * PP + pbase + tid] = s0;
        for (int q = 0; q < Q; q++)
            rec[(long)(Do + 1 + q) * PP + pbase + tid] =
                WS[tid * LDZ + q] - a.zh[(long)q * PP + pbase + tid] * WS[tid * LDZ + Q + q];
    }
}

// Forward contraction for WIDE output layers (Do > 4, fp64) on the FP64 tensor cores:
//   vacc[r, d] = sum_p bs[d, p] psi2'[r, p]                                  (aep_models.py:196-198)
// A CTA owns 64 rows at a time and sweeps ALL pair chunks: per chunk of 64 pairs the SIMT phase
// evaluates psi2'[64 x 64] into shared memory (one pair column and 16 rows per thread) and the
// product runs transposed as DMMA tiles, V^T[d, r] += bs[d, p] psi2'[r, p]^T, so that both
// operands are read with the conflict-free fragment pattern; warp w keeps the 8 rows 8w..8w+7 for
// every output block (<= 16 accumulators per thread).  The sums leave the CTA complete: plain
// stores, no atomics, no per-chunk partials (the SIMT row-owner kernel needed 255 registers and was
// shared-memory bound on the per-pair weight broadcasts).
template <int Q>
struct MMFwdWideMma {
    static constexpr int PCW = 64, TRF = 64, LDP = PCW + 4;
    static constexpr int RL = (1 + 2 * Q + 1) / 2 * 2;
    static constexpr size_t n_tab = ExpDom<double>::TAB;
    static constexpr size_t n_bs = 64 * LDP, n_psi = TRF * LDP, n_rc = TRF * RL;
    static constexpr size_t smem_bytes = 8 * (n_tab + n_bs + n_psi + n_rc);
    static_assert(Q <= 8, "8 staging lanes per row");
};

template <int Q>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) mm_fwd_wide_mma_kernel(MMArgs<double> a, int DOP8, int nchunks) {
    typedef MMFwdWideMma<Q> C;
    constexpr int PCW = C::PCW, TRF = C::TRF, LDP = C::LDP, RL = C::RL;
    constexpr double kS = ExpDom<double>::S;
    GPB_DYN_SMEM(dsm);
    double* s_tab = (double*)dsm;
    double* s_bs = s_tab + C::n_tab;     // [DOP8][LDP]   bs[d, p] of the current chunk
    double* s_psi = s_bs + C::n_bs;      // [TRF][LDP]    psi2'[r, p]
    double* s_rc = s_psi + C::n_psi;     // [TRF][RL]     expanded-form row constants
    GPB_SHARED double s_l2[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int lane16 = lane & (ExpDom<double>::REP - 1);
    const long PP = a.PP;
    const int Do = a.Do, DB = DOP8 / 8;
    const int pa = tid & (PCW - 1), ra = tid >> 6;

    for (int i = tid; i < (int)C::n_tab; i += kThreads)
        s_tab[i] = exp2((double)(i / ExpDom<double>::REP) * (1.0 / ExpDom<double>::ENT));
    if (tid < 8) s_l2[tid] = tid < a.Qa ? exp(2.0 * a.ls[tid]) : 1.0;
    sync_threads();

    const int nrb = (a.n + TRF - 1) / TRF;
    for (int rbk = blockIdx.x; rbk < nrb; rbk += gridDim.x) {
        const int t0 = rbk * TRF;
        const int tv = (a.n - t0) < TRF ? (a.n - t0) : TRF;
        // row constants: 8 lanes per row, two rows per 8-lane group (kernels.py:188-190)
        GPB_UNROLL
        for (int h = 0; h < TRF / 32; h++) {
            const int row = (tid >> 3) + 32 * h, q = tid & 7;
            const bool ok = row < tv;
            double mu = 0.0, c2 = 0.0, pr = 1.0;
            if (ok && q < a.Qa) {
                mu = a.mx[(long)(t0 + row) * a.Qa + q];
                const double lq = s_l2[q];
                c2 = 1.0 / (2.0 * a.vx[(long)(t0 + row) * a.Qa + q] + lq);
                pr = lq * c2;
            }
            const double c2s = c2 * kS;
            double a0 = -c2s * mu * mu;
            GPB_UNROLL
            for (int m = 1; m < 8; m <<= 1) {
                a0 += shfl_xor(a0, m);
                pr *= shfl_xor(pr, m);
            }
            if (q < Q) {
                s_rc[row * RL + 1 + q] = 2.0 * c2s * mu;
                s_rc[row * RL + 1 + Q + q] = -c2s;
            }
            if (q == 0) s_rc[row * RL] = (ok ? 0.5 * log(pr) : -1.0e5) * kS + a0;
        }
        double acc[8][2];
        GPB_UNROLL
        for (int i = 0; i < 8; i++) acc[i][0] = acc[i][1] = 0.0;
        sync_threads();
        // the weights / pair centres of chunk ch+1 are fetched into registers while chunk ch is
        // processed (global latency hidden under the SIMT and tensor phases)
        double g_bs[16], g_zh[Q];
        auto fetch = [&](int ch) {
            const long pb = (long)ch * PCW;
            GPB_UNROLL
            for (int i = 0; i < 16; i++) {
                const int idx = tid + i * kThreads, d = idx >> 6, p = idx & 63;
                g_bs[i] = d < Do ? a.bs[(long)d * PP + pb + p] : 0.0;
            }
            GPB_UNROLL
            for (int q = 0; q < Q; q++) g_zh[q] = a.zh[(long)q * PP + pb + pa];
        };
        fetch(0);
        for (int ch = 0; ch < nchunks; ch++) {
            // this chunk's weights -> shared memory
            GPB_UNROLL
            for (int i = 0; i < 16; i++) {
                const int idx = tid + i * kThreads, d = idx >> 6, p = idx & 63;
                if (d < DOP8) s_bs[d * LDP + p] = g_bs[i];
            }
            // SIMT phase: psi2' of this thread's pair column for its 16 rows, 8 at a time
            double zh[Q], zh2[Q];
            GPB_UNROLL
            for (int q = 0; q < Q; q++) {
                zh[q] = g_zh[q];
                zh2[q] = zh[q] * zh[q];
            }
            if (ch + 1 < nchunks) fetch(ch + 1);
            GPB_UNROLL
            for (int h = 0; h < 2; h++) {
                double x[8];
                GPB_UNROLL
                for (int i = 0; i < 8; i++) {
                    const double* rc = s_rc + (ra + 4 * (i + 8 * h)) * RL;
                    double xx = rc[0];
                    GPB_UNROLL
                    for (int q = 0; q < Q; q++) {
                        xx += rc[1 + q] * zh[q];
                        xx += rc[1 + Q + q] * zh2[q];
                    }
                    x[i] = xx;
                }
                exp_dom_n<8>(x, s_tab, lane16);
                GPB_UNROLL
                for (int i = 0; i < 8; i++) s_psi[(ra + 4 * (i + 8 * h)) * LDP + pa] = x[i];
            }
            sync_threads();
            // V^T[d, r] += bs[d, p] psi2'[r, p]: b = psi2'[r = 8 warp + g][p = 4k + t]
            GPB_UNROLL_N(2)
            for (int k = 0; k < PCW / 4; k++) {
                const double bv = s_psi[(8 * warp + g) * LDP + 4 * k + t];
                GPB_UNROLL
                for (int i = 0; i < 8; i++)
                    if (i < DB) dmma(acc[i][0], acc[i][1], s_bs[(8 * i + g) * LDP + 4 * k + t], bv);
            }
            sync_threads();     // s_bs / s_psi free
        }
        // c0 = V^T[d = 8i + g][r = 8 warp + 2t], c1: r + 1
        GPB_UNROLL
        for (int i = 0; i < 8; i++)
            if (i < DB) {
                const int d = 8 * i + g, r = 8 * warp + 2 * t;
                if (d < Do) {
                    if (r < tv) a.rowacc[(long)(t0 + r) * Do + d] = acc[i][0];
                    if (r + 1 < tv) a.rowacc[(long)(t0 + r + 1) * Do + d] = acc[i][1];
                }
            }
    }
}

template <typename T> struct Psi1Dom;   // exponent scale folded into c1 (argument of exp is -e)
template <> struct Psi1Dom<double> { static constexpr double kH = 0.5 * 64.0 / 0.693147180559945309417232; };
template <> struct Psi1Dom<float> { static constexpr double kH = 0.5 * 1.4426950408889634074; };
GPB_DEVICE double psi1_exp(double e, const double* tab, int lane16) { return exp_dom64(-e, tab, lane16); }
GPB_DEVICE double psi1_exp(float e, const double*, int) {
#ifndef GPB_CPU_EMU
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(-e));
    return (double)y;
#else
    return (double)exp2f(-e);
#endif
}

template <typename T, int QT>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) mm_psi1_fwd_kernel(
    const double* __restrict__ mx, const double* __restrict__ vx, const double* __restrict__ z,
    const double* __restrict__ ls, const double* __restrict__ sf, const double* __restrict__ A,
    const double* __restrict__ vacc, int n, int M, int Q, int Do, double* __restrict__ mout,
    double* __restrict__ vout, double* __restrict__ psi1save) {
    constexpr int TR = 32, MAXC = 2;              // M <= 512: at most 2 columns per thread
    GPB_DYN_SMEM(smem);
    const int LDT = M + 1;
    double* tile = (double*)smem;                 // [TR][LDT]
    double* s_cn = tile + (long)TR * LDT;         // [TR]   sf2 * prod sqrt(l^2 c1)
    double* tab = s_cn + TR;                      // [64*16]
    double* s_l2 = tab + 1024;                    // [QT]
    T* s_mu = (T*)(s_l2 + QT);                    // [TR][QT]
    T* s_c1 = s_mu + TR * QT;                     // [TR][QT]  c1 * kH
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lane16 = tid & 15;
    const double sf2 = exp(2.0 * sf[0]);
    for (int i = tid; i < 1024; i += kThreads) tab[i] = exp2((double)(i >> 4) * (1.0 / 64.0));
    if (tid < QT) s_l2[tid] = tid < Q ? exp(2.0 * ls[tid]) : 1.0;
    T zr[MAXC][QT];
    GPB_UNROLL
    for (int c = 0; c < MAXC; c++) {
        const int m = tid + c * kThreads;
        GPB_UNROLL
        for (int q = 0; q < QT; q++) zr[c][q] = (m < M && q < Q) ? (T)z[(long)m * Q + q] : (T)0;
    }
    const int ntiles = (n + TR - 1) / TR;
    for (int tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
        const int t0 = tile_i * TR;
        const int tv = (n - t0) < TR ? (n - t0) : TR;
        sync_threads();                            // previous tile consumed; tab / s_l2 visible
        for (int i = tid; i < TR * QT; i += kThreads) {
            const int r = i / QT, q = i - r * QT;
            double mu = 0, cc = 0;
            if (r < tv && q < Q) {
                mu = mx[(long)(t0 + r) * Q + q];
                cc = 1.0 / (vx[(long)(t0 + r) * Q + q] + s_l2[q]);
            }
            s_mu[i] = (T)mu;
            s_c1[i] = (T)(cc * Psi1Dom<T>::kH);
        }
        if (tid < TR) {
            double cn = sf2;
            if (tid < tv)
                for (int q = 0; q < Q; q++) cn *= sqrt(s_l2[q] / (vx[(long)(t0 + tid) * Q + q] + s_l2[q]));
            s_cn[tid] = tid < tv ? cn : 0.0;
        }
        sync_threads();
        GPB_UNROLL
        for (int c = 0; c < MAXC; c++) {
            const int m = tid + c * kThreads;
            if (m < M) {
                GPB_UNROLL_N(4)
                for (int r = 0; r < TR; r++) {
                    T e = 0;
                    GPB_UNROLL
                    for (int q = 0; q < QT; q++) {
                        const T diff = s_mu[r * QT + q] - zr[c][q];
                        e += diff * diff * s_c1[r * QT + q];
                    }
                    const double p = s_cn[r] * psi1_exp(e, tab, lane16);
                    tile[(long)r * LDT + m] = p;
                    if (psi1save != nullptr && r < tv) psi1save[(long)(t0 + r) * M + m] = p;
                }
            }
        }
        sync_threads();
        for (int rr = 0; rr < 4; rr++) {
            const int r = warp * 4 + rr;
            if (r >= tv) continue;                 // warp-uniform
            for (int d = 0; d < Do; d++) {
                double acc = 0;
                for (int m = lane; m < M; m += 32) acc += ldg(A + (long)d * M + m) * tile[(long)r * LDT + m];
                acc = warp_sum(acc);
                if (lane == 0) {
                    mout[(long)(t0 + r) * Do + d] = acc;
                    vout[(long)(t0 + r) * Do + d] = sf2 + vacc[(long)(t0 + r) * Do + d] - acc * acc;
                }
            }
        }
    }
}

// Backward, row-wise part: finishes dmx, dvx per row from the psi1 terms
// (kernels.py:355-378) and the reduced psi2 sums {s = dv . vacc, U, V} (kernels.py:419-431),
// and emits per-block partials of the row-summed hyper terms:
//   part[blk][0]      : dsf2  = sum_n (sum_m L1 + 2 s_n)/sf2
//   part[blk][1+q]    : dl_q  (psi1: Zmu2_denom.. , psi2: the n-dependent terms of kernels.py:441-442)
//   part[blk][1+Q]    : sum of scaled dv  (dv_sum, aep_models.py:247)
// One warp per row: lanes stride over the pseudo-points and stream the saved psi1 row
// (coalesced); z^T and A sit in shared memory; the 1+2Q row sums are warp-reduced and lane q
// finishes input dimension q.
template <int QT>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) mm_rows_bwd_kernel(
    const double* __restrict__ mx, const double* __restrict__ vx, const double* __restrict__ z,
    const double* __restrict__ ls, const double* __restrict__ sf, const double* __restrict__ A,
    const double* __restrict__ dm, const double* __restrict__ dv, const double* __restrict__ mout,
    const double* __restrict__ vacc, const double* __restrict__ rowacc,
    const double* __restrict__ psi1, int n, int M, int Q, int Do, double* __restrict__ dmx,
    double* __restrict__ dvx, double* __restrict__ part) {
    GPB_DYN_SMEM(smem);
    double* zsT = (double*)smem;                  // [QT][M]
    double* As = zsT + (long)QT * M;              // [Do][M]
    double* s_dma = As + (long)Do * M;            // [8][Do]
    double* s_red = s_dma + 8 * Do;               // [8][QT + 2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < QT * M; i += kThreads) {
        const int q = i / M, m = i - q * M;
        zsT[i] = q < Q ? z[(long)m * Q + q] : 0.0;
    }
    for (int i = tid; i < Do * M; i += kThreads) As[i] = A[i];
    const double sf2 = exp(2.0 * sf[0]);
    sync_threads();
    const int NS = 2 * QT;
    double p_sf2 = 0, p_dvsum = 0, p_dl = 0;      // lane 0 / lane 0 / lane q
    double* dma = s_dma + warp * Do;
    for (long row = (long)blockIdx.x * 8 + warp; row < n; row += (long)gridDim.x * 8) {
        double mu[QT], c1[QT], dmu[QT], dS[QT];
        GPB_UNROLL
        for (int q = 0; q < QT; q++) {
            mu[q] = q < Q ? mx[row * Q + q] : 0.0;
            c1[q] = q < Q ? 1.0 / (vx[row * Q + q] + exp(2.0 * ls[q])) : 0.0;
            dmu[q] = 0;
            dS[q] = 0;
        }
        double s = 0, dvs = 0;
        sync_warp();                               // previous row's dma fully consumed
        for (int d = lane; d < Do; d += 32) {
            const double dvd = dv[row * Do + d];
            dma[d] = dm[row * Do + d] - 2.0 * dvd * mout[row * Do + d];
            dvs += dvd;
            s += dvd * vacc[row * Do + d];
        }
        sync_warp();
        s = warp_sum(s);                           // sum_p Lam = sum_d dv_d * forward vacc_d
        dvs = warp_sum(dvs);
        double L1sum = 0;
        for (int m = lane; m < M; m += 32) {
            const double p1 = psi1[row * M + m];
            double g = 0;
            for (int d = 0; d < Do; d++) g += dma[d] * As[(long)d * M + m];
            const double L1 = g * p1;
            L1sum += L1;
            GPB_UNROLL
            for (int q = 0; q < QT; q++) {
                const double zm = zsT[(long)q * M + m] - mu[q];
                const double w = zm * c1[q];
                dmu[q] += L1 * w;
                dS[q] += L1 * (zm * w - 1.0) * c1[q];   // x 1/2 below
            }
        }
        L1sum = warp_sum(L1sum);
        double my_dmu = 0, my_dS = 0, my_mu = 0, my_c1 = 0;
        GPB_UNROLL
        for (int q = 0; q < QT; q++) {
            const double a = warp_sum(dmu[q]), b = warp_sum(dS[q]);
            if (q == lane) { my_dmu = a; my_dS = b; my_mu = mu[q]; my_c1 = c1[q]; }
        }
        if (lane == 0) {
            p_sf2 += (L1sum + 2.0 * s) / sf2;
            p_dvsum += dvs;
        }
        if (lane < Q) {
            const int q = lane;
            const double l = exp(ls[q]), lq = l * l;
            const double S = vx[row * Q + q];
            const double c2 = 1.0 / (2.0 * S + lq);
            const double U = rowacc[row * NS + q], V = rowacc[row * NS + QT + q];
            // psi1: sum_m L1 ((z-mu)^2 c1 + S/l^2) c1 l  =  (dS_acc + (1 + S/l^2) L1sum c1) l
            const double dl1 = (my_dS + (1.0 + S / lq) * my_c1 * L1sum) * l;
            // psi2 (n-dependent part of kernels.py:441-442)
            const double dl2 = 2.0 * l * ((S / lq * c2 + my_mu * my_mu * c2 * c2) * s -
                                          2.0 * my_mu * c2 * c2 * U + c2 * c2 * V);
            p_dl += dl1 + dl2;
            dmx[row * Q + q] = my_dmu - 2.0 * c2 * (my_mu * s - U);
            dvx[row * Q + q] = 0.5 * my_dS + 2.0 * c2 * c2 * (my_mu * my_mu * s - 2.0 * my_mu * U + V) - c2 * s;
        }
    }
    // per-block partials: sum the 8 warps
    double* rr = s_red + warp * (QT + 2);
    if (lane == 0) { rr[0] = p_sf2; rr[1] = p_dvsum; }
    if (lane < QT) rr[2 + lane] = p_dl;
    sync_threads();
    if (tid < Q + 2) {
        const int k = tid == 0 ? 0 : (tid == Q + 1 ? 1 : 2 + (tid - 1));   // -> [sf2 | dl_q | dvsum]
        double acc = 0;
        for (int w = 0; w < 8; w++) acc += s_red[w * (QT + 2) + k];
        part[(long)blockIdx.x * (2 + Q) + tid] = acc;
    }
}

// Backward, column-wise psi1 part (kernels.py:355-378 terms indexed by m; aep_models.py:239):
//   dA[d,m] = sum_n dm_all[n,d] psi1[n,m] ;  dZ1[m,q] = -sum_n L1 (z_mq - mu_nq) c1_nq,
//   L1[n,m] = (sum_d dm_all[n,d] A[d,m]) psi1[n,m]
// Thread per pseudo-point column (z, A, and the accumulators in registers), the rows of this
// block's range staged 32 at a time in shared memory, psi1 streamed (coalesced, U rows in
// flight) from the forward's save buffer.  DOB output dims per pass (Do <= 4: one pass); the
// first pass of a wider layer forms g over all Do with a runtime loop.
// partial record per block: [ dA (Do*M) | dZ1 (M*Q) ]
template <int QT, int DOB>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) mm_cols_bwd_kernel(
    const double* __restrict__ mx, const double* __restrict__ vx, const double* __restrict__ z,
    const double* __restrict__ ls, const double* __restrict__ A, const double* __restrict__ dm,
    const double* __restrict__ dv, const double* __restrict__ mout, const double* __restrict__ psi1,
    int n, int M, int Q, int Do, int d0, int rows_per_block, double* __restrict__ part) {
    constexpr int TR = 32, U = 8;
    GPB_DYN_SMEM(smem);
    double* s_mu = (double*)smem;          // [TR][QT]
    double* s_c1 = s_mu + TR * QT;          // [TR][QT]
    double* s_dma = s_c1 + TR * QT;         // [TR][Do]   dm_all (all output dims)
    const int tid = threadIdx.x;
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = (r_begin + rows_per_block) < n ? (r_begin + rows_per_block) : n;
    double* rec = part + (long)blockIdx.x * ((long)Do * M + (long)M * Q);
    const bool wide = Do > DOB;             // g needs output dims outside this pass' registers
    for (int m0 = 0; m0 < M; m0 += kThreads) {
        const int m = m0 + tid;
        const bool act = m < M;
        double zr[QT], dz[QT], Ar[DOB], dA[DOB];
        GPB_UNROLL
        for (int q = 0; q < QT; q++) {
            zr[q] = (act && q < Q) ? z[(long)m * Q + q] : 0.0;
            dz[q] = 0;
        }
        GPB_UNROLL
        for (int d = 0; d < DOB; d++) {
            Ar[d] = (act && d0 + d < Do) ? A[(long)(d0 + d) * M + m] : 0.0;
            dA[d] = 0;
        }
        for (int t0 = r_begin; t0 < r_end; t0 += TR) {
            const int tv = (r_end - t0) < TR ? (r_end - t0) : TR;
            sync_threads();
            for (int i = tid; i < TR * QT; i += kThreads) {
                const int r = i / QT, q = i - r * QT;
                const bool ok = r < tv && q < Q;
                s_mu[i] = ok ? mx[(long)(t0 + r) * Q + q] : 0.0;
                s_c1[i] = ok ? 1.0 / (vx[(long)(t0 + r) * Q + q] + exp(2.0 * ls[q])) : 0.0;
            }
            for (int i = tid; i < TR * Do; i += kThreads) {
                const int r = i / Do, d = i - r * Do;
                s_dma[i] = r < tv ? dm[(long)(t0 + r) * Do + d] -
                                        2.0 * dv[(long)(t0 + r) * Do + d] * mout[(long)(t0 + r) * Do + d]
                                  : 0.0;
            }
            sync_threads();
            if (act)
                for (int rb = 0; rb < tv; rb += U) {
                    double p1[U];
                    GPB_UNROLL
                    for (int u = 0; u < U; u++)
                        p1[u] = (rb + u) < tv ? psi1[(long)(t0 + rb + u) * M + m] : 0.0;
                    GPB_UNROLL
                    for (int u = 0; u < U; u++) {
                        const int r = (rb + u) < tv ? rb + u : rb;   // clamped rows carry p1 = 0
                        double g = 0;
                        GPB_UNROLL
                        for (int d = 0; d < DOB; d++) {
                            const double w = (d0 + d < Do) ? s_dma[r * Do + d0 + d] : 0.0;
                            g += w * Ar[d];
                            dA[d] += w * p1[u];
                        }
                        if (wide) {
                            g = 0;
                            if (d0 == 0)
                                for (int d = 0; d < Do; d++) g += s_dma[r * Do + d] * ldg(A + (long)d * M + m);
                        }
                        const double L1 = g * p1[u];
                        GPB_UNROLL
                        for (int q = 0; q < QT; q++) dz[q] -= L1 * (zr[q] - s_mu[r * QT + q]) * s_c1[r * QT + q];
                    }
                }
        }
        if (act) {
            GPB_UNROLL
            for (int d = 0; d < DOB; d++)
                if (d0 + d < Do) rec[(long)(d0 + d) * M + m] = dA[d];
            if (d0 == 0) {
                GPB_UNROLL
                for (int q = 0; q < QT; q++)
                    if (q < Q) rec[(long)Do * M + (long)m * Q + q] = dz[q];
            }
        }
    }
}

// Backward, pair -> matrix scatter (deterministic gather form).  Inputs are the pair sums
// already reduced over row splits: dBp[Do][PP], S0[PP], S1[Q][PP].
//   dB[d,a,b]   = dBp[d,p(a,b)]                                   (aep_models.py:240)
//   dZ2[a,q]    = sum_b { a!=b: S0/2 (z_b-z_a)/l^2 + S1_q ; a==b: 2 S1_q }   (kernels.py:436-440)
//   W_q (dl)    = sum_p S0[p] (z_a-z_b)^2/4  ->  dlW[a,q] partial rows, summed on the host side
GPB_KERNEL void mm_pair_finish_kernel(const double* __restrict__ pairsum, int DOCs, int Do,
                                      const double* __restrict__ z, const double* __restrict__ ls,
                                      int M, int Q, long PP, double* __restrict__ dB,
                                      double* __restrict__ dZ2, double* __restrict__ dlW) {
    // pairsum layout: [Do dBp rows][S0][S1 x Q], each of length PP
    const double* S0 = pairsum + (long)Do * PP;
    const double* S1 = pairsum + (long)(Do + 1) * PP;
    (void)DOCs;
    const long totalB = (long)Do * M * M;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < totalB;
         idx += (long)gridDim.x * blockDim.x) {
        int b = (int)(idx % M), a = (int)((idx / M) % M), d = (int)(idx / ((long)M * M));
        int hi = a > b ? a : b, lo = a > b ? b : a;
        dB[idx] = pairsum[(long)d * PP + (long)hi * (hi + 1) / 2 + lo];
    }
    // one warp per (a, q): the M-term sums over b run across the lanes (a thread per (a, q) walked them one
    // dependent global load at a time: 145 us at M = 256, a constant of the step that no row count amortises)
    const long totalZ = (long)M * Q;
    const int lane = threadIdx.x & 31;
    const long nwarp = (long)gridDim.x * (blockDim.x >> 5);
    for (long idx = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); idx < totalZ; idx += nwarp) {
        int q = (int)(idx % Q), a = (int)(idx / Q);
        const double il2 = exp(-2.0 * ls[q]);
        const double za = z[(long)a * Q + q];
        double g = 0, w = 0;
        for (int b = lane; b < M; b += 32) {
            int hi = a > b ? a : b, lo = a > b ? b : a;
            long p = (long)hi * (hi + 1) / 2 + lo;
            double s1 = S1[(long)q * PP + p];
            if (a == b) {
                g += 2.0 * s1;
            } else {
                double dzb = z[(long)b * Q + q] - za;
                g += 0.5 * S0[p] * dzb * il2 + s1;
                if (b < a) w += S0[p] * dzb * dzb * 0.25;   // each unordered pair once
            }
        }
        g = warp_sum(g);
        w = warp_sum(w);
        if (lane == 0) {
            dZ2[idx] = g;
            dlW[idx] = w;
        }
    }
}

// zero-padded typed operand copies for the deterministic layer kernels
template <typename T>
GPB_KERNEL void det_pad_kernel(const double* __restrict__ A, const double* __restrict__ B, int M,
                               int MP, int Do, T* __restrict__ Ap, T* __restrict__ Bp) {
    const long totalB = (long)Do * MP * MP;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < totalB;
         i += (long)gridDim.x * blockDim.x) {
        int j = (int)(i % MP), r = (int)((i / MP) % MP), d = (int)(i / ((long)MP * MP));
        Bp[i] = (r < M && j < M) ? (T)B[((long)d * M + r) * M + j] : (T)0;
        if (i < (long)Do * MP) {
            int c = (int)(i % MP), dd = (int)(i / MP);
            Ap[i] = c < M ? (T)A[(long)dd * M + c] : (T)0;
        }
    }
}

// det_bwd epilogue: fold the summed partial record [cs | dz | dl | dA] (padded MP columns)
// into dA[Do,M], dzu[M,D], dl[D], dsf2[1].  One block.
GPB_KERNEL void det_bwd_finish_kernel(const double* __restrict__ rec, const double* __restrict__ sf,
                                      int M, int MP, int D, int Do, double* __restrict__ dA,
                                      double* __restrict__ dzu, double* __restrict__ dl,
                                      double* __restrict__ dsf2) {
    GPB_SHARED double scratch[16];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < Do * M; i += nt) {
        int d = i / M, m = i - d * M;
        dA[i] = rec[(long)MP + 2L * MP * D + (long)d * MP + m];
    }
    for (int i = tid; i < M * D; i += nt) dzu[i] = rec[(long)MP + i];
    double s = 0;
    for (int m = tid; m < M; m += nt) s += rec[m];
    double r = block_sum(s, scratch);
    if (tid == 0) dsf2[0] = r / exp(2.0 * sf[0]);
    for (int q = 0; q < D; q++) {
        s = 0;
        for (int m = tid; m < M; m += nt) s += rec[(long)MP + (long)MP * D + (long)m * D + q];
        r = block_sum(s, scratch);
        if (tid == 0) dl[q] = r;
    }
}

// mm_bwd epilogue: dzu = dZ1 + dZ2 ; dl_q = dl_rows_q + 2 l_q W_q / l_q^4 (kernels.py:441-442);
// scalars copied out.  One block.
GPB_KERNEL void mm_final_kernel(const double* __restrict__ colsum /*[Do*M + M*Q]*/,
                                const double* __restrict__ rowsum /*[2+Q]*/,
                                const double* __restrict__ dZ2, const double* __restrict__ dlW,
                                const double* __restrict__ ls, int M, int Q, int Do,
                                double* __restrict__ dA, double* __restrict__ dzu,
                                double* __restrict__ dl, double* __restrict__ dsf2,
                                double* __restrict__ dvsum) {
    GPB_SHARED double scratch[16];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < Do * M; i += nt) dA[i] = colsum[i];
    for (int i = tid; i < M * Q; i += nt) dzu[i] = colsum[(long)Do * M + i] + dZ2[i];
    for (int q = 0; q < Q; q++) {
        double s = 0;
        for (int a = tid; a < M; a += nt) s += dlW[(long)a * Q + q];
        double r = block_sum(s, scratch);
        if (tid == 0) {
            double l = exp(ls[q]);
            dl[q] = rowsum[1 + q] + 2.0 * l * r / (l * l * l * l);
        }
    }
    if (tid == 0) {
        dsf2[0] = rowsum[0];
        dvsum[0] = rowsum[1 + Q];
    }
}

// =========================================================================
// a3 / a4 / a10 building block: batched inverse + log-determinant of SPD M x M matrices
// (Kuu, Kuu^-1 + theta_1, Kuu^-1 + beta theta_1: base_models.py:464,471,476, aep_models.py:68,78,91,
// 525,533 -- np.linalg.inv / slogdet in the reference).  fp64 throughout.
// One thread-block CLUSTER per matrix (kInvCluster CTAs, hardware cluster barrier); blocked
// Gauss-Jordan without pivoting (valid for SPD), NB = 32:
//   per block step k:   D = A_kk (Schur complement so far);  logdet += log det D
//     every CTA:   stage the row panel R = A[k,:] and its own column block A[rows,k] (cp.async),
//                  invert D in shared memory (unblocked GJ, 4 warps, reciprocal by Newton iteration)
//     CTA r, its rows i not in k:   P_i = A_ik D^-1;  A_ij -= P_i R_j (j not in k; FP64 tensor cores);  A_ik = -P_i
//     column slice r of the k rows: A_kj = D^-1 R_j (j not in k);  A_kk = D^-1
// The matrix stays in global memory (L2 resident: 0.5 MB at M = 256); a step reads the panel once into shared
// memory and applies the rank-NB update to the CTA's row slice with DMMA 8x8x4 (P: A fragments, R: B fragments;
// strides = 4 / 8 mod 16 doubles keep both fragment loads conflict free).  Two cluster barriers per step.
// After M/NB steps the work matrix holds A^-1.
#ifndef GPB_CPU_EMU
constexpr int kInvCluster = 8;
GPB_DEVICE double pivot_rcp(double x) {     // 1/x off the long IEEE-division path: the pivots form a serial chain
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
#else
constexpr int kInvCluster = 1;
static inline double pivot_rcp(double x) { return 1.0 / x; }
#endif

template <int NB>
struct SpdInvCfg {
    static constexpr int PS = NB + 4;                       // P / column-block row stride (= 4 mod 16)
    static int rows8(int M) { return ((M + kInvCluster - 1) / kInvCluster + 7) / 8 * 8; }
    static int rs(int M) { return (M + 15) / 16 * 16 + 8; } // panel row stride (= 8 mod 16)
    static size_t smem_bytes(int M) {
        return sizeof(double) * (2 * (size_t)NB * (NB + 1) + (size_t)NB * rs(M) + 2 * (size_t)rows8(M) * PS + 64);
    }
};

template <int NB>
GPB_KERNEL void GPB_CLUSTER(kInvCluster) GPB_LAUNCH_BOUNDS(512) spd_inverse_kernel(
    const double* __restrict__ A, int M, double* __restrict__ W /* [batch, M, M]: out = A^-1 */,
    double* __restrict__ logdet /* [batch] */) {
    GPB_DYN_SMEM(smem);
    constexpr int NT = 512, NW = NT / 32, PS = SpdInvCfg<NB>::PS, TB = 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int rank = cluster_rank();
    const int mat = blockIdx.x / kInvCluster;
    const int rows_per = (M + kInvCluster - 1) / kInvCluster;
    const int row_lo = rank * rows_per < M ? rank * rows_per : M;
    const int row_hi = (row_lo + rows_per) < M ? (row_lo + rows_per) : M;
    const int nrows = row_hi - row_lo;
    const int rows8 = (rows_per + 7) / 8 * 8, RS = (M + 15) / 16 * 16 + 8;   // = SpdInvCfg::rows8 / rs
    double* D = (double*)smem;                 // [NB][NB+1]
    double* D2 = D + NB * (NB + 1);            // [NB][NB+1]  ping-pong twin of D
    double* R = D2 + NB * (NB + 1);            // [NB][RS]    row panel (values before the step)
    double* P = R + (size_t)NB * RS;           // [rows8][PS] A_ik D^-1
    double* Ac = P + (size_t)rows8 * PS;       // [rows8][PS] A_ik (values before the step)
    double* s_ld = Ac + (size_t)rows8 * PS;    // [1]
    double* s_piv = s_ld + 8;                  // [NB]
    const double* Am = A + (size_t)mat * M * M;
    double* Wm = W + (size_t)mat * M * M;
    const bool even = (M & 1) == 0 && ((size_t)Wm & 15) == 0;   // rows 16-byte aligned: 16-byte asynchronous copies
    // work copy of this CTA's rows
    for (long i = tid; i < (long)nrows * M; i += NT) Wm[(long)row_lo * M + i] = Am[(long)row_lo * M + i];
    if (tid == 0) s_ld[0] = 0.0;
    cluster_sync();
    for (int k0 = 0; k0 < M; k0 += NB) {
        const int nb = (M - k0) < NB ? (M - k0) : NB;
        // ---- stage the row panel (rows past nb zero) and this CTA's column block ----
        if (even) {
            const int Mh = M >> 1;
            for (int i = tid; i < NB * Mh; i += NT) {
                const int c = i / Mh, jj = 2 * (i - c * Mh);
                cp_async16_zfill(&R[c * RS + jj], &Wm[(long)(k0 + (c < nb ? c : 0)) * M + jj], c < nb);
            }
        } else {
            for (int i = tid; i < NB * M; i += NT) {
                const int c = i / M, jj = i - c * M;
                cp_async8_zfill(&R[c * RS + jj], &Wm[(long)(k0 + (c < nb ? c : 0)) * M + jj], c < nb);
            }
        }
        for (int i = tid; i < rows8 * NB; i += NT) {
            const int r = i / NB, c = i - r * NB;
            const bool ok = r < nrows && c < nb;
            cp_async8_zfill(&Ac[r * PS + c], &Wm[ok ? (long)(row_lo + r) * M + k0 + c : 0], ok);
        }
        cp_async_commit();
        cp_async_wait<0>();
        sync_threads();
        // ---- warps 0-3 invert the diagonal block: lane = row, warp = group of NB/4 columns; ping-pong between
        //      D and D2 (a Gauss-Jordan step reads the old block and writes the new one), one 128-thread named
        //      barrier per pivot.  The chain of M pivots is the serial part of the whole inversion: the
        //      reciprocal is a Newton iteration and the logs of the pivots are taken afterwards. ----
        if (tid < 128) {
            constexpr int CG = NB / 4;
            const int a = tid & 31, cg0 = (tid >> 5) * CG;
            double* Dc = D;
            double* Dn = D2;
            for (int b = cg0; b < cg0 + CG; b++)
                Dc[a * (NB + 1) + b] = (a < nb && b < nb) ? R[a * RS + k0 + b] : (a == b ? 1.0 : 0.0);
            sync_group<1, 128>();
            for (int j = 0; j < nb; j++) {
                const double piv = Dc[j * (NB + 1) + j];
                const double aj = Dc[a * (NB + 1) + j];
                double jb[CG], ab[CG];
                GPB_UNROLL
                for (int bb = 0; bb < CG; bb++) {
                    jb[bb] = Dc[j * (NB + 1) + cg0 + bb];
                    ab[bb] = Dc[a * (NB + 1) + cg0 + bb];
                }
                const double ip = pivot_rcp(piv);
                const double f = (a == j) ? 0.0 : aj * ip;
                if (tid == j) s_piv[j] = piv;
                GPB_UNROLL
                for (int bb = 0; bb < CG; bb++) {
                    const int b = cg0 + bb;
                    const double v = (a == j) ? jb[bb] * ip : ab[bb] - f * jb[bb];
                    Dn[a * (NB + 1) + b] = (b == j) ? ((a == j) ? ip : -aj * ip) : v;
                }
                sync_group<1, 128>();
                double* tp = Dc; Dc = Dn; Dn = tp;
            }
            if (Dc != D) {                         // odd number of pivots: result sits in D2
                for (int b = cg0; b < cg0 + CG; b++) D[a * (NB + 1) + b] = Dc[a * (NB + 1) + b];
            }
            if (tid < 32) {
                double l = tid < nb ? log(s_piv[tid]) : 0.0;
                l = warp_sum(l);
                if (tid == 0) s_ld[0] += l;
            }
        }
        sync_threads();
        // ---- P = A[rows, k] D^-1 for this CTA's rows (zero rows: the k rows themselves and the padding) ----
        for (int i = tid; i < rows8 * NB; i += NT) {
            const int r = i / NB, c = i - r * NB;
            const int row = row_lo + r;
            double acc = 0;
            if (r < nrows && c < nb && (row < k0 || row >= k0 + nb)) {
                GPB_UNROLL_N(8)
                for (int b = 0; b < nb; b++) acc += Ac[r * PS + b] * D[b * (NB + 1) + c];
            }
            P[r * PS + c] = acc;
        }
        cluster_sync();            // every CTA holds R, D^-1, P: the k rows / k columns may now change
        // ---- trailing update of this CTA's rows on the FP64 tensor cores: 8 x 8 tiles, TB per warp at a time
        //      (their old values are fetched before the MMAs) ----
        const int tiles_r = rows8 / 8, tiles_c = (M + 7) / 8, ntile = tiles_r * tiles_c;
        for (int tile0 = warp; tile0 < ntile; tile0 += NW * TB) {
            double w0[TB], w1[TB], c0[TB], c1[TB];
            int rl[TB], col[TB];
            bool rok[TB];
            GPB_UNROLL
            for (int u = 0; u < TB; u++) {
                const int tile = tile0 + u * NW;
                const bool tv = tile < ntile;
                const int tr = tv ? tile % tiles_r : 0, tc = tv ? tile / tiles_r : 0;
                rl[u] = tr * 8 + g;
                col[u] = tc * 8 + 2 * t4;
                const int row = row_lo + rl[u];
                rok[u] = tv && rl[u] < nrows && (row < k0 || row >= k0 + nb);
                w0[u] = (rok[u] && col[u] < M) ? Wm[(long)row * M + col[u]] : 0.0;
                w1[u] = (rok[u] && col[u] + 1 < M) ? Wm[(long)row * M + col[u] + 1] : 0.0;
                c0[u] = 0.0;
                c1[u] = 0.0;
            }
            GPB_UNROLL
            for (int ks = 0; ks < NB / 4; ks++) {
                GPB_UNROLL
                for (int u = 0; u < TB; u++) {
                    const double pa = P[rl[u] * PS + 4 * ks + t4];
                    const double rb = R[(4 * ks + t4) * RS + (col[u] - 2 * t4) + g];
                    dmma(c0[u], c1[u], pa, rb);
                }
            }
            GPB_UNROLL
            for (int u = 0; u < TB; u++) {
                if (!rok[u]) continue;
                const long base = (long)(row_lo + rl[u]) * M;
                if (col[u] < M) {
                    const int cc = col[u];
                    Wm[base + cc] = (cc >= k0 && cc < k0 + nb) ? -P[rl[u] * PS + cc - k0] : w0[u] - c0[u];
                }
                if (col[u] + 1 < M) {
                    const int cc = col[u] + 1;
                    Wm[base + cc] = (cc >= k0 && cc < k0 + nb) ? -P[rl[u] * PS + cc - k0] : w1[u] - c1[u];
                }
            }
        }
        // ---- the k rows themselves: A_kj = D^-1 R_j, A_kk = D^-1.  Split by COLUMN slices over
        //      the cluster (not by row ownership), so that no CTA is a straggler at the barrier ----
        const int ncol = nrows;                    // column slice of this CTA = its row range
        for (int i = tid; i < nb * ncol; i += NT) {
            const int a = i / ncol, j = row_lo + (i - a * ncol);
            double v;
            if (j >= k0 && j < k0 + nb) v = D[a * (NB + 1) + (j - k0)];
            else {
                v = 0;
                GPB_UNROLL_N(8)
                for (int c = 0; c < nb; c++) v += D[a * (NB + 1) + c] * R[c * RS + j];
            }
            Wm[(long)(k0 + a) * M + j] = v;
        }
        cluster_sync();            // step complete everywhere before the next panel is staged
    }
    if (rank == 0 && tid == 0) logdet[mat] = s_ld[0];
}

// M <= 256: the same blocked Gauss-Jordan with the matrix held in REGISTERS.  The problem is padded to 256 x 256 with
// an identity block and cut into 8 panels of 32 rows; a CTA holds 256 / CL panel rows as DMMA accumulator tiles
// (8 x 8, 16 warps, static tile -> warp map), so a step is
//   stage panel k from global (cp.async) -> invert D (4 warps) -> P rows: D^-1 (rows of panel k) or -A_ik D^-1 (others)
//   -> tile += P R on the FP64 tensor cores (panel-k rows start from 0; k columns take P itself)
//   -> the holder of panel k + 1 writes those rows back -> ONE cluster barrier.
// Nothing but the next panel travels through global memory / L2, and the pivot chain computes the next pivot
// speculatively so that only reciprocal + one FMA separate two pivots.  CL = 8 on the GPU (4 row tiles per CTA),
// 1 in the CPU emulator (one block holds all 32 row tiles).
template <int CL>
struct SpdInv256Cfg {
    static constexpr int NB = 32, MP = 256, PS = NB + 4, RS = MP + 8, RPC = MP / CL;
    static constexpr size_t smem_bytes = sizeof(double) * (2 * NB * (NB + 1) + NB * RS + 2 * RPC * PS + 64);
};

template <int CL>
GPB_KERNEL void GPB_CLUSTER(CL) GPB_LAUNCH_BOUNDS(512) spd_inverse256_kernel(
    const double* __restrict__ A, int M, double* __restrict__ W /* [batch, M, M]: out = A^-1 */,
    double* __restrict__ logdet /* [batch] */) {
    typedef SpdInv256Cfg<CL> C;
    GPB_DYN_SMEM(smem);
    constexpr int NB = C::NB, NT = 512, NW = NT / 32, PS = C::PS, RS = C::RS, RPC = C::RPC;
    constexpr int TR = RPC / 8, TPW = TR * (C::MP / 8) / NW;       // row tiles per CTA, tiles per warp
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int rank = cluster_rank();
    const int mat = blockIdx.x / CL;
    const int row_base = rank * RPC;
    double* D = (double*)smem;                 // [NB][NB+1]
    double* D2 = D + NB * (NB + 1);            // [NB][NB+1]  ping-pong twin of D
    double* R = D2 + NB * (NB + 1);            // [NB][RS]    row panel k (values before the step)
    double* P = R + (size_t)NB * RS;           // [RPC][PS]
    double* Ac = P + (size_t)RPC * PS;         // [RPC][PS]   A_ik (values before the step)
    double* s_ld = Ac + (size_t)RPC * PS;      // [1]
    double* s_piv = s_ld + 8;                  // [NB]
    const double* Am = A + (size_t)mat * M * M;
    double* Wm = W + (size_t)mat * M * M;
    const bool even = (M & 1) == 0 && ((size_t)Wm & 15) == 0 && ((size_t)Am & 15) == 0;
    const int nsteps = (M + NB - 1) / NB;
    const int M8 = (M + 7) & ~7;
    const bool cta_live = row_base < M;

    // tile i of this warp: rows trow[i] + g (CTA-local), columns tcol[i] + 2 t4, + 1
    double c0[TPW], c1[TPW];
    GPB_UNROLL
    for (int i = 0; i < TPW; i++) {
        const int tile = warp + NW * i;
        const int row = row_base + (tile % TR) * 8 + g, col = (tile / TR) * 8 + 2 * t4;
        c0[i] = (row < M && col < M) ? Am[(long)row * M + col] : (row == col ? 1.0 : 0.0);
        c1[i] = (row < M && col + 1 < M) ? Am[(long)row * M + col + 1] : (row == col + 1 ? 1.0 : 0.0);
    }
    if (tid == 0) s_ld[0] = 0.0;
    for (int k = 0; k < nsteps; k++) {
        const int k0 = k * NB;
        if (cta_live) {
            // ---- stage panel k (rows / columns past M: the identity padding) ----
            const double* src = k == 0 ? Am : Wm;
            if (even) {
                const int U = M8 >> 1;
                for (int i = tid; i < NB * U; i += NT) {
                    const int c = i / U, jj = 2 * (i - c * U), row = k0 + c;
                    if (row < M) cp_async16_zfill(&R[c * RS + jj], &src[jj < M ? (long)row * M + jj : 0], jj < M);
                    else {
                        R[c * RS + jj] = jj == row ? 1.0 : 0.0;
                        R[c * RS + jj + 1] = jj + 1 == row ? 1.0 : 0.0;
                    }
                }
            } else {
                for (int i = tid; i < NB * M8; i += NT) {
                    const int c = i / M8, jj = i - c * M8, row = k0 + c;
                    if (row < M) cp_async8_zfill(&R[c * RS + jj], &src[jj < M ? (long)row * M + jj : 0], jj < M);
                    else R[c * RS + jj] = jj == row ? 1.0 : 0.0;
                }
            }
            cp_async_commit();
            // ---- this CTA's column block k, out of the register tiles ----
            GPB_UNROLL
            for (int i = 0; i < TPW; i++) {
                const int tile = warp + NW * i;
                const int tc = tile / TR, rl = (tile % TR) * 8 + g;
                if ((tc >> 2) == k) {
                    Ac[rl * PS + tc * 8 + 2 * t4 - k0] = c0[i];
                    Ac[rl * PS + tc * 8 + 2 * t4 + 1 - k0] = c1[i];
                }
            }
            cp_async_wait<0>();
            sync_threads();
            // ---- warps 0-3 invert the diagonal block: lane = row, warp = group of NB/4 columns; ping-pong
            //      between D and D2, one 128-thread named barrier per pivot.  Every thread forms the next pivot
            //      itself from the old block, so the serial chain per pivot is reciprocal + one FMA. ----
            if (tid < 128) {
                constexpr int CG = NB / 4;
                const int a = tid & 31, cg0 = (tid >> 5) * CG;
                double* Dc = D;
                double* Dn = D2;
                for (int b = cg0; b < cg0 + CG; b++)       // (the panel is staged up to column M8 only)
                    Dc[a * (NB + 1) + b] = (k0 + a < M && k0 + b < M) ? R[a * RS + k0 + b] : (a == b ? 1.0 : 0.0);
                sync_group<1, 128>();
                double piv = Dc[0];
                for (int j = 0; j < NB; j++) {
                    // row a != j:  new = old - (a_aj a_jb) / piv ;  row j:  new = a_jb / piv  (multiplier -1, old 0);
                    // column j is overwritten afterwards by the thread that holds it
                    const bool isj = a == j;
                    const double aj = Dc[a * (NB + 1) + j];
                    const double mj = isj ? -1.0 : aj, keep = isj ? 0.0 : 1.0;
                    double tj[CG], ab[CG];
                    GPB_UNROLL
                    for (int bb = 0; bb < CG; bb++) {
                        tj[bb] = mj * Dc[j * (NB + 1) + cg0 + bb];
                        ab[bb] = keep * Dc[a * (NB + 1) + cg0 + bb];
                    }
                    const int j1 = j + 1 < NB ? j + 1 : j;
                    const double n11 = Dc[j1 * (NB + 1) + j1];
                    const double nt = Dc[j1 * (NB + 1) + j] * Dc[j * (NB + 1) + j1];
                    const double ip = pivot_rcp(piv);
                    if (tid == j) s_piv[j] = piv;
                    piv = fma(-nt, ip, n11);                      // pivot j + 1
                    GPB_UNROLL
                    for (int bb = 0; bb < CG; bb++) Dn[a * (NB + 1) + cg0 + bb] = fma(-tj[bb], ip, ab[bb]);
                    if (j >= cg0 && j < cg0 + CG) Dn[a * (NB + 1) + j] = isj ? ip : -aj * ip;
                    sync_group<1, 128>();
                    double* tp = Dc; Dc = Dn; Dn = tp;
                }
                // NB is even: the result sits in D
                if (tid < 32) {
                    double l = log(s_piv[tid]);
                    l = warp_sum(l);
                    if (tid == 0) s_ld[0] += l;
                }
            }
            sync_threads();
            // ---- P rows: D^-1 for the rows of panel k, -A_ik D^-1 for the others (DMMA), 0 for the padding ----
            for (int tile = warp; tile < TR * (NB / 8); tile += NW) {
                const int rl0 = (tile % TR) * 8, cb = (tile / TR) * 8;
                const int row0 = row_base + rl0;
                double p0 = 0.0, p1 = 0.0;
                if (row0 < M8) {
                    if (row0 >= k0 && row0 < k0 + NB) {
                        p0 = D[(row0 + g - k0) * (NB + 1) + cb + 2 * t4];
                        p1 = D[(row0 + g - k0) * (NB + 1) + cb + 2 * t4 + 1];
                    } else {
                        GPB_UNROLL
                        for (int ks = 0; ks < NB / 4; ks++)
                            dmma(p0, p1, -Ac[(rl0 + g) * PS + 4 * ks + t4], D[(4 * ks + t4) * (NB + 1) + cb + g]);
                        if (row0 + g >= M) {
                            p0 = 0.0;
                            p1 = 0.0;
                        }
                    }
                }
                P[(rl0 + g) * PS + cb + 2 * t4] = p0;
                P[(rl0 + g) * PS + cb + 2 * t4 + 1] = p1;
            }
            sync_threads();
            // ---- tiles += P R on the FP64 tensor cores ----
            GPB_UNROLL
            for (int i = 0; i < TPW; i++) {
                const int tile = warp + NW * i;
                const int tc = tile / TR, rl0 = (tile % TR) * 8;
                if (row_base + rl0 >= M || tc * 8 >= M) continue;           // padding tile (warp uniform)
                if (row_base + rl0 >= k0 && row_base + rl0 < k0 + NB) {
                    c0[i] = 0.0;
                    c1[i] = 0.0;
                }
                GPB_UNROLL
                for (int ks = 0; ks < NB / 4; ks++)
                    dmma(c0[i], c1[i], P[(rl0 + g) * PS + 4 * ks + t4], R[(4 * ks + t4) * RS + tc * 8 + g]);
                if ((tc >> 2) == k) {
                    c0[i] = P[(rl0 + g) * PS + tc * 8 + 2 * t4 - k0];
                    c1[i] = P[(rl0 + g) * PS + tc * 8 + 2 * t4 + 1 - k0];
                }
            }
            // ---- the holder of panel k + 1 publishes those rows ----
            if (k + 1 < nsteps) {
                GPB_UNROLL
                for (int i = 0; i < TPW; i++) {
                    const int tile = warp + NW * i;
                    const int row = row_base + (tile % TR) * 8 + g, col = (tile / TR) * 8 + 2 * t4;
                    if (row >= k0 + NB && row < k0 + 2 * NB && row < M) {
                        if (col < M) Wm[(long)row * M + col] = c0[i];
                        if (col + 1 < M) Wm[(long)row * M + col + 1] = c1[i];
                    }
                }
            }
        }
        cluster_barrier();
    }
    GPB_UNROLL
    for (int i = 0; i < TPW; i++) {
        const int tile = warp + NW * i;
        const int row = row_base + (tile % TR) * 8 + g, col = (tile / TR) * 8 + 2 * t4;
        if (row < M) {
            if (col < M) Wm[(long)row * M + col] = c0[i];
            if (col + 1 < M) Wm[(long)row * M + col + 1] = c1[i];
        }
    }
    if (rank == 0 && tid == 0) logdet[mat] = s_ld[0];
}

// FMA-pipe peak microbenchmark (roofline denominator): 8 independent chains per thread
template <typename T>
GPB_KERNEL void fma_peak_kernel(long iters, double* __restrict__ sink) {
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
      a6 = a0 + 6, a7 = a0 + 7;
    const T b = (T)0.999999, c = (T)1e-7;
    for (long i = 0; i < iters; i++) {
        a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
        a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
    }
    // one real store per block keeps every chain live
    double r = (double)(a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7);
    r = warp_sum(r);
    if (threadIdx.x == 0) sink[blockIdx.x] = r;
}

}  // namespace gpb
