// gpb_tail.cuh -- kernels of the replicated O(Dout M^3) tail (SURVEY.md section 8 rows a3, a4, a10
// and the data-independent part of a8 / a9): everything between the parameters and the per-row
// kernels (q(u), cavity, log-partitions: base_models.py:454-488,630-658, aep_models.py:62-114,513-546)
// and between the reduced statistics and the parameter gradients (aep_models.py:252-297,462-511,
// 548-586, base_models.py:490-516, vfe_models.py:363-394,518-541, kernels.py:447-475).
//
// The reference writes this phase as numpy einsum / linalg calls; here it is a short program of
// batched fp64 primitives (GpbTailOp, include/geepee_b200.h):
//   GEMM     C[b] = alpha op(A[b]) op(B[b]) + beta C0[b]      DMMA.8x8x4 tiles, cp.async staging
//   LINCOMB  D[b] = sum_s c_s op(S_s[b]) + c u[b] v[b]^T + c I   (optionally summed over the batch)
//   MATVEC   y[b] = c op(A[b]) x[b] + c' op(A'[b]) x'[b] + c'' w[b] + c''' w'[b]
//   DOTS     scalar = sum_t c_t <x_t, y_t>
//   UNPACK_R / PACK_R   log-diagonal upper-triangular packing of theta_1 = R^T R
//   KHYPER   d tr(M^T Kzz) / d{sf, ls, zu} folded with the direct kernel derivatives
//   GATHER   flat = scale * concat(srcs)
// Matrices are row-major with explicit leading dimensions and per-operand batch strides
// (0 = the operand is shared by the whole batch), so that Kuuinv is never replicated and a
// "sum over output dimensions of X_d^T Y_d" is ONE GEMM with K = Dout * M.
#pragma once

namespace gpb {

struct TailOp {            // mirrors GpbTailOp (include/geepee_b200.h)
    int kind, flags;
    int batch, m, n, k;
    const double* src[6];
    long sstride[6];
    int ld[6];
    double coef[8];
    double* dst;
    long dstride;
    int ldd;
};

// ---- GEMM ---------------------------------------------------------------------------------
// CTA tile BM x BN, K chunks of 32 through a 3-stage cp.async ring (the tail's products are small
// and latency bound: two chunks are always in flight).  VEC: 16-byte copies (all leading dimensions,
// batch strides and base addresses even / 16-byte aligned); otherwise 8-byte copies, so that operands
// with odd leading dimensions work too.  Out-of-range elements are zero filled.  Shared-memory
// strides are 4 resp. 8 mod 16 doubles so that every DMMA fragment load of a warp touches 32
// different banks:
//   A as [m][k] (stride 36)   a = A[g][t] -> 36 g + t        B as [k][n] (stride BN + 8)   b = B[t][g] -> (BN+8) t + g
//   A as [k][m] (stride BM+8) a = A^T[t][g]                  B as [n][k] (stride 36)       b = B^T[g][t]
template <int BM, int BN, int WM, int WN>
struct TailGemmCfg {
    static constexpr int KC = 32, STAGES = 3;
    static constexpr int NT = WM * WN * 32;
    static constexpr int TM = BM / WM / 8, TN = BN / WN / 8;     // DMMA tiles per warp
    static constexpr int LDK = KC + 4;
    static constexpr int A_ELEMS = (BM * LDK > KC * (BM + 8)) ? BM * LDK : KC * (BM + 8);
    static constexpr int B_ELEMS = (BN * LDK > KC * (BN + 8)) ? BN * LDK : KC * (BN + 8);
    static constexpr size_t smem_bytes = sizeof(double) * STAGES * (A_ELEMS + B_ELEMS);
};

// one operand tile -> shared memory.  ROWS x KC logical tile (rows = m or n index, KC = k index);
// `kmajor_src`: the source is stored [k][rows] (contiguous along rows), else [rows][k].
template <int ROWS, int KC, int NT, bool VEC>
GPB_DEVICE void tail_gemm_stage(double* dst, const double* __restrict__ src, int ld, bool kmajor_src, int r0,
                                int k0, int R, int K, int tid) {
    constexpr int LDK = KC + 4, LDR = ROWS + 8;
    constexpr int W = VEC ? 2 : 1;
    if (!kmajor_src) {      // [rows][k]: consecutive threads along k
        for (int i = tid; i < ROWS * (KC / W); i += NT) {
            const int r = i / (KC / W), c = (i % (KC / W)) * W;
            const int left = ((r0 + r) < R) ? (K - (k0 + c)) : 0;
            const double* g = src + (left > 0 ? (long)(r0 + r) * ld + (k0 + c) : 0);
            if (VEC) cp_async16_bytes(dst + r * LDK + c, g, left >= 2 ? 16 : (left == 1 ? 8 : 0));
            else cp_async8_zfill(dst + r * LDK + c, g, left > 0);
        }
    } else {                // [k][rows]: consecutive threads along rows
        for (int i = tid; i < KC * (ROWS / W); i += NT) {
            const int c = i / (ROWS / W), r = (i % (ROWS / W)) * W;
            const int left = ((k0 + c) < K) ? (R - (r0 + r)) : 0;
            const double* g = src + (left > 0 ? (long)(k0 + c) * ld + (r0 + r) : 0);
            if (VEC) cp_async16_bytes(dst + c * LDR + r, g, left >= 2 ? 16 : (left == 1 ? 8 : 0));
            else cp_async8_zfill(dst + c * LDR + r, g, left > 0);
        }
    }
}

template <int BM, int BN, int WM, int WN, bool VEC>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(WM * WN * 32) tail_gemm_kernel(TailOp o) {
    typedef TailGemmCfg<BM, BN, WM, WN> C;
    constexpr int KC = C::KC, NT = C::NT, TM = C::TM, TN = C::TN, LDK = C::LDK, ST = C::STAGES;
    GPB_DYN_SMEM(dsm);
    double* sA = (double*)dsm;                    // [ST][A_ELEMS]
    double* sB = sA + ST * C::A_ELEMS;            // [ST][B_ELEMS]
    const bool ta = o.flags & 1, tb = (o.flags >> 1) & 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const double* A = o.src[0] + (long)b * o.sstride[0];
    const double* B = o.src[1] + (long)b * o.sstride[1];
    const int M = o.m, N = o.n, K = o.k;

    double acc[TM][TN][2];
    GPB_UNROLL
    for (int i = 0; i < TM; i++)
        GPB_UNROLL
        for (int j = 0; j < TN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto stage = [&](int buf, int k0) {
        // op(A) is m x k: stored [m][k] (not transposed) or [k][m]; op(B) is k x n: stored [k][n] or [n][k]
        tail_gemm_stage<BM, KC, NT, VEC>(sA + buf * C::A_ELEMS, A, o.ld[0], ta, m0, k0, M, K, tid);
        tail_gemm_stage<BN, KC, NT, VEC>(sB + buf * C::B_ELEMS, B, o.ld[1], !tb, n0, k0, N, K, tid);
        cp_async_commit();
    };

    const int nk = (K + KC - 1) / KC;
    GPB_UNROLL
    for (int s = 0; s < ST - 1; s++) {
        if (s < nk) stage(s, s * KC);
        else cp_async_commit();          // keep the group count uniform
    }
    for (int kc = 0; kc < nk; kc++) {
        const int buf = kc % ST;
        cp_async_wait<ST - 2>();         // chunk kc has landed (for this thread)
        sync_threads();                  // ... for every thread; and chunk kc-1's buffer is free
        if (kc + ST - 1 < nk) stage((kc + ST - 1) % ST, (kc + ST - 1) * KC);
        else cp_async_commit();
        const double* a_s = sA + buf * C::A_ELEMS;
        const double* b_s = sB + buf * C::B_ELEMS;
        GPB_UNROLL
        for (int ks = 0; ks < KC; ks += 4) {
            double af[TM], bf[TN];
            GPB_UNROLL
            for (int i = 0; i < TM; i++) {
                const int r = wm * (TM * 8) + i * 8 + g;
                af[i] = ta ? a_s[(ks + t) * (BM + 8) + r] : a_s[r * LDK + ks + t];
            }
            GPB_UNROLL
            for (int j = 0; j < TN; j++) {
                const int c = wn * (TN * 8) + j * 8 + g;
                bf[j] = tb ? b_s[c * LDK + ks + t] : b_s[(ks + t) * (BN + 8) + c];
            }
            GPB_UNROLL
            for (int i = 0; i < TM; i++)
                GPB_UNROLL
                for (int j = 0; j < TN; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    // epilogue: dst = alpha acc + beta C0   (C0 = src[2], may alias dst)
    const double alpha = o.coef[0], beta = o.coef[1];
    const double* C0 = o.src[2] ? o.src[2] + (long)b * o.sstride[2] : nullptr;
    double* D = o.dst + (long)b * o.dstride;
    GPB_UNROLL
    for (int i = 0; i < TM; i++)
        GPB_UNROLL
        for (int j = 0; j < TN; j++)
            GPB_UNROLL
            for (int e = 0; e < 2; e++) {
                const int r = m0 + wm * (TM * 8) + i * 8 + g;
                const int c = n0 + wn * (TN * 8) + j * 8 + 2 * t + e;
                if (r < M && c < N) {
                    double v = alpha * acc[i][j][e];
                    if (C0) v += beta * C0[(long)r * o.ld[2] + c];
                    D[(long)r * o.ldd + c] = v;
                }
            }
}

// ---- LINCOMB ------------------------------------------------------------------------------
// dst[b][i][j] = sum_{s<4} coef[s] S_s[b][i][j | j][i]  +  coef[4] u[b][i] v[b][j]  +  coef[5] (i == j)
// flags: bit s (s < 4) = S_s is read transposed; bit 8 = sum over the batch into ONE matrix; bit 9 = src[2],
// src[3] are a second outer-product pair (coef[2] u'[b][i] v'[b][j]) instead of two matrices.
// src[4] = u, src[5] = v (both or neither).  Vectors are 1 x n matrices.
GPB_KERNEL void tail_lincomb_kernel(TailOp o) {
    const long per = (long)o.m * o.n;
    const bool red = (o.flags >> 8) & 1;
    const long total = red ? per : per * o.batch;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % o.n), i = (int)((idx / o.n) % o.m);
        const int b0 = red ? 0 : (int)(idx / per), b1 = red ? o.batch : b0 + 1;
        double acc = 0.0;
        const bool outer2 = (o.flags >> 9) & 1;
        for (int b = b0; b < b1; b++) {
            GPB_UNROLL
            for (int s = 0; s < 4; s++) {
                if (o.src[s] && !(outer2 && s >= 2)) {
                    const double* S = o.src[s] + (long)b * o.sstride[s];
                    acc += o.coef[s] * (((o.flags >> s) & 1) ? S[(long)j * o.ld[s] + i] : S[(long)i * o.ld[s] + j]);
                }
            }
            if (o.src[4])
                acc += o.coef[4] * o.src[4][(long)b * o.sstride[4] + i] * o.src[5][(long)b * o.sstride[5] + j];
            if (outer2)
                acc += o.coef[2] * o.src[2][(long)b * o.sstride[2] + i] * o.src[3][(long)b * o.sstride[3] + j];
            if (i == j) acc += o.coef[5];
        }
        o.dst[(red ? 0 : (long)b0 * o.dstride) + (long)i * o.ldd + j] = acc;
    }
}

// ---- MATVEC -------------------------------------------------------------------------------
// y[b][i] = coef[0] sum_k op(A0[b])[i][k] x0[b][k] + coef[1] sum_k op(A1[b])[i][k] x1[b][k]
//           + coef[2] w0[b][i] + coef[3] w1[b][i]
// src = {A0, x0, A1, x1, w0, w1}; flags bit 0 / bit 1: A0 / A1 stored transposed ([k][m]).
// One warp per output element (M <= 512: these are a few thousand dot products).
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) tail_matvec_kernel(TailOp o) {
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    const long total = (long)o.batch * o.m;
    for (long e = wid; e < total; e += nw) {
        const int b = (int)(e / o.m), i = (int)(e % o.m);
        double acc = 0.0;
        GPB_UNROLL
        for (int s = 0; s < 2; s++) {
            const double* A = o.src[2 * s];
            if (!A) continue;
            A += (long)b * o.sstride[2 * s];
            const double* x = o.src[2 * s + 1] + (long)b * o.sstride[2 * s + 1];
            const bool tr = (o.flags >> s) & 1;
            double p = 0.0;
            for (int k = lane; k < o.k; k += 32)
                p += (tr ? A[(long)k * o.ld[2 * s] + i] : A[(long)i * o.ld[2 * s] + k]) * x[k];
            acc += o.coef[s] * p;
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            if (o.src[4]) acc += o.coef[2] * o.src[4][(long)b * o.sstride[4] + i];
            if (o.src[5]) acc += o.coef[3] * o.src[5][(long)b * o.sstride[5] + i];
            o.dst[(long)b * o.dstride + i] = acc;
        }
    }
}

// ---- DOTS ---------------------------------------------------------------------------------
// dst[0] = (flags bit 0 ? dst[0] : 0) + coef[6] + sum_{t<3} coef[t] sum_{e < ld[2t]} x_t[e] (y_t ? y_t[e] : 1)
// src = {x0, y0, x1, y1, x2, y2}; one CTA (the tail's vectors and matrices are <= Dout M^2 doubles).
GPB_KERNEL void GPB_LAUNCH_BOUNDS(1024) tail_dots_kernel(TailOp o) {
    GPB_SHARED double scratch[32];
    double acc = 0.0;
    GPB_UNROLL
    for (int t = 0; t < 3; t++) {
        const double* x = o.src[2 * t];
        if (!x) continue;
        const double* y = o.src[2 * t + 1];
        double p = 0.0;
        for (long e = threadIdx.x; e < (long)o.ld[2 * t]; e += blockDim.x) p += y ? x[e] * y[e] : x[e];
        acc += o.coef[t] * p;
    }
    const double s = block_sum(acc, scratch);
    if (threadIdx.x == 0) o.dst[0] = ((o.flags & 1) ? o.dst[0] : 0.0) + o.coef[6] + s;
}

// ---- SUM (large arrays, e.g. the [n, Dout] gradient of the layer above) --------------------
// part[block] = sum of this block's grid-stride share; folded by reduce_partials_kernel.
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) tail_sum_partial_kernel(const double* __restrict__ x, long count,
                                                              double* __restrict__ part) {
    GPB_SHARED double scratch[8];
    double p = 0.0;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x)
        p += x[e];
    const double s = block_sum(p, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// ---- R packing (base_models.py:645-653 forward, 505-514 backward) ---------------------------
// UNPACK_R: R[b][i][j] = i < j ? e[b][idx(i,j)] : (i == j ? exp(e[b][idx]) : 0), idx = row-major upper triangle
GPB_DEVICE long triu_index(int i, int j, int M) { return (long)i * M - (long)i * (i - 1) / 2 + (j - i); }
GPB_KERNEL void tail_unpack_r_kernel(TailOp o) {
    const int M = o.m;
    const long per = (long)M * M, total = per * o.batch;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % M), i = (int)((idx / M) % M), b = (int)(idx / per);
        double v = 0.0;
        if (i <= j) {
            v = o.src[0][(long)b * o.sstride[0] + triu_index(i, j, M)];
            if (i == j) v = exp(v);
        }
        o.dst[(long)b * o.dstride + (long)i * o.ldd + j] = v;
    }
}
// PACK_R: out[b][idx(i,j)] = coef[0] dR[b][i][j] (i == j ? R[b][i][i] : 1)   (i <= j)
GPB_KERNEL void tail_pack_r_kernel(TailOp o) {
    const int M = o.m;
    const long P = (long)M * (M + 1) / 2, total = P * o.batch;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / P);
        const long p = idx % P;
        // invert idx(i, j): largest i with triu_index(i, i) <= p
        int i = (int)(((2.0 * M + 1.0) - sqrt((2.0 * M + 1.0) * (2.0 * M + 1.0) - 8.0 * (double)p)) * 0.5);
        while (i > 0 && triu_index(i, i, M) > p) i--;
        while (i + 1 < M && triu_index(i + 1, i + 1, M) <= p) i++;
        const int j = i + (int)(p - triu_index(i, i, M));
        double v = o.coef[0] * o.src[0][(long)b * o.sstride[0] + (long)i * o.ld[0] + j];
        if (i == j) v *= o.src[1][(long)b * o.sstride[1] + (long)i * o.ld[1] + i];
        o.dst[(long)b * o.dstride + p] = v;
    }
}

// ---- KHYPER (kernels.py:447-475 + aep_models.py:455-460,497-504) ---------------------------
// With Kzz = Kuu - jitter I, W = Mm o Kzz:
//   dsf      = 2 sf2 (dsf2 + dvsum) + 2 sum_ab W_ab
//   dls[q]   = dl[q] l_q + sum_ab W_ab (z_aq - z_bq)^2 / l_q^2
//   dzu[a,q] = dzu0[a,q] + sum_b (Mm_ab + Mm_ba) Kzz_ab (z_bq - z_aq) / l_q^2
// src = {Mm, Kuu, zu, ls, sf, stats}, stats = the contiguous record [dzu0[M*D] | dl[D] | dsf2 | dvsum] (the
// tail of a layer's packed statistics).  dst = [dsf | dls[D] | dzu[M*D]] scaled by coef[1].
// coef[0] = jitter; m = M, k = D <= DMAX.  Two launches: CTAs of 8 warps take 8 rows a each (one warp
// per row: dzu[a,:] is complete there) and leave their share of the D + 1 scalar sums in `part`
// ([gridDim.x][DMAX + 1], the op's scratch operand dst + 1 + D + M*D); a one-block finish kernel adds
// the shares in a fixed order (deterministic) and writes dsf and dls.
template <int DMAX>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) tail_khyper_kernel(TailOp o, double* __restrict__ part) {
    GPB_SHARED double s_ls[DMAX + 1][9];      // per-warp partials of the D lengthscale sums + the sf sum
    GPB_SHARED double s_il2[DMAX];
    const int M = o.m, D = o.k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const double* Mm = o.src[0];
    const double* Kuu = o.src[1];
    const double* z = o.src[2];
    const double* dzu0 = o.src[5];
    const double jitter = o.coef[0], scale = o.coef[1];
    if ((int)threadIdx.x < D) s_il2[threadIdx.x] = exp(-2.0 * o.src[3][threadIdx.x]);
    sync_threads();
    double gl[DMAX], gsf = 0.0;
    GPB_UNROLL
    for (int q = 0; q < DMAX; q++) gl[q] = 0.0;
    for (int a = blockIdx.x * nwarp + warp; a < M; a += gridDim.x * nwarp) {
        double gz[DMAX], za[DMAX];
        GPB_UNROLL
        for (int q = 0; q < DMAX; q++) {
            gz[q] = 0.0;
            za[q] = q < D ? z[(long)a * D + q] : 0.0;
        }
        for (int b = lane; b < M; b += 32) {
            const double kzz = Kuu[(long)a * o.ld[1] + b] - (a == b ? jitter : 0.0);
            const double mab = Mm[(long)a * o.ld[0] + b], mba = Mm[(long)b * o.ld[0] + a];
            const double w = mab * kzz, ws = (mab + mba) * kzz;
            gsf += w;
            GPB_UNROLL
            for (int q = 0; q < DMAX; q++) {
                if (q < D) {
                    const double dz = z[(long)b * D + q] - za[q];
                    gz[q] += ws * dz * s_il2[q];
                    gl[q] += w * dz * dz * s_il2[q];
                }
            }
        }
        GPB_UNROLL
        for (int q = 0; q < DMAX; q++) {
            if (q < D) {
                const double s = warp_sum(gz[q]);
                if (lane == 0) o.dst[1 + D + (long)a * D + q] = scale * (dzu0[(long)a * D + q] + s);
            }
        }
    }
    GPB_UNROLL
    for (int q = 0; q < DMAX; q++) {
        if (q < D) {
            const double s = warp_sum(gl[q]);
            if (lane == 0) s_ls[q][warp] = s;
        }
    }
    {
        const double s = warp_sum(gsf);
        if (lane == 0) s_ls[DMAX][warp] = s;
    }
    sync_threads();
    if ((int)threadIdx.x <= D) {
        const int q = (int)threadIdx.x < D ? (int)threadIdx.x : DMAX;
        double s = 0.0;
        for (int w = 0; w < nwarp; w++) s += s_ls[q][w];
        part[(long)blockIdx.x * (DMAX + 1) + q] = s;
    }
}
GPB_KERNEL void tail_khyper_finish_kernel(TailOp o, const double* __restrict__ part, int nblocks, int DMAX) {
    const int M = o.m, D = o.k;
    const double* dl = o.src[5] + (long)M * D;
    const double* dsf2 = dl + D;
    const double* dvsum = dsf2 + 1;
    const double scale = o.coef[1];
    const int q = threadIdx.x;
    if (q < D) {
        double s = 0.0;
        for (int g = 0; g < nblocks; g++) s += part[(long)g * (DMAX + 1) + q];
        o.dst[1 + q] = scale * (dl[q] * exp(o.src[3][q]) + s);
    }
    if (q == 32) {
        double s = 0.0;
        for (int g = 0; g < nblocks; g++) s += part[(long)g * (DMAX + 1) + DMAX];
        const double sf2 = exp(2.0 * o.src[4][0]);
        o.dst[0] = scale * (2.0 * sf2 * (dsf2[0] + dvsum[0]) + 2.0 * s);
    }
}

// ---- GATHER -------------------------------------------------------------------------------
struct GatherArgs {
    static constexpr int MAXN = 40;
    const double* src[MAXN];
    long off[MAXN + 1];      // prefix sums of the element counts
    int n;
    double scale;
    double* dst;
};
GPB_KERNEL void tail_gather_kernel(GatherArgs a) {
    const long total = a.off[a.n];
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        int s = 0;
        while (idx >= a.off[s + 1]) s++;
        a.dst[idx] = a.scale * a.src[s][idx - a.off[s]];
    }
}

// ---- COPY: several contiguous arrays to several destinations in one launch ------------------
struct CopyArgs {
    static constexpr int MAXN = 24;
    const double* src[MAXN];
    double* dst[MAXN];
    long off[MAXN + 1];
    int n;
};
GPB_KERNEL void tail_copy_kernel(CopyArgs a) {
    const long total = a.off[a.n];
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        int s = 0;
        while (idx >= a.off[s + 1]) s++;
        a.dst[s][idx - a.off[s]] = a.src[s][idx - a.off[s]];
    }
}

}  // namespace gpb
