// gpb_umma.cuh -- fp32-psi mode, deterministic-input layer forward on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM, B operand tiles fetched by the TMA engine with cp.async.bulk).
//
//   T[n, d, :] = kfu[n, :] B_d        (aep_models.py:142-158: the O(n Dout M^2) part of the forward)
//   mout[n,d]  = kfu[n,:] . A_d ,     vout[n,d] = sf2 + kfu[n,:] . T[n,d,:]
//
// fp32-psi mode promises fp32-level arithmetic (1e-3 parity), and kind::tf32 keeps only 10 mantissa bits
// of each operand, so every operand is split x = hi + lo with hi exactly representable in TF32 and the
// product is formed as hi*hi + hi*lo + lo*hi (three MMAs, fp32 accumulation in TMEM; the dropped lo*lo
// term is 2^-22 relative) -- the "3xTF32" scheme.
//
// One CTA = 256 threads = one 128-row tile; threads r and r + 128 share row r of the tile (TMEM lane r).
//   * Kfu is generated chunk by chunk (16 pseudo-points) straight into shared memory in the canonical UMMA
//     K-major / no-swizzle layout [k/4][row][4] (8-row x 16-byte core matrices: SBO = 128 B, LBO = rows * 16 B),
//     as a hi and a lo tile; the fp32 values also go to the Ksave buffer the backward streams.
//   * the matching B chunks (hi | lo, pre-split and pre-laid-out once per call by det_umma_prep_kernel:
//     [Do][M/16][2][4][MP][4]) are contiguous in global memory and arrive through ONE cp.async.bulk per
//     output dimension, completion signalled on an mbarrier (expect_tx);
//   * thread 0 issues the tcgen05.mma instructions of a chunk and commits them to the stage's mbarrier;
//     two stages, so the tensor core works on chunk c while the threads generate chunk c + 1;
//   * 512 TMEM columns hold the accumulators of 512 / MP output dimensions (MP = padded M); wider layers
//     take several passes.  Epilogue: tcgen05.ld (thread = row = lane), dot with the row's kfu in fp64,
//     T written to the Tsave buffer.
// Not available in the CPU emulator build (tests/emu): the SIMT fp32 kernel stays the fallback there and
// for calls that do not save Kfu / T (prediction).
#pragma once
#ifndef GPB_CPU_EMU

namespace gpb {

template <int MP>
struct DetUmmaCfg {
    static constexpr int KC = 16;                      // pseudo-points per chunk = 2 MMA K-steps (K = 8 for tf32)
    static constexpr int NCH = MP / KC;
    static constexpr int NI = MP > 256 ? 256 : MP;     // N per instruction
    static constexpr int NH = MP / NI;
    static constexpr int DG = 512 / MP;                // output dims per pass (TMEM: 512 columns)
    static constexpr int A_BYTES = 128 * KC * 4;       // one of (hi, lo)
    static constexpr int B_BYTES = MP * KC * 4;        // one of (hi, lo) of one output dim
    static constexpr int STAGE = 2 * A_BYTES + DG * 2 * B_BYTES;
    // two stages | scaled pseudo-inputs [MP][DP] | mean weights [DG][MP]
    static constexpr size_t smem_bytes(int DP) { return 2 * (size_t)STAGE + sizeof(float) * ((size_t)MP * DP + 512); }
};

GPB_DEVICE uint32_t umma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): start address, leading / stride
// byte offsets in 16-byte units, version 1 (Blackwell)
GPB_DEVICE uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
GPB_DEVICE void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
GPB_DEVICE void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// once per call: scaled pseudo-inputs and the pre-split, pre-laid-out B operand
//   Zs[m][DP]                    = z[m][q] / l_q                      (0 beyond M / D)
//   Bu[d][c][hl][j][b][e]        = hi | lo of Bp[d][a = 16 c + 4 j + e][b]
GPB_KERNEL void det_umma_prep_kernel(const float* __restrict__ Bp, const double* __restrict__ z,
                                     const double* __restrict__ ls, int M, int MP, int D, int DP, int Do,
                                     float* __restrict__ Bu, float* __restrict__ Zs) {
    const long total = (long)Do * MP * MP;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i % MP), a = (int)((i / MP) % MP), d = (int)(i / ((long)MP * MP));
        const float v = Bp[i];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        const float lo = v - hi;
        const int c = a >> 4, j = (a >> 2) & 3, e = a & 3;
        const long base = (((long)d * (MP / 16) + c) * 2) * 4 * MP * 4;
        Bu[base + ((long)j * MP + b) * 4 + e] = hi;
        Bu[base + (long)4 * MP * 4 + ((long)j * MP + b) * 4 + e] = lo;
    }
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)MP * DP; i += (long)gridDim.x * blockDim.x) {
        const int q = (int)(i % DP), m = (int)(i / DP);
        Zs[i] = (m < M && q < D) ? (float)(z[(long)m * D + q] * exp(-ls[q])) : 0.0f;
    }
}

struct DetUmmaArgs {
    const double* x;     // [n, D]
    const double* ls;    // [D]
    const double* sf;    // [1]
    const float* Zs;     // [MP, DP]
    const float* Ap;     // [Do, MP]
    const float* Bu;     // [Do, MP/16, 2, 4, MP, 4]
    int n, M, D, Do;
    double* mout;        // [n, Do]
    double* vout;        // [n, Do]
    float* Ksave;        // [n, MP]
    float* Tsave;        // [n, Do, MP]
};

// one 256-bit store (STG.E.ENL2.256, sm_100): 8 consecutive floats, 32-byte aligned -- a whole sector per instruction
GPB_DEVICE void st_global_v8(float* p, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6), "f"(a7) : "memory");
}

// tcgen05.ld of 8 / 32 consecutive accumulator columns of the warp's 32 lanes
GPB_DEVICE void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
GPB_DEVICE void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

template <int MP, int DP>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) det_fwd_umma_kernel(DetUmmaArgs a) {
    typedef DetUmmaCfg<MP> C;
    constexpr int KC = C::KC, NCH = C::NCH, NI = C::NI, NH = C::NH, DG = C::DG;
    constexpr int ZV = DP / 4;                          // float4 per pseudo-input
    // M <= 256: the thread keeps its 8 kernel values of every chunk in registers (M/2 floats) and reads exactly
    // those accumulator columns in the epilogue; wider layers re-read the row's kfu from the Ksave buffer (L2).
    constexpr bool KREG = MP <= 256;
    constexpr int NKG = KREG ? NCH / 4 : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_full[2], s_done[2];
    __shared__ uint32_t s_tmem;
    float* sZ = (float*)(smem + 2 * (size_t)C::STAGE);  // [MP][DP] scaled pseudo-inputs, whole call
    float* sAp = sZ + MP * DP;                          // [DG][MP] mean weights of the pass
    double* s_part = (double*)smem;                     // [128][2 DG] half 1 -> half 0 (stage 0 is free by then)
    // 256 threads: thread (row = tid % 128, half = tid / 128).  Both halves of a row generate 8 of the 16 kernel
    // values of a chunk and read half of the accumulator columns in the epilogue (warps w and w + 4 own the same
    // 32 TMEM lanes).
    const int tid = threadIdx.x, warp = tid >> 5, rowt = tid & 127, half = tid >> 7;
    const int n = a.n, M = a.M, D = a.D, Do = a.Do;

    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(umma_smem_u32(&s_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(umma_smem_u32(&s_done[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(umma_smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < MP * DP / 4; i += 256) ((float4*)sZ)[i] = __ldg((const float4*)a.Zs + i);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = s_tmem;
    const uint32_t tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (2 at bits 7-9 / 10-12), both K-major,
    // N >> 3 at bits 17-22, M >> 4 at bits 24-28
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NI >> 3) << 17) | ((128u >> 4) << 24);
    const float c0 = (float)(2.0 * a.sf[0] * 1.4426950408889634);      // log2(sf2)
    const double sf2 = exp(2.0 * a.sf[0]);
    uint32_t n_full[2] = {0, 0}, n_done[2] = {0, 0};   // completed phases observed per barrier
    uint32_t used[2] = {0, 0};                          // chunks issued into each stage so far

    const int ntiles = (n + 127) / 128;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row = tile * 128 + rowt;
        const bool rv = row < n;
        float xs[DP];
        GPB_UNROLL
        for (int q = 0; q < DP; q++) xs[q] = (rv && q < D) ? (float)(a.x[(long)row * D + q] * exp(-a.ls[q])) : 0.0f;
        for (int d0 = 0; d0 < Do; d0 += DG) {
            const int dg = (Do - d0) < DG ? (Do - d0) : DG;
            if (d0 == 0 ? (tile == (int)blockIdx.x || Do > DG) : true) {      // mean weights of this pass
                for (int i = tid; i < DG * MP; i += 256) {
                    const int dd = i / MP;
                    sAp[i] = dd < dg ? __ldg(a.Ap + (long)(d0 + dd) * MP + (i - dd * MP)) : 0.0f;
                }
                __syncthreads();
            }
            double mo[DG];
            GPB_UNROLL
            for (int dd = 0; dd < DG; dd++) mo[dd] = 0.0;
            // the chunk loop stays rolled in groups of 4 (a fully unrolled tile is ~190 KB of code and runs out of
            // the instruction cache); the register copy of kfu is filed under (group, chunk in group, point) with
            // a select over the groups, the only dynamic index
            float kreg[NKG][4][8];
#pragma unroll 1
            for (int g = 0; g < NCH / 4; g++)
            GPB_UNROLL
            for (int u4 = 0; u4 < 4; u4++) {
                const int c = 4 * g + u4;
                const int s = u4 & 1;
                unsigned char* st = smem + (size_t)s * C::STAGE;
                float* sA_hi = (float*)st;
                float* sA_lo = (float*)(st + C::A_BYTES);
                unsigned char* sB = st + 2 * C::A_BYTES;
                const int m0 = c * KC + 8 * half;
                // the MMAs that read this stage two chunks ago must have completed
                if (used[s] > n_done[s]) {
                    mbar_wait(umma_smem_u32(&s_done[s]), n_done[s] & 1);
                    n_done[s]++;
                }
                if (tid == 0) {      // B chunks of this pass' output dims through the TMA engine
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                 ::"r"(umma_smem_u32(&s_full[s])), "r"((uint32_t)(dg * 2 * C::B_BYTES)) : "memory");
                    for (int dd = 0; dd < dg; dd++) {
                        const float* src = a.Bu + (((long)(d0 + dd) * NCH + c) * 2) * 4 * MP * 4;
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                            ::"r"(umma_smem_u32(sB + (size_t)dd * 2 * C::B_BYTES)), "l"(src),
                              "r"((uint32_t)(2 * C::B_BYTES)), "r"(umma_smem_u32(&s_full[s])) : "memory");
                    }
                }
                // this thread's 8 kernel values of the chunk: hi / lo tiles in the UMMA layout + Ksave
                float kall[8];
                GPB_UNROLL
                for (int jj = 0; jj < 2; jj++) {
                    const int j = 2 * half + jj;
                    float kv[4], hi[4], lo[4];
                    GPB_UNROLL
                    for (int e = 0; e < 4; e++) {
                        const int i = 4 * jj + e, m = m0 + i;
                        const float4* zp = (const float4*)(sZ + m * DP);
                        float r2 = 0.0f;
                        GPB_UNROLL
                        for (int q4 = 0; q4 < ZV; q4++) {
                            const float4 zq = zp[q4];
                            const float d0_ = xs[4 * q4] - zq.x, d1_ = xs[4 * q4 + 1] - zq.y;
                            const float d2_ = xs[4 * q4 + 2] - zq.z, d3_ = xs[4 * q4 + 3] - zq.w;
                            r2 = fmaf(d0_, d0_, r2);
                            r2 = fmaf(d1_, d1_, r2);
                            r2 = fmaf(d2_, d2_, r2);
                            r2 = fmaf(d3_, d3_, r2);
                        }
                        float k;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(k) : "f"(fmaf(-0.72134752044448170368f, r2, c0)));
                        k = (rv && m < M) ? k : 0.0f;
                        kv[e] = k;
                        if (KREG) {
                            GPB_UNROLL
                            for (int gg = 0; gg < NKG; gg++) kreg[gg][u4][i] = (g == gg) ? k : kreg[gg][u4][i];
                        }
                        hi[e] = __uint_as_float(__float_as_uint(k) & 0xFFFFE000u);
                        lo[e] = k - hi[e];
                        GPB_UNROLL
                        for (int dd = 0; dd < DG; dd++) mo[dd] += (double)k * (double)sAp[dd * MP + m];
                    }
                    *(float4*)(sA_hi + (j * 128 + rowt) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *(float4*)(sA_lo + (j * 128 + rowt) * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    GPB_UNROLL
                    for (int e = 0; e < 4; e++) kall[4 * jj + e] = kv[e];
                }
                if (rv && d0 == 0)
                    st_global_v8(a.Ksave + (long)row * MP + c * KC + 8 * half, kall[0], kall[1], kall[2], kall[3], kall[4],
                                 kall[5], kall[6], kall[7]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> async proxy
                __syncthreads();
                if (tid == 0) {
                    mbar_wait(umma_smem_u32(&s_full[s]), n_full[s] & 1);       // B landed
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    GPB_UNROLL
                    for (int ks = 0; ks < KC / 8; ks++) {
                        const uint64_t a_hi = umma_desc(umma_smem_u32(sA_hi) + ks * 2 * (128 * 16), 128 * 16, 128);
                        const uint64_t a_lo = umma_desc(umma_smem_u32(sA_lo) + ks * 2 * (128 * 16), 128 * 16, 128);
                        for (int dd = 0; dd < dg; dd++) {
                            const uint32_t bh = umma_smem_u32(sB + (size_t)dd * 2 * C::B_BYTES) + ks * 2 * (MP * 16);
                            const uint32_t bl = bh + C::B_BYTES;
                            GPB_UNROLL
                            for (int h = 0; h < NH; h++) {
                                const uint32_t td = tbase + dd * MP + h * NI;
                                const uint64_t b_hi = umma_desc(bh + h * NI * 16, MP * 16, 128);
                                const uint64_t b_lo = umma_desc(bl + h * NI * 16, MP * 16, 128);
                                umma_tf32(td, a_hi, b_hi, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                                umma_tf32(td, a_hi, b_lo, idesc, 1u);
                                umma_tf32(td, a_lo, b_hi, idesc, 1u);
                            }
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(umma_smem_u32(&s_done[s])) : "memory");
                }
                n_full[s]++;        // (only thread 0 waits on it; the phase counters stay uniform)
                used[s]++;
            }
            // all MMAs of this pass: both stages' last commits (commits complete in order)
            GPB_UNROLL
            for (int s = 0; s < 2; s++)
                if (used[s] > n_done[s]) {
                    mbar_wait(umma_smem_u32(&s_done[s]), n_done[s] & 1);
                    n_done[s]++;
                }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // epilogue: thread = (row = TMEM lane, half of the columns); products and short sums in fp32,
            // accumulated in fp64
            double vacc[DG];
            GPB_UNROLL
            for (int dd = 0; dd < DG; dd++) vacc[dd] = 0.0;
            const long rws = rv ? row : 0;
            if (KREG) {
                GPB_UNROLL
                for (int dd = 0; dd < DG; dd++) {
                    if (dd >= dg) break;
                    float* tp = a.Tsave + (rws * Do + d0 + dd) * MP + 8 * half;
                    GPB_UNROLL
                    for (int c4 = 0; c4 < NCH; c4 += 4) {        // 4 chunks' columns per TMEM wait
                        uint32_t v[4][8];
                        GPB_UNROLL
                        for (int u = 0; u < 4; u++) tmem_ld8(tlane + dd * MP + (c4 + u) * KC + 8 * half, v[u]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        double part = 0.0;
                        GPB_UNROLL
                        for (int u = 0; u < 4; u++) {
                            const float* kr = kreg[KREG ? c4 / 4 : 0][u];
                            float t[8];
                            GPB_UNROLL
                            for (int i = 0; i < 8; i++) t[i] = __uint_as_float(v[u][i]);
                            part += (double)(kr[0] * t[0] + kr[1] * t[1]) + (double)(kr[2] * t[2] + kr[3] * t[3]);
                            part += (double)(kr[4] * t[4] + kr[5] * t[5]) + (double)(kr[6] * t[6] + kr[7] * t[7]);
                            if (rv) st_global_v8(tp + (c4 + u) * KC, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
                        }
                        vacc[dd] += part;
                    }
                }
            } else {
                GPB_UNROLL
                for (int dd = 0; dd < DG; dd++) {
                    if (dd >= dg) break;
                    constexpr int NCB = MP / 64;              // 32-column blocks per half
                    float4 kn[8];                             // the row's kfu of the next block (prefetched)
                    {
                        const float4* kp = (const float4*)(a.Ksave + rws * MP + (half * NCB) * 32);
                        GPB_UNROLL
                        for (int u = 0; u < 8; u++) kn[u] = kp[u];
                    }
                    for (int cbl = 0; cbl < NCB; cbl++) {
                        const int cb = half * NCB + cbl;
                        uint32_t v[32];
                        tmem_ld32(tlane + dd * MP + cb * 32, v);
                        float4 kc[8];
                        GPB_UNROLL
                        for (int u = 0; u < 8; u++) kc[u] = kn[u];
                        if (cbl + 1 < NCB) {
                            const float4* kp = (const float4*)(a.Ksave + rws * MP + (cb + 1) * 32);
                            GPB_UNROLL
                            for (int u = 0; u < 8; u++) kn[u] = kp[u];
                        }
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        float4* tp = (float4*)(a.Tsave + (rws * Do + d0 + dd) * MP + cb * 32);
                        double part = 0.0;
                        GPB_UNROLL
                        for (int u = 0; u < 8; u++) {
                            const float t0 = __uint_as_float(v[4 * u]), t1 = __uint_as_float(v[4 * u + 1]);
                            const float t2 = __uint_as_float(v[4 * u + 2]), t3 = __uint_as_float(v[4 * u + 3]);
                            part += (double)(kc[u].x * t0 + kc[u].y * t1) + (double)(kc[u].z * t2 + kc[u].w * t3);
                            if (rv) tp[u] = make_float4(t0, t1, t2, t3);
                        }
                        vacc[dd] += part;
                    }
                }
            }
            if (half == 1) {
                GPB_UNROLL
                for (int dd = 0; dd < DG; dd++) {
                    s_part[rowt * 2 * DG + 2 * dd] = vacc[dd];
                    s_part[rowt * 2 * DG + 2 * dd + 1] = mo[dd];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();        // TMEM and both stages free for the next pass / tile; partner's sums visible
            if (half == 0 && rv) {
                GPB_UNROLL
                for (int dd = 0; dd < DG; dd++) {
                    if (dd >= dg) break;
                    a.vout[(long)row * Do + d0 + dd] = sf2 + vacc[dd] + s_part[rowt * 2 * DG + 2 * dd];
                    a.mout[(long)row * Do + d0 + dd] = mo[dd] + s_part[rowt * 2 * DG + 2 * dd + 1];
                }
            }
            __syncthreads();        // s_part (stage 0) and sAp free
        }
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

// =========================================================================
// fp32-psi mode, moment-matched layer forward (a2 + a6, narrow layers Dout <= 4):
//     vacc[n, d] = sum_p bs[d, p] psi2'[n, p],      psi2'[n, p] = cn_n exp(-sum_q c2_nq (mu_nq - zh_pq)^2)
// The exponent is bilinear in per-row and per-pair features (kernels.py:201-234 expanded),
//     log2 psi2'[n, p] = sum_k F[n, k] G[p, k],   F = log2(e) [2 c2 mu (Q) | -c2 (Q) | log cn - sum c2 mu^2],
//                                                 G = [zh (Q) | zh^2 (Q) | 1],
// so a tile of 128 rows x 256 pairs of exponents is ONE tcgen05.mma (K = 8 features; x3 for the TF32 hi/lo split,
// x2 when 2Q + 1 > 8) into a TMEM accumulator, and what the CUDA cores are left with per row and pair is the
// tcgen05.ld share, one ex2 and Dout FMAs -- the SIMT kernel spends ~10 issue slots there.  The kernel is bound by
// the 16 ex2/clk/SM of the SFU.
//   * one persistent CTA per SM walks row tiles of 128; the row features are built once per tile straight into the
//     UMMA K-major layout (hi / lo tiles);
//   * the pair operand (pre-split, pre-laid-out by mm_tc_prep_kernel: [chunk][hi|lo][k/4][256][4]) and the chunk's
//     weights bs[d][256] arrive by cp.async.bulk into a 4-stage ring (mbarrier expect_tx);
//   * TMEM holds two 256-column accumulators: thread 0 issues the MMAs of chunk c, then all 256 threads (row = lane,
//     half of the columns each) run the exp / FMA epilogue of chunk c - 1.
template <int KS>
struct MMTcCfg {
    static constexpr int NP = 256;                      // pairs per chunk
    static constexpr int A_BYTES = 128 * 8 * KS * 4;    // hi or lo of the row-feature tile
    static constexpr int G_BYTES = NP * 8 * KS * 4;     // hi or lo of one pair chunk
    static constexpr int NSTAGE = 4;
    static constexpr int STAGE = 2 * G_BYTES + 4 * NP * 4;          // + weights of up to 4 output dims
    static constexpr size_t smem_bytes = 2 * (size_t)A_BYTES + (size_t)NSTAGE * STAGE + 128;
};

// Gu[chunk][hl][k/4][pair in chunk][k%4] = hi | lo of G[p][k]
GPB_KERNEL void mm_tc_prep_kernel(const float* __restrict__ zh, long PP, int Q, int KS, float* __restrict__ Gu) {
    const int K = 8 * KS;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < PP * K; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const long p = i / K;
        double v = 0.0;
        if (k < Q) v = zh[(long)k * PP + p];
        else if (k < 2 * Q) {
            const double z = zh[(long)(k - Q) * PP + p];
            v = z * z;
        } else if (k == 2 * Q) v = 1.0;
        const float f = (float)v;
        const float hi = __uint_as_float(__float_as_uint(f) & 0xFFFFE000u);
        const float lo = f - hi;
        const long chunk = p >> 8;
        const int pc = (int)(p & 255);
        float* base = Gu + chunk * (2L * 256 * K);
        base[((long)(k >> 2) * 256 + pc) * 4 + (k & 3)] = hi;
        base[256L * K + ((long)(k >> 2) * 256 + pc) * 4 + (k & 3)] = lo;
    }
}

struct MMTcArgs {
    const double* mx;    // [n, Q]
    const double* vx;    // [n, Q]
    const double* ls;    // [Q]
    const float* Gu;     // [PP/256][2][2 KS][256][4]
    const float* bs;     // [Do, PP]
    int n, Q, Do, d0, dn; // dn output dims d0 .. d0 + dn - 1 in this pass (<= 4)
    long PP;
    double* rowacc;      // [n, Do]
};

// 2^x on the FMA pipe (the SFU's 16 ex2/clk/SM bound mm_pairs_tc_kernel): round-to-nearest split x = n + f with the
// 1.5 * 2^23 trick, degree-5 Taylor polynomial of 2^f on [-1/2, 1/2] (truncation 2.4e-6 relative, fp32-psi mode's bar is
// 1e-3), n added into the exponent field.  x is clamped at -126 (result ~1e-38 instead of 0).
GPB_DEVICE float ex2_fma(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = fmaf(f, 1.3333558e-3f, 9.6181291e-3f);
    p = fmaf(p, f, 5.5504109e-2f);
    p = fmaf(p, f, 2.4022651e-1f);
    p = fmaf(p, f, 6.9314718e-1f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

GPB_DEVICE void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// 288 threads: warps 0-7 build the row features and run the exp / FMA epilogue (thread = row, half of the chunk's
// columns); warp 8 is the producer: one elected thread streams the pair chunks (cp.async.bulk) and issues the MMAs.
// The two sides only meet on mbarriers:
//   s_full[stage]   producer's bulk copies landed                  (expect_tx)
//   s_done[buffer]  MMAs of a chunk complete                       (tcgen05.commit)
//   s_free[buffer]  epilogue finished with a TMEM accumulator      (256 arrivals)
//   s_empty[stage]  epilogue finished with a stage's weights       (256 arrivals)
//   s_aready        row features of the tile are in shared memory  (256 arrivals)
template <int KS, int DN>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(288) mm_pairs_tc_kernel(MMTcArgs a) {
    typedef MMTcCfg<KS> C;
    constexpr int NP = C::NP, NST = C::NSTAGE, K = 8 * KS;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_full[NST], s_empty[NST], s_done[2], s_free[2], s_aready;
    __shared__ uint32_t s_tmem;
    __shared__ double s_part[128][4];
    float* sA_hi = (float*)smem;
    float* sA_lo = (float*)(smem + C::A_BYTES);
    unsigned char* stages = smem + 2 * C::A_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, rowt = tid & 127, half = (tid >> 7) & 1;
    const int n = a.n, Q = a.Q;
    const int nch = (int)(a.PP / NP);

    if (tid == 0) {
        for (int s = 0; s < NST; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(umma_smem_u32(&s_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(umma_smem_u32(&s_empty[s])));
        }
        for (int s = 0; s < 2; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(umma_smem_u32(&s_done[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(umma_smem_u32(&s_free[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(umma_smem_u32(&s_aready)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(umma_smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = s_tmem;
    const int ntiles = (n + 127) / 128;

    if (warp == 8) {
        // ------------------------------ producer ------------------------------
        if (tid == 256) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t tx_bytes = (uint32_t)(2 * C::G_BYTES + DN * NP * 4);
            long sl_load = 0, sl_mma = 0;        // chunk slots loaded / multiplied so far (all tiles)
            int ntile_done = 0;
            auto load = [&](int c) {             // chunk c of the current tile -> ring stage sl_load % NST
                const int s = (int)(sl_load % NST);
                if (sl_load >= NST) mbar_wait(umma_smem_u32(&s_empty[s]), (uint32_t)((sl_load / NST - 1) & 1));
                unsigned char* st = stages + (size_t)s * C::STAGE;
                const uint32_t bar = umma_smem_u32(&s_full[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(umma_smem_u32(st)), "l"(a.Gu + (long)c * (2L * NP * K)), "r"((uint32_t)(2 * C::G_BYTES)),
                               "r"(bar) : "memory");
                GPB_UNROLL
                for (int d = 0; d < DN; d++)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(umma_smem_u32(st + 2 * C::G_BYTES + d * NP * 4)),
                                   "l"(a.bs + (long)(a.d0 + d) * a.PP + (long)c * NP), "r"((uint32_t)(NP * 4)), "r"(bar)
                                 : "memory");
                sl_load++;
            };
            auto mma = [&]() {                   // next chunk in order
                const int s = (int)(sl_mma % NST);
                const int bf = (int)(sl_mma & 1);
                mbar_wait(umma_smem_u32(&s_full[s]), (uint32_t)((sl_mma / NST) & 1));
                if (sl_mma >= 2) mbar_wait(umma_smem_u32(&s_free[bf]), (uint32_t)(((sl_mma >> 1) - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t gh = umma_smem_u32(stages + (size_t)s * C::STAGE), gl = gh + C::G_BYTES;
                const uint32_t td = tbase + (uint32_t)bf * NP;
                GPB_UNROLL
                for (int ks = 0; ks < KS; ks++) {
                    const uint64_t a_hi = umma_desc(umma_smem_u32(sA_hi) + ks * 2 * (128 * 16), 128 * 16, 128);
                    const uint64_t a_lo = umma_desc(umma_smem_u32(sA_lo) + ks * 2 * (128 * 16), 128 * 16, 128);
                    const uint64_t g_hi = umma_desc(gh + ks * 2 * (NP * 16), NP * 16, 128);
                    const uint64_t g_lo = umma_desc(gl + ks * 2 * (NP * 16), NP * 16, 128);
                    umma_tf32(td, a_hi, g_hi, idesc, ks > 0 ? 1u : 0u);
                    umma_tf32(td, a_hi, g_lo, idesc, 1u);
                    umma_tf32(td, a_lo, g_hi, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             ::"r"(umma_smem_u32(&s_done[bf])) : "memory");
                sl_mma++;
            };
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                // MMA(m) is issued as soon as the epilogue of chunk m - 2 has released its accumulator, i.e. while
                // the epilogue of chunk m - 1 runs; the same event frees the ring stage chunk m - 2 + NST goes into.
                // The first MMA of a tile waits for the tile's row features.
                int cl = 0;
                for (; cl < NST && cl < nch; cl++) load(cl);
                mbar_wait(umma_smem_u32(&s_aready), (uint32_t)(ntile_done & 1));
                for (int m = 0; m < nch; m++) {
                    mma();
                    if (m >= 2 && cl < nch) load(cl++);
                }
                ntile_done++;
            }
        }
    } else {
        // ------------------------------ features + epilogue ------------------------------
        const uint32_t tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        long sl = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int row = tile * 128 + rowt;
            const bool rv = row < n;
            {   // row features of the tile (every thread forms its row's and stores its half of the k groups);
                // all MMAs of the previous tile are complete: this thread has waited for the last one
                float f[K];
                GPB_UNROLL
                for (int k = 0; k < K; k++) f[k] = 0.0f;
                const double L2E = 1.4426950408889634074;
                double a0 = rv ? 0.0 : -1.0e5;
                if (rv) {
                    GPB_UNROLL
                    for (int q = 0; q < (K - 1) / 2; q++) {
                        if (q < Q) {
                            const double mu = a.mx[(long)row * Q + q];
                            const double lq = exp(2.0 * a.ls[q]);
                            const double c2 = 1.0 / (2.0 * a.vx[(long)row * Q + q] + lq);
                            a0 += 0.5 * log(lq * c2) - c2 * mu * mu;
                            f[q] = (float)(2.0 * c2 * mu * L2E);
                            GPB_UNROLL
                            for (int k = 0; k < K; k++)
                                if (k == Q + q) f[k] = (float)(-c2 * L2E);
                        }
                    }
                }
                GPB_UNROLL
                for (int k = 0; k < K; k++)
                    if (k == 2 * Q) f[k] = (float)(a0 * L2E);
                GPB_UNROLL
                for (int j = 0; j < 2 * KS; j++) {
                    float hi[4], lo[4];
                    GPB_UNROLL
                    for (int e = 0; e < 4; e++) {
                        hi[e] = __uint_as_float(__float_as_uint(f[4 * j + e]) & 0xFFFFE000u);
                        lo[e] = f[4 * j + e] - hi[e];
                    }
                    if ((j & 1) == half) {
                        *(float4*)(sA_hi + (j * 128 + rowt) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *(float4*)(sA_lo + (j * 128 + rowt) * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(umma_smem_u32(&s_aready));
            double acc[DN];
            GPB_UNROLL
            for (int d = 0; d < DN; d++) acc[d] = 0.0;
            for (int c = 0; c < nch; c++, sl++) {
                const int s = (int)(sl % NST), bf = (int)(sl & 1);
                mbar_wait(umma_smem_u32(&s_full[s]), (uint32_t)((sl / NST) & 1));     // the weights landed
                mbar_wait(umma_smem_u32(&s_done[bf]), (uint32_t)((sl >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const float* sbs = (const float*)(stages + (size_t)s * C::STAGE + 2 * C::G_BYTES);
                const uint32_t tcol = tlane + (uint32_t)bf * NP + half * 128;
                uint32_t v[2][32];
                tmem_ld32(tcol, v[0]);
                GPB_UNROLL
                for (int cb = 0; cb < 4; cb++) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (cb < 3) tmem_ld32(tcol + (cb + 1) * 32, v[(cb + 1) & 1]);   // in flight during the math
                    const int col0 = half * 128 + cb * 32;
                    float part[DN][2];
                    GPB_UNROLL
                    for (int d = 0; d < DN; d++) part[d][0] = part[d][1] = 0.0f;
                    GPB_UNROLL
                    for (int u = 0; u < 8; u++) {
                        float e[4];
                        GPB_UNROLL
                        for (int j = 0; j < 4; j++) {
                            // single-output passes have issue slots to spare: GPB_MM_TC_POLY of every 4 exponentials go
                            // to the FMA pipe, the rest to the SFU (measured: -5.5 % at Dout = 1, +5 % at Dout = 2)
                            if (j >= 4 - (DN == 1 ? GPB_MM_TC_POLY : 0)) e[j] = ex2_fma(__uint_as_float(v[cb & 1][4 * u + j]));
                            else asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[j]) : "f"(__uint_as_float(v[cb & 1][4 * u + j])));
                        }
                        GPB_UNROLL
                        for (int d = 0; d < DN; d++) {
                            const float4 w = *(const float4*)(sbs + d * NP + col0 + 4 * u);
                            part[d][0] = fmaf(e[0], w.x, part[d][0]);
                            part[d][1] = fmaf(e[1], w.y, part[d][1]);
                            part[d][0] = fmaf(e[2], w.z, part[d][0]);
                            part[d][1] = fmaf(e[3], w.w, part[d][1]);
                        }
                    }
                    GPB_UNROLL
                    for (int d = 0; d < DN; d++) acc[d] += (double)(part[d][0] + part[d][1]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(umma_smem_u32(&s_free[bf]));      // accumulator and stage back to the producer
                mbar_arrive(umma_smem_u32(&s_empty[s]));
            }
            if (half == 1) {
                GPB_UNROLL
                for (int d = 0; d < DN; d++) s_part[rowt][d] = acc[d];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0 && rv) {
                GPB_UNROLL
                for (int d = 0; d < DN; d++) a.rowacc[(long)row * a.Do + a.d0 + d] = acc[d] + s_part[rowt][d];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");     // s_part free
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

// =========================================================================
// fp32-psi mode, deterministic-input layer: dB[d] = sum_n dv[n, d] kfu[n, :] kfu[n, :]^T  (aep_models.py:493) from the
// saved fp32 Kfu, on tcgen05.  Here the GEMM's inner dimension is the data row: A = (dv kfu)^T and B = kfu^T must
// have 4 consecutive ROWS of one pseudo-point in each 16-byte piece of the K-major UMMA layout, so the saved tile
// [row][pseudo-point] is transposed on its way into the operand tiles:
//   * cp.async ring (4 stages) of raw 16-row chunks [16][MP] + their dv;
//   * thread = pseudo-point m: reads its column of the chunk (a warp reads 32 consecutive floats: conflict free),
//     forms dv k and the TF32 hi / lo splits and stores four 16-byte pieces per operand tile
//     (A_hi, A_lo, B_hi, B_lo; a warp stores 512 contiguous bytes: conflict free);
//   * thread 0 issues the 3xTF32 MMAs of the chunk: rows 0-127 of the output against all MP columns and, for
//     MP = 256, rows 128-255 against columns 128-255 only (upper block triangle: 384 TMEM columns);
//   * fp32 accumulation in TMEM is flushed to the CTA's fp64 partial record every 2048 rows (coalesced, the
//     record is stored transposed), so no sum runs longer than that in fp32;
//   * grid = (row splits, Dout); det_syrk_finish_kernel (tr = 1) folds the splits.
template <int MP>
struct SyrkUmmaCfg {
    static constexpr int RK = 16;                        // rows per chunk = 2 MMA K-steps
    static constexpr int RAW = RK * MP * 4 + 128;        // raw chunk + dv[16]
    static constexpr int NRAW = 4;
    static constexpr int TILE = RK * MP * 4;             // one operand tile (hi or lo of A or B)
    static constexpr int OST = 4 * TILE;                 // operand stage: A_hi | A_lo | B_hi | B_lo
    static constexpr int NCOL = MP == 256 ? 384 : 128;   // accumulator columns
    static constexpr int FLUSH = 2048 / RK;              // chunks between flushes
    static constexpr size_t smem_bytes = (size_t)NRAW * RAW + 2 * (size_t)OST + 128;
};

struct SyrkUmmaArgs {
    const float* Ksave;   // [n, MP]
    const double* dv;     // [n, Do]
    int n, Do, rows_per_split;
    double* part;         // [nsplit][Do][nbu][128 * 128], blocks stored transposed
};

template <int MP>
GPB_KERNEL void GPB_LAUNCH_BOUNDS(256) det_syrk_umma_kernel(SyrkUmmaArgs a) {
    typedef SyrkUmmaCfg<MP> C;
    constexpr int RK = C::RK, NRAW = C::NRAW, NCOL = C::NCOL;
    constexpr int NBU = MP == 256 ? 3 : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_done[2];
    __shared__ uint32_t s_tmem;
    unsigned char* raws = smem;
    unsigned char* osts = smem + (size_t)NRAW * C::RAW;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = blockIdx.y, split = blockIdx.x;
    const long row_lo = (long)split * a.rows_per_split;
    const long row_hi = (row_lo + a.rows_per_split) < a.n ? (row_lo + a.rows_per_split) : a.n;
    const int nchunk = row_hi > row_lo ? (int)((row_hi - row_lo + RK - 1) / RK) : 0;

    if (tid == 0) {
        for (int s = 0; s < 2; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(umma_smem_u32(&s_done[s])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(umma_smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = s_tmem;
    const uint32_t tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t idesc_full = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(MP >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_half = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((128u >> 4) << 24);

    auto issue = [&](int c) {          // raw chunk c -> ring stage c % NRAW (all threads; an empty group past the end)
        if (c < nchunk) {
            unsigned char* st = raws + (size_t)(c % NRAW) * C::RAW;
            const long r0 = row_lo + (long)c * RK;
            for (int i = tid; i < RK * (MP / 4); i += 256) {
                const int r = i / (MP / 4), m4 = i - r * (MP / 4);
                const bool ok = r0 + r < row_hi;
                cp_async16_zfill(st + (size_t)(r * MP + 4 * m4) * 4, a.Ksave + (ok ? (r0 + r) * MP + 4 * m4 : 0), ok);
            }
            if (tid < RK) {
                const bool ok = r0 + tid < row_hi;
                cp_async8_zfill(st + RK * MP * 4 + tid * 8, a.dv + (ok ? (r0 + tid) * a.Do + d : 0), ok);
            }
        }
        cp_async_commit();
    };
    // flush: accumulators (+)= into this CTA's partial record; thread = output row (TMEM lane), half of the columns
    auto flush = [&](bool first) {
        double* rec = a.part + ((long)split * a.Do + d) * NBU * (128 * 128);
        const int c_lo = (warp >> 2) * (NCOL / 2), c_hi = c_lo + NCOL / 2;
        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tlane + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            GPB_UNROLL
            for (int j = 0; j < 32; j++) {
                const int c = c0 + j;
                // columns 0..MP-1: output rows 0..127, columns 0..MP-1; columns 256..383: block (1, 1)
                const int ub = c >> 7, jj = c & 127;
                double* o = rec + (long)ub * (128 * 128) + jj * 128 + (tid & 127);
                const double x = (double)__uint_as_float(v[j]);
                *o = first ? x : *o + x;
            }
        }
    };

    uint32_t used[2] = {0, 0}, seen[2] = {0, 0};
    bool first_flush = true;
    int since = 0;                      // chunks accumulated in TMEM since the last flush
    for (int c = 0; c < NRAW - 1; c++) issue(c);
    for (int c = 0; c < nchunk; c++) {
        const int s = c & 1;
        issue(c + NRAW - 1);
        cp_async_wait<NRAW - 1>();      // raw chunk c has landed (this thread's copies)
        if (used[s] > seen[s]) {        // the MMAs that read operand stage s two chunks ago are complete
            mbar_wait(umma_smem_u32(&s_done[s]), seen[s] & 1);
            seen[s]++;
        }
        __syncthreads();                // ... everyone's copies
        const float* raw = (const float*)(raws + (size_t)(c % NRAW) * C::RAW);
        const double* rdv = (const double*)(raws + (size_t)(c % NRAW) * C::RAW + RK * MP * 4);
        unsigned char* ost = osts + (size_t)s * C::OST;
        if (tid < MP) {
            GPB_UNROLL
            for (int kg = 0; kg < RK / 4; kg++) {
                float bh[4], bl[4], ah[4], al[4];
                GPB_UNROLL
                for (int e = 0; e < 4; e++) {
                    const float k = raw[(4 * kg + e) * MP + tid];
                    const float ka = k * (float)rdv[4 * kg + e];
                    bh[e] = __uint_as_float(__float_as_uint(k) & 0xFFFFE000u);
                    bl[e] = k - bh[e];
                    ah[e] = __uint_as_float(__float_as_uint(ka) & 0xFFFFE000u);
                    al[e] = ka - ah[e];
                }
                const size_t off = (size_t)kg * (MP * 16) + (size_t)tid * 16;
                *(float4*)(ost + off) = make_float4(ah[0], ah[1], ah[2], ah[3]);
                *(float4*)(ost + C::TILE + off) = make_float4(al[0], al[1], al[2], al[3]);
                *(float4*)(ost + 2 * C::TILE + off) = make_float4(bh[0], bh[1], bh[2], bh[3]);
                *(float4*)(ost + 3 * C::TILE + off) = make_float4(bl[0], bl[1], bl[2], bl[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = umma_smem_u32(ost), al = ah + C::TILE, bh = ah + 2 * C::TILE, bl = ah + 3 * C::TILE;
            GPB_UNROLL
            for (int ks = 0; ks < RK / 8; ks++) {
                const uint32_t ko = ks * 2 * (MP * 16);
                const uint32_t acc = (since > 0 || ks > 0) ? 1u : 0u;
                // output rows 0..127 against all MP columns
                umma_tf32(tbase, umma_desc(ah + ko, MP * 16, 128), umma_desc(bh + ko, MP * 16, 128), idesc_full, acc);
                umma_tf32(tbase, umma_desc(ah + ko, MP * 16, 128), umma_desc(bl + ko, MP * 16, 128), idesc_full, 1u);
                umma_tf32(tbase, umma_desc(al + ko, MP * 16, 128), umma_desc(bh + ko, MP * 16, 128), idesc_full, 1u);
                if (MP == 256) {        // output rows 128..255 against columns 128..255
                    const uint32_t ro = 128 * 16;
                    umma_tf32(tbase + 256, umma_desc(ah + ko + ro, MP * 16, 128), umma_desc(bh + ko + ro, MP * 16, 128), idesc_half, acc);
                    umma_tf32(tbase + 256, umma_desc(ah + ko + ro, MP * 16, 128), umma_desc(bl + ko + ro, MP * 16, 128), idesc_half, 1u);
                    umma_tf32(tbase + 256, umma_desc(al + ko + ro, MP * 16, 128), umma_desc(bh + ko + ro, MP * 16, 128), idesc_half, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(umma_smem_u32(&s_done[s])) : "memory");
        }
        used[s]++;
        since++;
        if (since == C::FLUSH || c == nchunk - 1) {
            GPB_UNROLL
            for (int q = 0; q < 2; q++)
                if (used[q] > seen[q]) {
                    mbar_wait(umma_smem_u32(&s_done[q]), seen[q] & 1);
                    seen[q]++;
                }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            flush(first_flush);
            first_flush = false;
            since = 0;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
    }
    if (nchunk == 0) {                   // empty split: a zero record
        double* rec = a.part + ((long)split * a.Do + d) * NBU * (128 * 128);
        for (int i = tid; i < NBU * 128 * 128; i += 256) rec[i] = 0.0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

}  // namespace gpb
#endif  // GPB_CPU_EMU
