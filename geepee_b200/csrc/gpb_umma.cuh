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
                    if (rv && d0 == 0)
                        *(float4*)(a.Ksave + (long)row * MP + c * KC + 4 * j) = make_float4(kv[0], kv[1], kv[2], kv[3]);
                }
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
                            if (rv) {
                                float4* tq = (float4*)(tp + (c4 + u) * KC);
                                tq[0] = make_float4(t[0], t[1], t[2], t[3]);
                                tq[1] = make_float4(t[4], t[5], t[6], t[7]);
                            }
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

}  // namespace gpb
#endif  // GPB_CPU_EMU
