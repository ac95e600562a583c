// gpb_pairsx.cuh -- fp64 pair kernels with the exponent on the FP64 tensor cores (a2 fused with a6 / a9,
// narrow layers: Dout <= 4, Q <= 4; kernels.py:201-234 + aep_models.py:196-198,240-243 + kernels.py:402-444).
//
// The exponent of psi2'[n,p] = cn_n exp(-sum_q c2_nq (mu_nq - zh_pq)^2) is a bilinear form in per-row and
// per-pair features,
//     xs[n,p] = a0_n + sum_q b_nq zh_pq + sum_q c_nq zh_pq^2,     b = 2 kS c2 mu,  c = -kS c2,
// i.e. a GEMM with inner dimension 2Q.  For Q = 2 that is exactly ONE DMMA.8x8x4 per 8 rows x 8 pairs
// (two for Q = 3, 4) with the row constant a0 as the accumulator input: one issue slot instead of the
// eight DFMAs of the SIMT kernel, and no register-operand traffic.  The pair kernels are bound by
// instruction issue, not by the fp64 pipe (ncu: issue slots 60 % busy, fp64 pipe 62-67 %), so the slots
// the tensor instruction frees are what the kernel gains.
//
// Mapping (g = lane / 4, t = lane % 4): a warp owns NPG groups of 8 pairs for the whole kernel and walks
// the rows 8 at a time.  The DMMA result fragment gives lane (g, t) the exponents of row g and pairs
// 2t, 2t+1 of every group; everything after it (exp, contraction with the weights, backward sums) is SIMT
// on those values with the pair constants in registers.  Per-pair sums over rows are therefore spread over
// the 8 lanes with the same t and folded ONCE at the end of the kernel; per-row sums over pairs are spread
// over the 4 lanes of a quad, the 8 warps and the pair chunks (blocks): per-lane partials go to shared
// memory slot of the warp (after two quad shuffles), and after the tile's only barrier the 8 slots are summed and
// one fp64 atomic per (row, value) goes to the global row sums.
//
// Row records are precomputed once per launch by mm_rowfeat_kernel (the divisions / logarithms of
// kernels.py:188-190 are per row, not per row and pair chunk) and staged 32 rows at a time with cp.async:
//     [ b_1..b_Q, c_1..c_Q, 0.. (4 KQ values: the A fragments) | a0 | dv_1..dv_DOC | pad ]     (RLG doubles)
// In shared memory the records are RL = 12 or 20 doubles apart (4 mod 16): the fragment loads of a half
// warp fall into 16 different banks.
#pragma once

namespace gpb {

// pair groups (of 8) a warp owns: the backward keeps 2 (Q + DOC) + ... values per pair in registers, and two
// resident blocks per SM leave 128 registers per thread (ptxas -v: no spills at these settings)
constexpr int mmx_npg(int Q, int DOC, bool BWD) { return !BWD ? 4 : ((2 * Q + 2 * DOC <= 8) ? 2 : 1); }

template <int Q, int DOC, bool BWD>
struct MMXCfg {
    static constexpr int KQ = (2 * Q + 3) / 4;            // DMMA k-steps of the exponent
    static constexpr int NPG = mmx_npg(Q, DOC, BWD);      // pair groups (of 8) per warp
    static constexpr int NRG = 2;                         // row groups (of 8) per loop trip
    static constexpr int PCX = 8 * NPG * 8;               // pairs per block
    static constexpr int TR = 64;                         // rows per staged tile (one barrier per tile)
    static constexpr int RL = KQ == 1 ? 12 : 20;          // row record stride in shared memory (doubles)
    static constexpr int RLG = (4 * KQ + 1 + (BWD ? DOC : 0) + 1) / 2 * 2;   // ... in global memory (even)
    static constexpr int NS = BWD ? 2 * Q : DOC;          // per-row sums
    static constexpr int NSP = (NS + 1) / 2 * 2;
    static constexpr size_t smem_bytes = sizeof(double) * ((size_t)ExpDom<double>::TAB + 2 * TR * RL + 2 * 8 * TR * NSP);
};

// one record per row (rows n .. n_pad-1: null records whose psi2' underflows and whose dv is 0)
template <int Q>
GPB_KERNEL void mm_rowfeat_kernel(const double* __restrict__ mx, const double* __restrict__ vx,
                                  const double* __restrict__ ls, const double* __restrict__ dv, int n, int n_pad,
                                  int Qa, int Do, int DOC, int RL, double* __restrict__ out) {
    constexpr int KQ4 = (2 * Q + 3) / 4 * 4;
    constexpr double kS = ExpDom<double>::S;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_pad; r += gridDim.x * blockDim.x) {
        double* rec = out + (long)r * RL;
        for (int k = 0; k < RL; k++) rec[k] = 0.0;
        if (r >= n) {
            rec[KQ4] = -1.0e5 * kS + ExpBits::OFF;
            continue;
        }
        double lcn = 0.0, a0 = 0.0;
        for (int q = 0; q < Q; q++) {
            double mu = 0, c2 = 0;
            if (q < Qa) {
                mu = mx[(long)r * Qa + q];
                const double lq = exp(2.0 * ls[q]);
                c2 = 1.0 / (2.0 * vx[(long)r * Qa + q] + lq);
                lcn += 0.5 * log(lq * c2);
            }
            const double c2s = c2 * kS;
            rec[q] = 2.0 * c2s * mu;
            rec[Q + q] = -c2s;
            a0 -= c2s * mu * mu;
        }
        rec[KQ4] = lcn * kS + a0 + ExpBits::OFF;
        if (dv)
            for (int d = 0; d < DOC; d++) rec[KQ4 + 1 + d] = d < Do ? dv[(long)r * Do + d] : 0.0;
    }
}

template <int Q, int DOC, bool BWD>
GPB_KERNEL void GPB_LAUNCH_BOUNDS2(256, 2) mm_pairsx_kernel(MMArgs<double> a, const double* __restrict__ rowfeat) {
    typedef MMXCfg<Q, DOC, BWD> C;
    constexpr int KQ = C::KQ, NPG = C::NPG, NRG = C::NRG, TR = C::TR, RL = C::RL, NS = C::NS, NSP = C::NSP;
    constexpr int kTab = ExpDom<double>::TAB;
    constexpr double kS = ExpDom<double>::S;
    GPB_DYN_SMEM(dsm);
    double* s_tab = (double*)dsm;
    double* s_row = s_tab + kTab;                  // [2][TR * RL]
    double* s_acc = s_row + 2 * TR * RL;           // [2][8 warps][TR][NSP] row sums of the tile, one slot per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long PP = a.PP;
    const long pbase = (long)blockIdx.x * C::PCX + (long)warp * NPG * 8;

    // B fragments of the exponent: feature 4 ks + t of pair column g of every group
    double bfr[NPG][KQ];
    GPB_UNROLL
    for (int gi = 0; gi < NPG; gi++) {
        const long p = pbase + gi * 8 + g;
        GPB_UNROLL
        for (int ks = 0; ks < KQ; ks++) {
            const int k = 4 * ks + t;
            double v = 0.0;
            if (k < Q) v = a.zh[(long)k * PP + p];
            else if (k < 2 * Q) {
                const double z = a.zh[(long)(k - Q) * PP + p];
                v = z * z;
            }
            bfr[gi][ks] = v;
        }
    }
    // constants / accumulators of this lane's own pairs 2t, 2t+1 of every group
    double bs[NPG][2][DOC];
    double zh[BWD ? NPG : 1][2][Q], zh2[BWD ? NPG : 1][2][Q];
    double accB[BWD ? NPG : 1][2][DOC], accS1[BWD ? NPG : 1][2][Q];
    GPB_UNROLL
    for (int gi = 0; gi < NPG; gi++)
        GPB_UNROLL
        for (int e = 0; e < 2; e++) {
            const long p = pbase + gi * 8 + 2 * t + e;
            GPB_UNROLL
            for (int d = 0; d < DOC; d++) bs[gi][e][d] = d < a.Do ? a.bs[(long)d * PP + p] : 0.0;
            if (BWD) {
                GPB_UNROLL
                for (int q = 0; q < Q; q++) {
                    zh[gi][e][q] = a.zh[(long)q * PP + p];
                    zh2[gi][e][q] = zh[gi][e][q] * zh[gi][e][q];
                    accS1[gi][e][q] = 0.0;
                }
                GPB_UNROLL
                for (int d = 0; d < DOC; d++) accB[gi][e][d] = 0.0;
            }
        }
    for (int i = tid; i < kTab; i += 256) s_tab[i] = exp_bits_table(i / ExpDom<double>::REP);
    const int lane16 = lane & (ExpDom<double>::REP - 1);
    const int half_bit = 8 + (a.n < 0);      // = 8, kept in a register (LOP3 operand of exp_dom_bits_n)

    const int r_begin = blockIdx.y * a.rows_per_split;            // multiples of TR
    const int r_end = (r_begin + a.rows_per_split) < a.n ? (r_begin + a.rows_per_split) : a.n;

    auto stage = [&](int buf, int t0) {
        constexpr int RLG = C::RLG;
        const double* src = rowfeat + (long)t0 * RLG;
        double* dst = s_row + buf * TR * RL;
        for (int i = tid; i < TR * (RLG / 2); i += 256) {
            const int row = i / (RLG / 2), c = i - row * (RLG / 2);
            cp_async16(dst + row * RL + 2 * c, src + row * RLG + 2 * c);
        }
        cp_async_commit();
    };
    if (r_begin < r_end) stage(0, r_begin);
    cp_async_wait<0>();
    sync_threads();
    int buf = 0;
    for (int t0 = r_begin; t0 < r_end; t0 += TR, buf ^= 1) {
        const int tv = (r_end - t0) < TR ? (r_end - t0) : TR;
        if (t0 + TR < r_end) stage(buf ^ 1, t0 + TR);
        const double* rows = s_row + buf * TR * RL;
        double* acc_t = s_acc + (size_t)(buf * 8 + warp) * TR * NSP;     // this warp's slots of the tile
        GPB_UNROLL_N(1)
        for (int trip = 0; trip < TR / (8 * NRG); trip++) {
            double x[NRG][NPG * 2];
            double v[NRG][NSP];
            double dvr[NRG][BWD ? DOC : 1], brow[NRG][BWD ? Q : 1], crow[NRG][BWD ? Q : 1];
            GPB_UNROLL
            for (int rg = 0; rg < NRG; rg++) {
                const double* rec = rows + (trip * 8 * NRG + rg * 8 + g) * RL;
                double af[KQ];
                GPB_UNROLL
                for (int ks = 0; ks < KQ; ks++) af[ks] = rec[4 * ks + t];
                const double a0 = rec[4 * KQ];
                if (BWD) {
                    GPB_UNROLL
                    for (int q = 0; q < Q; q++) {
                        brow[rg][q] = 0.5 * rec[q];        // kS c2 mu
                        crow[rg][q] = rec[Q + q];          // -kS c2
                    }
                    GPB_UNROLL
                    for (int d = 0; d < DOC; d++) dvr[rg][d] = rec[4 * KQ + 1 + d];
                }
                GPB_UNROLL
                for (int gi = 0; gi < NPG; gi++) {
                    x[rg][2 * gi] = a0;
                    x[rg][2 * gi + 1] = a0;
                    GPB_UNROLL
                    for (int ks = 0; ks < KQ; ks++) dmma(x[rg][2 * gi], x[rg][2 * gi + 1], af[ks], bfr[gi][ks]);
                }
                GPB_UNROLL
                for (int s = 0; s < NSP; s++) v[rg][s] = 0.0;
            }
            GPB_UNROLL
            for (int rg = 0; rg < NRG; rg++) exp_dom_bits_n<NPG * 2>(x[rg], s_tab, lane16, half_bit);
            if (!BWD) {
                GPB_UNROLL
                for (int rg = 0; rg < NRG; rg++)
                    GPB_UNROLL
                    for (int gi = 0; gi < NPG; gi++)
                        GPB_UNROLL
                        for (int e = 0; e < 2; e++)
                            GPB_UNROLL
                            for (int d = 0; d < DOC; d++) v[rg][d] += bs[gi][e][d] * x[rg][2 * gi + e];
            } else {
                double lam[NRG][NPG * 2];
                GPB_UNROLL
                for (int rg = 0; rg < NRG; rg++)
                    GPB_UNROLL
                    for (int gi = 0; gi < NPG; gi++)
                        GPB_UNROLL
                        for (int e = 0; e < 2; e++) {
                            double coef = 0.0;
                            GPB_UNROLL
                            for (int d = 0; d < DOC; d++) coef += dvr[rg][d] * bs[gi][e][d];
                            lam[rg][2 * gi + e] = coef * x[rg][2 * gi + e];
                        }
                GPB_UNROLL
                for (int d = 0; d < DOC; d++)
                    GPB_UNROLL
                    for (int rg = 0; rg < NRG; rg++)
                        GPB_UNROLL
                        for (int gi = 0; gi < NPG; gi++)
                            GPB_UNROLL
                            for (int e = 0; e < 2; e++) accB[gi][e][d] += dvr[rg][d] * x[rg][2 * gi + e];
                GPB_UNROLL
                for (int rg = 0; rg < NRG; rg++)
                    GPB_UNROLL
                    for (int gi = 0; gi < NPG; gi++)
                        GPB_UNROLL
                        for (int e = 0; e < 2; e++)
                            GPB_UNROLL
                            for (int q = 0; q < Q; q++)      // kS c2 (mu - zh) = kS c2 mu + zh (-kS c2)
                                accS1[gi][e][q] += lam[rg][2 * gi + e] * (zh[gi][e][q] * crow[rg][q] + brow[rg][q]);
                GPB_UNROLL
                for (int gi = 0; gi < NPG; gi++)
                    GPB_UNROLL
                    for (int e = 0; e < 2; e++)
                        GPB_UNROLL
                        for (int q = 0; q < Q; q++) {
                            GPB_UNROLL
                            for (int rg = 0; rg < NRG; rg++) v[rg][q] += lam[rg][2 * gi + e] * zh[gi][e][q];
                            GPB_UNROLL
                            for (int rg = 0; rg < NRG; rg++) v[rg][Q + q] += lam[rg][2 * gi + e] * zh2[gi][e][q];
                        }
            }
            // row sums: fold the 4 lanes of a quad (they hold different pairs of the same row) with two
            // shuffles, then one lane per quad stores into the warp's own slot of the tile in shared memory
            // (every (warp, row) meets once per tile: plain stores, 8 lanes = 8 consecutive records; the first
            // version used fp64 shared-memory atomics, which are CAS spin loops that the 8 warps -- walking the
            // same rows in step -- kept colliding in)
            GPB_UNROLL
            for (int rg = 0; rg < NRG; rg++) {
                const int row = trip * 8 * NRG + rg * 8 + g;
                GPB_UNROLL
                for (int s_ = 0; s_ < NS; s_++) {
                    v[rg][s_] += shfl_xor(v[rg][s_], 1);
                    v[rg][s_] += shfl_xor(v[rg][s_], 2);
                }
                if (t == 0) {
                    GPB_UNROLL
                    for (int s_ = 0; s_ < NS; s_++) acc_t[row * NSP + s_] = v[rg][s_];
                }
            }
        }
        cp_async_wait<0>();
        sync_threads();         // the only barrier per tile: partials complete, next tile staged
        // tile complete: one fp64 atomic per (row, value) into the global row sums; re-arm the accumulators
        for (int o = tid; o < TR * NS; o += 256) {
            const int row = o / NS, s_ = o - row * NS;
            const double* ap = s_acc + (size_t)buf * 8 * TR * NSP + row * NSP + s_;
            double acc = 0.0;
            GPB_UNROLL
            for (int w8 = 0; w8 < 8; w8++) acc += ap[(size_t)w8 * TR * NSP];
            if (row < tv) {
                if (BWD) atomic_add(a.rowacc + (long)(t0 + row) * NS + s_, acc);
                else if (s_ < a.Do) atomic_add(a.rowacc + (long)(t0 + row) * a.Do + s_, acc);
            }
        }
    }
    if (BWD) {
        // fold the 8 row lanes (same t) of every per-pair sum, then lanes g == 0 write the records
        double* rec = a.pairpart + (long)blockIdx.y * (DOC + 1 + Q) * PP;
        GPB_UNROLL
        for (int gi = 0; gi < NPG; gi++)
            GPB_UNROLL
            for (int e = 0; e < 2; e++) {
                GPB_UNROLL
                for (int m = 4; m < 32; m <<= 1) {
                    GPB_UNROLL
                    for (int d = 0; d < DOC; d++) accB[gi][e][d] += shfl_xor(accB[gi][e][d], m);
                    GPB_UNROLL
                    for (int q = 0; q < Q; q++) accS1[gi][e][q] += shfl_xor(accS1[gi][e][q], m);
                }
                if (g == 0) {
                    const long p = pbase + gi * 8 + 2 * t + e;
                    const double epv = a.ep[p];
                    double s0 = 0.0;
                    GPB_UNROLL
                    for (int d = 0; d < DOC; d++) {
                        rec[(long)d * PP + p] = epv * accB[gi][e][d];
                        s0 += bs[gi][e][d] * accB[gi][e][d];
                    }
                    rec[(long)DOC * PP + p] = s0;
                    GPB_UNROLL
                    for (int q = 0; q < Q; q++)
                        rec[(long)(DOC + 1 + q) * PP + p] = accS1[gi][e][q] * (1.0 / kS);
                }
            }
    }
}

}  // namespace gpb
