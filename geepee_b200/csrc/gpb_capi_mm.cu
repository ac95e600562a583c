// gpb_capi_mm.cu -- moment-matched (uncertain-input) layer entry points (a2+a6, a2+a9).
#include "gpb_common.cuh"
#include "gpb_pairsx.cuh"
#include "gpb_umma.cuh"

namespace {

// ------------------------------- moment-matched layer -----------------------------------
struct MMPlan {
    int Qt, DOC, RP, PC;
    long P, PP;
    int nchunks, nsplit, rows_per_split, npass;
    int rows_grid, cols_grid, cols_rows_per_block, fwd_grid;
    // fp64 backward of wide layers (Do > 4, Qt <= 8): tensor-core kernel over 64-pair chunks
    int wide_mma, w_nchunks, w_nsplit, w_rows_per_split;
    // fp64 narrow layers (Do <= 4, Qt <= 4): pair kernels with the exponent on the FP64 tensor cores
    int xpath, x_nchunks, x_nsplit, x_rows_per_split, x_rlg;
    long x_npad;
    // fp32 narrow layers, forward: exponent GEMM on tcgen05 (gpb_umma.cuh), K = 8 tc_ks features
    int tc, tc_ks;
    // workspace byte offsets are carved in order by mm_carve
};
// row splits of the tensor-exponent pair kernels: two resident blocks per SM; take the split count (<= 48,
// >= 4 row tiles per block) whose block count fills whole waves best, the smallest one on ties
void mmx_splits(int n, int nchunks, int* nsplit, int* rows_per_split) {
    const long slots = 2L * sm_count();
    int best = 1;
    double best_eff = 0.0;
    for (int ns = 1; ns <= 48; ns++) {
        if (ns > 1 && cdiv(n, ns) < 4 * 64) break;
        const long rps = cdiv(cdiv(n, ns), 64) * 64;
        const long total = (long)nchunks * cdiv(n, rps);
        const double eff = (double)total / (double)(cdiv(total, slots) * slots);
        if (eff > best_eff + 0.01) { best_eff = eff; best = ns; }
    }
    const long rps = cdiv(cdiv(n, best), 64) * 64;      // whole row tiles (MMXCfg::TR)
    *rows_per_split = (int)rps;
    *nsplit = (int)cdiv(n, rps);
}
int q_template(int Q) {
    if (Q <= 6) return Q;
    if (Q <= 8) return 8;
    if (Q <= 16) return 16;
    return -1;
}
// pairs per thread: mirrors gpb::MMCfg<T,Q,DOC>::RP (fp32 holds twice as many)
int mm_rp(int tbytes, int Qt, int DOC) {
    int rp = (2 * Qt + 2 * DOC + 2) <= GPB_MM_RP4_MAX ? 4 : ((2 * Qt + 2 * DOC + 2) <= 26 ? 2 : 1);
    if (GPB_MM_RP64 > 0 && GPB_MM_RP64 < rp) rp = GPB_MM_RP64;
    return tbytes == 4 ? 2 * rp : rp;
}

MMPlan mm_plan(int tbytes, int n, int M, int Q, int Do, int backward) {
    MMPlan p;
    p.Qt = q_template(Q);
    // output dims held in registers per pass: exact for Do <= 4, else 8 / 16 / 32 with one pair
    // per thread (wide layers, e.g. SGPLVM with Do = 50: 2 passes instead of 13)
    p.DOC = Do == 1 ? 1 : (Do == 2 ? 2 : (Do <= 4 ? 4 : (Do <= 8 ? 8 : (Do <= 16 ? 16 : 32))));
    // forward: the DOC row sums per pair chunk go through the warp transposition, whose cost grows
    // with DOC while only one pair per thread is left to amortise it -> 4 dims per pass there
    if (!backward && p.DOC > 4) p.DOC = 4;
    p.npass = (int)cdiv(Do, p.DOC);
    p.RP = mm_rp(tbytes, p.Qt, p.DOC);
    p.PC = 256 * p.RP;
    p.P = (long)M * (M + 1) / 2;
    p.PP = cdiv(p.P, p.PC) * p.PC;
    p.nchunks = (int)(p.PP / p.PC);
    // row splits: as many blocks as fit an integer number of waves (1 block/SM), 8 waves when the
    // problem is large enough, never less than 4 row tiles per block
    const int TR = 32;
    int best = 1;
    for (int waves = 8; waves >= 1; waves--) {
        int ns = (int)((long)waves * sm_count() / p.nchunks);
        if (ns >= 1 && cdiv(n, ns) >= 4 * TR) { best = ns; break; }
    }
    long rps = cdiv(n, best);
    rps = cdiv(rps, TR) * TR;
    p.rows_per_split = (int)rps;
    p.nsplit = (int)cdiv(n, rps);
    p.rows_grid = (int)cdiv(n, 8);               // one warp per row, 8 rows per block trip
    if (p.rows_grid > 8 * sm_count()) p.rows_grid = 8 * sm_count();
    if (p.rows_grid < 1) p.rows_grid = 1;
    p.fwd_grid = (int)cdiv(n, 32);               // psi1 forward: row tiles of 32
    if (p.fwd_grid > 2 * sm_count()) p.fwd_grid = 2 * sm_count();
    if (p.fwd_grid < 1) p.fwd_grid = 1;
    int cb = 2 * sm_count();
    long crpb = cdiv(n, cb);
    if (crpb < 32) crpb = 32;
    p.cols_rows_per_block = (int)crpb;
    p.cols_grid = (int)cdiv(n, crpb);
    p.xpath = (GPB_MM_XPATH && (!backward || GPB_MM_XPATH_BWD) && tbytes == 8 && Do <= 4 && p.Qt <= 4) ? 1 : 0;
    p.x_nchunks = p.x_nsplit = p.x_rows_per_split = p.x_rlg = 0;
    p.x_npad = 0;
    if (p.xpath) {
        const int kq = (2 * p.Qt + 3) / 4;
        p.x_nchunks = (int)(p.PP / (64 * gpb::mmx_npg(p.Qt, p.DOC, backward != 0)));
        p.x_rlg = (4 * kq + 1 + (backward ? p.DOC : 0) + 1) / 2 * 2;
        p.x_npad = cdiv(n, 64) * 64;
        mmx_splits(n, p.x_nchunks, &p.x_nsplit, &p.x_rows_per_split);
    }
    p.tc = 0;
    p.tc_ks = (2 * Q + 1 + 7) / 8;
#ifndef GPB_CPU_EMU
    p.tc = (GPB_MM_TC && !backward && tbytes == 4 && Do <= 4 && p.tc_ks <= 2 && p.PP % 256 == 0) ? 1 : 0;
#endif
    p.wide_mma = (backward && tbytes == 8 && Do > 4 && p.Qt <= 8) ? 1 : 0;
    p.w_nchunks = (int)(p.PP / 64);
    p.w_nsplit = 1;
    p.w_rows_per_split = n;
    if (p.wide_mma) {   // one CTA per SM: row splits that fill an integer number of waves
        int bestw = 1;
        for (int waves = 8; waves >= 1; waves--) {
            int ns = (int)((long)waves * sm_count() / p.w_nchunks);
            if (ns >= 1 && cdiv(n, ns) >= 4 * TR) { bestw = ns; break; }
        }
        long wr = cdiv(cdiv(n, bestw), TR) * TR;
        p.w_rows_per_split = (int)wr;
        p.w_nsplit = (int)cdiv(n, wr);
    }
    return p;
}

template <typename T>
struct MMWs {
    T *zh, *ep, *bs;
    double *rowacc, *pairpart, *pairsum, *rowpart, *rowsum, *colpart, *colsum, *dZ2, *dlW;
    double* rowfeat;
    float* gu;           // tcgen05 forward: pre-split pair features
    size_t bytes;
};
template <typename T>
MMWs<T> mm_carve(const MMPlan& p, int n, int M, int Q, int Do, int backward, void* ws, size_t cap) {
    MMWs<T> w;
    Carver cv(ws, cap);
    w.zh = (T*)cv.take(sizeof(T) * p.Qt * p.PP);
    w.ep = (T*)cv.take(sizeof(T) * p.PP);
    w.bs = (T*)cv.take(sizeof(T) * Do * p.PP);
    w.rowfeat = p.xpath ? (double*)cv.take(sizeof(double) * (size_t)p.x_npad * p.x_rlg) : nullptr;
    w.gu = p.tc ? (float*)cv.take(sizeof(float) * (size_t)p.PP * 8 * p.tc_ks * 2) : nullptr;
    if (!backward) {
        w.rowacc = nullptr;   // the forward accumulates straight into the caller's vacc[n,Do]
        w.pairpart = w.pairsum = w.rowpart = w.rowsum = w.colpart = w.colsum = w.dZ2 = w.dlW = nullptr;
    } else {
        w.rowacc = (double*)cv.take(sizeof(double) * (size_t)n * (2 * p.Qt));
        size_t npart = (size_t)(p.xpath ? p.x_nsplit : p.nsplit) * (p.DOC + 1 + p.Qt) * p.PP;
        if (p.wide_mma) npart = (size_t)p.w_nsplit * (Do + 1 + p.Qt) * p.PP;
        w.pairpart = (double*)cv.take(sizeof(double) * npart);
        w.pairsum = (double*)cv.take(sizeof(double) * (size_t)(Do + 1 + p.Qt) * p.PP);
        w.rowpart = (double*)cv.take(sizeof(double) * (size_t)p.rows_grid * (2 + Q));
        w.rowsum = (double*)cv.take(sizeof(double) * (2 + Q));
        w.colpart = (double*)cv.take(sizeof(double) * (size_t)p.cols_grid * ((size_t)Do * M + (size_t)M * Q));
        w.colsum = (double*)cv.take(sizeof(double) * ((size_t)Do * M + (size_t)M * Q));
        w.dZ2 = (double*)cv.take(sizeof(double) * (size_t)M * Q);
        w.dlW = (double*)cv.take(sizeof(double) * (size_t)M * Q);
    }
    w.bytes = cv.off;
    return w;
}

// The fp64 pair kernels keep the replicated exp table in dynamic shared memory; together with
// their static buffers they exceed the 48 KB a kernel gets without opting in.
template <typename K>
int mm_pairs_smem(K kern, size_t dyn) {
#ifndef GPB_CPU_EMU
    if (dyn == 0) return GPB_OK;
    static K done = nullptr;   // one static per kernel instantiation
    if (done == kern) return GPB_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024);
    if (e != cudaSuccess) return fail(GPB_ERR_CUDA, "mm_pairs: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    done = kern;
#else
    (void)kern; (void)dyn;
#endif
    return GPB_OK;
}

// tensor-exponent pair kernels (gpb_pairsx.cuh): row records first, then the pair kernel
template <int Q, int DOC, bool BWD>
int mm_pairsx_launch(const MMPlan& p, gpb::MMArgs<double> a, double* rowfeat, void* stream) {
    typedef gpb::MMXCfg<Q, DOC, BWD> C;
    static_assert(C::PCX == 64 * gpb::mmx_npg(Q, DOC, BWD), "mm_plan assumes this pair-chunk size");
    auto kern = gpb::mm_pairsx_kernel<Q, DOC, BWD>;
    if (mm_pairs_smem(kern, C::smem_bytes)) return GPB_ERR_CUDA;
    if (p.x_rlg != C::RLG) return fail(GPB_ERR_ARG, "mm_pairsx: record length mismatch (%d vs %d)", p.x_rlg, C::RLG);
    a.rows_per_split = p.x_rows_per_split;
    prof_begin(BWD ? 4 : 3, stream);
    auto feat = gpb::mm_rowfeat_kernel<Q>;
    GPB_LAUNCH(feat, dim3(elementwise_grid(p.x_npad)), dim3(256), 0, stream, a.mx, a.vx, a.ls,
               BWD ? a.dv : (const double*)nullptr, a.n, (int)p.x_npad, a.Qa, a.Do, DOC, C::RLG, rowfeat);
    GPB_LAUNCH(kern, dim3(p.x_nchunks, p.x_nsplit), dim3(256), C::smem_bytes, stream, a, (const double*)rowfeat);
    prof_end(BWD ? 4 : 3, stream);
    return GPB_OK;
}
template <int Q, bool BWD>
int mm_pairsx_doc(const MMPlan& p, const gpb::MMArgs<double>& a, double* rowfeat, void* stream) {
    switch (p.DOC) {
        case 1: return mm_pairsx_launch<Q, 1, BWD>(p, a, rowfeat, stream);
        case 2: return mm_pairsx_launch<Q, 2, BWD>(p, a, rowfeat, stream);
        case 4: return mm_pairsx_launch<Q, 4, BWD>(p, a, rowfeat, stream);
    }
    return fail(GPB_ERR_ARG, "mm_pairsx: bad DOC %d", p.DOC);
}
template <bool BWD>
int mm_pairsx_dispatch(const MMPlan& p, const gpb::MMArgs<double>& a, double* rowfeat, void* stream) {
    if constexpr (BWD && !GPB_MM_XPATH_BWD) {
        (void)p; (void)a; (void)rowfeat; (void)stream;
        return fail(GPB_ERR_ARG, "mm_pairsx: backward variant not built (GPB_MM_XPATH_BWD)");
    } else
    switch (p.Qt) {
        case 1: return mm_pairsx_doc<1, BWD>(p, a, rowfeat, stream);
        case 2: return mm_pairsx_doc<2, BWD>(p, a, rowfeat, stream);
        case 3: return mm_pairsx_doc<3, BWD>(p, a, rowfeat, stream);
        case 4: return mm_pairsx_doc<4, BWD>(p, a, rowfeat, stream);
    }
    return fail(GPB_ERR_ARG, "mm_pairsx: input dim template %d unsupported", p.Qt);
}

template <typename T, int Q, int DOC, bool BWD>
int mm_pairs_launch(const MMPlan& p, const gpb::MMArgs<T>& a, void* stream) {
    constexpr int NS = BWD ? 2 * Q : DOC;
    const size_t smem = sizeof(double) * gpb::ExpDom<T>::TAB + gpb::RowXpose<T, NS>::kBytes;
    prof_begin(BWD ? 4 : 3, stream);
    if constexpr (BWD && DOC == 32) {
        if (p.npass > 1) {   // Do > 32: generic multi-pass kernel
            auto kern = gpb::mm_pairs_kernel<T, Q, DOC, BWD, true>;
            if (mm_pairs_smem(kern, smem)) return GPB_ERR_CUDA;
            GPB_LAUNCH(kern, dim3(p.nchunks, p.nsplit), dim3(256), smem, stream, a);
            prof_end(4, stream);
            return GPB_OK;
        }
    }
    if constexpr (!BWD && sizeof(T) == 8 && DOC <= 4 && Q <= 4) {   // fp64 forward: two-CTA register budget
        auto kern = gpb::mm_pairs_fwd64_kernel<Q, DOC>;
        if (mm_pairs_smem(kern, smem)) return GPB_ERR_CUDA;
        GPB_LAUNCH(kern, dim3(p.nchunks, p.nsplit), dim3(256), smem, stream, a);
    } else {
        auto kern = gpb::mm_pairs_kernel<T, Q, DOC, BWD, false>;
        if (mm_pairs_smem(kern, smem)) return GPB_ERR_CUDA;
        GPB_LAUNCH(kern, dim3(p.nchunks, p.nsplit), dim3(256), smem, stream, a);
    }
    prof_end(BWD ? 4 : 3, stream);
    return GPB_OK;
}
template <typename T, int Q, bool BWD>
int mm_pairs_doc(const MMPlan& p, const gpb::MMArgs<T>& a, void* stream) {
    switch (p.DOC) {
        case 1: return mm_pairs_launch<T, Q, 1, BWD>(p, a, stream);
        case 2: return mm_pairs_launch<T, Q, 2, BWD>(p, a, stream);
        case 4: return mm_pairs_launch<T, Q, 4, BWD>(p, a, stream);
        case 8: return mm_pairs_launch<T, Q, 8, BWD>(p, a, stream);
        case 16: return mm_pairs_launch<T, Q, 16, BWD>(p, a, stream);
        case 32: return mm_pairs_launch<T, Q, 32, BWD>(p, a, stream);
    }
    return fail(GPB_ERR_ARG, "mm: bad DOC %d", p.DOC);
}
template <typename T, bool BWD>
int mm_pairs_dispatch(const MMPlan& p, const gpb::MMArgs<T>& a, void* stream) {
    switch (p.Qt) {
        case 1: return mm_pairs_doc<T, 1, BWD>(p, a, stream);
        case 2: return mm_pairs_doc<T, 2, BWD>(p, a, stream);
        case 3: return mm_pairs_doc<T, 3, BWD>(p, a, stream);
        case 4: return mm_pairs_doc<T, 4, BWD>(p, a, stream);
        case 5: return mm_pairs_doc<T, 5, BWD>(p, a, stream);
        case 6: return mm_pairs_doc<T, 6, BWD>(p, a, stream);
        case 8: return mm_pairs_doc<T, 8, BWD>(p, a, stream);
        case 16: return mm_pairs_doc<T, 16, BWD>(p, a, stream);
    }
    return fail(GPB_ERR_ARG, "mm: input dim template %d unsupported", p.Qt);
}

template <typename T, int QT>
int mm_psi1_fwd_launch(const MMPlan& p, const double* mx, const double* vx, const double* z,
                       const double* ls, const double* sf, const double* A, const double* vacc, int n,
                       int M, int Q, int Do, double* mout, double* vout, double* psi1save, void* stream) {
    auto kern = gpb::mm_psi1_fwd_kernel<T, QT>;
    size_t smem = sizeof(double) * ((size_t)32 * (M + 1) + 32 + 1024 + QT) + sizeof(T) * 2 * 32 * QT;
    int rc = allow_smem(kern, smem);
    if (rc) return rc;
    prof_begin(7, stream);
    GPB_LAUNCH(kern, dim3(p.fwd_grid), dim3(256), smem, stream, mx, vx, z, ls, sf, A, vacc, n, M, Q, Do,
               mout, vout, psi1save);
    prof_end(7, stream);
    return GPB_CHECK_LAUNCH();
}
template <int QT>
int mm_rows_bwd_launch(const MMPlan& p, const double* mx, const double* vx, const double* z,
                       const double* ls, const double* sf, const double* A, const double* dm,
                       const double* dv, const double* mout, const double* vacc, const double* rowacc,
                       const double* psi1, int n, int M, int Q, int Do, double* dmx, double* dvx,
                       double* rowpart, void* stream) {
    auto kern = gpb::mm_rows_bwd_kernel<QT>;
    size_t smem = sizeof(double) * ((size_t)QT * M + (size_t)Do * M + 8 * (size_t)Do + 8 * (QT + 2));
    int rc = allow_smem(kern, smem);
    if (rc) return rc;
    prof_begin(5, stream);
    GPB_LAUNCH(kern, dim3(p.rows_grid), dim3(256), smem, stream, mx, vx, z, ls, sf, A, dm, dv, mout, vacc,
               rowacc, psi1, n, M, Q, Do, dmx, dvx, rowpart);
    prof_end(5, stream);
    return GPB_CHECK_LAUNCH();
}
#define GPB_QT_SWITCH(CALL)                                   \
    switch (p.Qt) {                                           \
        case 1: return CALL(1);                               \
        case 2: return CALL(2);                               \
        case 3: return CALL(3);                               \
        case 4: return CALL(4);                               \
        case 5: return CALL(5);                               \
        case 6: return CALL(6);                               \
        case 8: return CALL(8);                               \
        case 16: return CALL(16);                             \
    }                                                         \
    return fail(GPB_ERR_ARG, "mm: input dim template %d unsupported", p.Qt)
template <int QT>
int mm_cols_bwd_launch(const MMPlan& p, const double* mx, const double* vx, const double* z,
                       const double* ls, const double* A, const double* dm, const double* dv,
                       const double* mout, const double* psi1, int n, int M, int Q, int Do,
                       double* colpart, void* stream) {
    const size_t smem = sizeof(double) * 32 * (2 * (size_t)QT + Do);
    // output dims per pass: every pass re-streams psi1[n,M], so wide layers take 16 at a time
    // (Do = 50: 4 passes instead of 13)
    const int DOB = Do == 1 ? 1 : (Do == 2 ? 2 : (Do <= 4 ? 4 : (Do <= 8 ? 8 : 16)));
    prof_begin(6, stream);
#define GPB_COLS(DOBV)                                                                                          \
    {                                                                                                           \
        auto kern = gpb::mm_cols_bwd_kernel<QT, DOBV>;                                                          \
        GPB_LAUNCH(kern, dim3(p.cols_grid), dim3(256), smem, stream, mx, vx, z, ls, A, dm, dv, mout, psi1, n, M, \
                   Q, Do, d0, p.cols_rows_per_block, colpart);                                                  \
    }
    for (int d0 = 0; d0 < Do; d0 += DOB) {
        if (DOB == 1) GPB_COLS(1)
        else if (DOB == 2) GPB_COLS(2)
        else if (DOB == 4) GPB_COLS(4)
        else if (DOB == 8) GPB_COLS(8)
        else GPB_COLS(16)
    }
#undef GPB_COLS
    prof_end(6, stream);
    return GPB_CHECK_LAUNCH();
}
int mm_cols_bwd_dispatch(const MMPlan& p, const double* mx, const double* vx, const double* z,
                         const double* ls, const double* A, const double* dm, const double* dv,
                         const double* mout, const double* psi1, int n, int M, int Q, int Do,
                         double* colpart, void* stream) {
#define GPB_CALL(QT) mm_cols_bwd_launch<QT>(p, mx, vx, z, ls, A, dm, dv, mout, psi1, n, M, Q, Do, colpart, stream)
    GPB_QT_SWITCH(GPB_CALL);
#undef GPB_CALL
}
template <typename T>
int mm_psi1_fwd_dispatch(const MMPlan& p, const double* mx, const double* vx, const double* z,
                         const double* ls, const double* sf, const double* A, const double* vacc, int n,
                         int M, int Q, int Do, double* mout, double* vout, double* psi1save, void* stream) {
#define GPB_CALL(QT) mm_psi1_fwd_launch<T, QT>(p, mx, vx, z, ls, sf, A, vacc, n, M, Q, Do, mout, vout, psi1save, stream)
    GPB_QT_SWITCH(GPB_CALL);
#undef GPB_CALL
}
int mm_rows_bwd_dispatch(const MMPlan& p, const double* mx, const double* vx, const double* z,
                         const double* ls, const double* sf, const double* A, const double* dm,
                         const double* dv, const double* mout, const double* vacc, const double* rowacc,
                         const double* psi1, int n, int M, int Q, int Do, double* dmx, double* dvx,
                         double* rowpart, void* stream) {
#define GPB_CALL(QT) mm_rows_bwd_launch<QT>(p, mx, vx, z, ls, sf, A, dm, dv, mout, vacc, rowacc, psi1, n, M, Q, Do, dmx, dvx, rowpart, stream)
    GPB_QT_SWITCH(GPB_CALL);
#undef GPB_CALL
}

// wide-output forward (Do > 4): row-owner kernel, one launch
template <typename T, int QT, int DOW>
int mm_fwd_wide_launch(const MMPlan& p, gpb::MMArgs<T> a, int n, void* stream) {
    typedef gpb::MMWideCfg<T, QT, DOW> C;
    auto kern = gpb::mm_fwd_wide_kernel<T, QT, DOW>;
    int rc = allow_smem(kern, C::smem_bytes + 1024);
    if (rc) return rc;
    const int nchunks = (int)cdiv(p.P, C::PCW);
    int nsplit = (int)((long)6 * sm_count() / nchunks);
    if (nsplit < 1) nsplit = 1;
    long rps = cdiv(n, nsplit);
    rps = cdiv(rps, 256) * 256;
    a.rows_per_split = (int)rps;
    nsplit = (int)cdiv(n, rps);
    prof_begin(3, stream);
    GPB_LAUNCH(kern, dim3(nchunks, nsplit), dim3(256), C::smem_bytes, stream, a);
    prof_end(3, stream);
    return GPB_CHECK_LAUNCH();
}
template <typename T, int QT>
int mm_fwd_wide_dow(const MMPlan& p, const gpb::MMArgs<T>& a, int n, int Do, void* stream) {
    if (Do <= 8) return mm_fwd_wide_launch<T, QT, 8>(p, a, n, stream);
    if (Do <= 16) return mm_fwd_wide_launch<T, QT, 16>(p, a, n, stream);
    if (Do <= 32) return mm_fwd_wide_launch<T, QT, 32>(p, a, n, stream);
    return mm_fwd_wide_launch<T, QT, 64>(p, a, n, stream);
}
template <typename T>
int mm_fwd_wide_dispatch(const MMPlan& p, const gpb::MMArgs<T>& a, int n, int Do, void* stream) {
#define GPB_CALL(QT) mm_fwd_wide_dow<T, QT>(p, a, n, Do, stream)
    GPB_QT_SWITCH(GPB_CALL);
#undef GPB_CALL
}

template <int QT>
int mm_bwd_wide_mma_launch(const MMPlan& p, gpb::MMArgs<double> a, void* stream) {
    if constexpr (QT <= 8) {
        typedef gpb::MMWideMma<QT> C;
        constexpr int NW = GPB_MM_WIDE_WARPS;
        auto kern = gpb::mm_bwd_wide_mma_kernel<QT, NW>;
        if (mm_pairs_smem(kern, C::smem_bytes)) return GPB_ERR_CUDA;
        a.rows_per_split = p.w_rows_per_split;
        const int DOP8 = (a.Do + 7) / 8 * 8;
        prof_begin(4, stream);
        GPB_LAUNCH(kern, dim3(p.w_nchunks, p.w_nsplit), dim3(NW * 32), C::smem_bytes, stream, a, DOP8);
        prof_end(4, stream);
        return GPB_OK;
    } else {
        (void)p; (void)a; (void)stream;
        return fail(GPB_ERR_ARG, "mm_bwd_wide_mma: Q template %d unsupported", QT);
    }
}
int mm_bwd_wide_mma_dispatch(const MMPlan& p, const gpb::MMArgs<double>& a, void* stream) {
#define GPB_CALL(QT) mm_bwd_wide_mma_launch<QT>(p, a, stream)
    GPB_QT_SWITCH(GPB_CALL);
#undef GPB_CALL
}

template <int QT>
int mm_fwd_wide_mma_launch(const MMPlan& p, const gpb::MMArgs<double>& a, void* stream) {
    if constexpr (QT <= 8) {
        typedef gpb::MMFwdWideMma<QT> C;
        auto kern = gpb::mm_fwd_wide_mma_kernel<QT>;
        if (mm_pairs_smem(kern, C::smem_bytes)) return GPB_ERR_CUDA;
        const int DOP8 = (a.Do + 7) / 8 * 8;
        const int nrb = (int)cdiv(a.n, C::TRF);
        const int grid = nrb < sm_count() ? nrb : sm_count();     // one persistent CTA per SM
        prof_begin(3, stream);
        GPB_LAUNCH(kern, dim3(grid), dim3(256), C::smem_bytes, stream, a, DOP8, (int)(p.PP / C::PCW));
        prof_end(3, stream);
        return GPB_OK;
    } else {
        (void)p; (void)a; (void)stream;
        return fail(GPB_ERR_ARG, "mm_fwd_wide_mma: Q template %d unsupported", QT);
    }
}
int mm_fwd_wide_mma_dispatch(const MMPlan& p, const gpb::MMArgs<double>& a, void* stream) {
#define GPB_CALL(QT) mm_fwd_wide_mma_launch<QT>(p, a, stream)
    GPB_QT_SWITCH(GPB_CALL);
#undef GPB_CALL
}

// fp32 forward of narrow layers with the exponent GEMM on tcgen05 (gpb_umma.cuh)
#ifndef GPB_CPU_EMU
template <int KS, int DN>
int mm_pairs_tc_pass(const gpb::MMTcArgs& t, int grid, void* stream) {
    typedef gpb::MMTcCfg<KS> C;
    auto kern = gpb::mm_pairs_tc_kernel<KS, DN>;
    // One CTA per SM, enforced through the shared-memory request: a CTA owns all 512 TMEM columns, so a second
    // resident CTA would sit in tcgen05.alloc until the first one has walked ALL its tiles (measured: the kernel
    // took 1.5x as long whenever a side-stream kernel delayed the placement of a few CTAs).
    const size_t smem = C::smem_bytes > 120 * 1024 ? C::smem_bytes : 120 * 1024;
    int rc = allow_smem(kern, smem);
    if (rc) return rc;
    GPB_LAUNCH(kern, dim3(grid), dim3(288), smem, stream, t);
    return GPB_OK;
}
template <int KS>
int mm_pairs_tc_launch(const MMPlan& p, const gpb::MMArgs<float>& a, float* gu, void* stream) {
    auto prep = gpb::mm_tc_prep_kernel;
    GPB_LAUNCH(prep, dim3(elementwise_grid(p.PP * 8 * KS)), dim3(256), 0, stream, a.zh, p.PP, a.Qa, KS, gu);
    const int ntiles = (int)cdiv(a.n, 128);
    const int grid = ntiles < sm_count() ? ntiles : sm_count();
    gpb::MMTcArgs t;
    t.mx = a.mx; t.vx = a.vx; t.ls = a.ls; t.Gu = gu; t.bs = a.bs; t.n = a.n; t.Q = a.Qa; t.Do = a.Do; t.PP = p.PP;
    t.rowacc = a.rowacc;
    prof_begin(3, stream);
    int rc = GPB_OK;
    for (int d0 = 0; d0 < a.Do && !rc; d0 += 4) {
        t.d0 = d0;
        t.dn = (a.Do - d0) < 4 ? (a.Do - d0) : 4;
        switch (t.dn) {
            case 1: rc = mm_pairs_tc_pass<KS, 1>(t, grid, stream); break;
            case 2: rc = mm_pairs_tc_pass<KS, 2>(t, grid, stream); break;
            case 3: rc = mm_pairs_tc_pass<KS, 3>(t, grid, stream); break;
            default: rc = mm_pairs_tc_pass<KS, 4>(t, grid, stream); break;
        }
    }
    prof_end(3, stream);
    if (rc) return rc;
    return GPB_CHECK_LAUNCH();
}
#endif

template <typename T>
int mm_check(int n, int M, int Q, int Do) {
    if (n < 1 || M < 1 || Q < 1 || Do < 1) return fail(GPB_ERR_ARG, "mm: empty problem");
    if (q_template(Q) < 0) return fail(GPB_ERR_ARG, "mm: Q=%d unsupported (max 16)", Q);
    if (Do > 64) return fail(GPB_ERR_ARG, "mm: Do=%d unsupported (max 64)", Do);
    return GPB_OK;
}

template <typename T>
int mm_fwd_t(const double* mx, const double* vx, const double* z, const double* ls, const double* sf,
             const double* A, const double* B, int n, int M, int Q, int Do, double* mout,
             double* vout, double* vacc, double* psi1save, void* ws, size_t ws_bytes, void* stream) {
    int rc = mm_check<T>(n, M, Q, Do);
    if (rc) return rc;
    MMPlan p = mm_plan((int)sizeof(T), n, M, Q, Do, 0);
    MMWs<T> w = mm_carve<T>(p, n, M, Q, Do, 0, ws, ws_bytes);
    if (w.bytes > ws_bytes) return fail(GPB_ERR_WS, "mm_fwd: workspace %zu < %zu", ws_bytes, w.bytes);
    auto tab = gpb::mm_pair_table_kernel<T>;
    GPB_LAUNCH(tab, dim3(elementwise_grid(p.PP)), dim3(256), 0, stream, z, ls, sf, B, M, Q, p.Qt, Do,
               p.P, p.PP, w.zh, w.ep, w.bs);
    w.rowacc = vacc;
    dev_memset(w.rowacc, sizeof(double) * (size_t)n * Do, stream);
    gpb::MMArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.mx = mx; a.vx = vx; a.ls = ls; a.zh = w.zh; a.ep = w.ep; a.bs = w.bs; a.dv = nullptr;
    a.n = n; a.Qa = Q; a.Do = Do; a.PP = p.PP; a.rows_per_split = p.rows_per_split;
    a.rowacc = w.rowacc; a.pairpart = nullptr; a.full_coef = 0; a.lam_pass = 0;
    if (Do > 4) {      // wide layers: psi2 evaluated once per row and pair
        a.d0 = 0;
        bool done = false;
        if constexpr (sizeof(T) == 8) {
            if (p.Qt <= 8) {    // fp64: contraction on the FP64 tensor cores
                rc = mm_fwd_wide_mma_dispatch(p, a, stream);
                if (rc) return rc;
                done = true;
            }
        }
        if (!done) {            // fp32 / Q = 16: SIMT row-owner kernel
            rc = mm_fwd_wide_dispatch<T>(p, a, n, Do, stream);
            if (rc) return rc;
        }
    } else {
        bool done = false;
        if constexpr (sizeof(T) == 8) {
            if (p.xpath) {      // exponent on the FP64 tensor cores
                a.d0 = 0;
                rc = mm_pairsx_dispatch<false>(p, a, w.rowfeat, stream);
                if (rc) return rc;
                done = true;
            }
        }
#ifndef GPB_CPU_EMU
        if constexpr (sizeof(T) == 4) {
            if (p.tc) {         // exponent on the 5th-generation tensor cores
                a.d0 = 0;
                rc = p.tc_ks == 1 ? mm_pairs_tc_launch<1>(p, a, w.gu, stream) : mm_pairs_tc_launch<2>(p, a, w.gu, stream);
                if (rc) return rc;
                done = true;
            }
        }
#endif
        for (int pass = 0; pass < (done ? 0 : p.npass); pass++) {
            a.d0 = pass * p.DOC;
            rc = mm_pairs_dispatch<T, false>(p, a, stream);
            if (rc) return rc;
        }
    }
    rc = GPB_CHECK_LAUNCH();
    if (rc) return rc;
    if (M > 512) return fail(GPB_ERR_ARG, "mm: M=%d unsupported (max 512)", M);
    return mm_psi1_fwd_dispatch<T>(p, mx, vx, z, ls, sf, A, w.rowacc, n, M, Q, Do, mout, vout, psi1save, stream);
}

template <typename T>
int mm_bwd_t(const double* mx, const double* vx, const double* z, const double* ls, const double* sf,
             const double* A, const double* B, const double* dm, const double* dv,
             const double* mout, const double* vacc, const double* psi1, int n, int M, int Q, int Do, double* dA,
             double* dB, double* dzu, double* dl, double* dsf2, double* dvsum, double* dmx,
             double* dvx, void* ws, size_t ws_bytes, void* stream) {
    int rc = mm_check<T>(n, M, Q, Do);
    if (rc) return rc;
    MMPlan p = mm_plan((int)sizeof(T), n, M, Q, Do, 1);
    MMWs<T> w = mm_carve<T>(p, n, M, Q, Do, 1, ws, ws_bytes);
    if (w.bytes > ws_bytes) return fail(GPB_ERR_WS, "mm_bwd: workspace %zu < %zu", ws_bytes, w.bytes);
    auto tab = gpb::mm_pair_table_kernel<T>;
    GPB_LAUNCH(tab, dim3(elementwise_grid(p.PP)), dim3(256), 0, stream, z, ls, sf, B, M, Q, p.Qt, Do,
               p.P, p.PP, w.zh, w.ep, w.bs);
    const int NS = 2 * p.Qt;
    dev_memset(w.rowacc, sizeof(double) * (size_t)n * NS, stream);
    gpb::MMArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.mx = mx; a.vx = vx; a.ls = ls; a.zh = w.zh; a.ep = w.ep; a.bs = w.bs; a.dv = dv;
    a.n = n; a.Qa = Q; a.Do = Do; a.PP = p.PP; a.rows_per_split = p.rows_per_split;
    a.rowacc = w.rowacc; a.pairpart = w.pairpart;
    a.full_coef = p.npass > 1 ? 1 : 0;
    const long recstride = (long)(p.DOC + 1 + p.Qt) * p.PP;
    bool wide_done = false;
    if constexpr (sizeof(T) == 8) {
        if (p.wide_mma) {   // wide fp64 layers: one tensor-core kernel, no d-passes
            rc = mm_bwd_wide_mma_dispatch(p, a, stream);
            if (rc) return rc;
            const long wlen = (long)(Do + 1 + p.Qt) * p.PP;
            launch_reduce_partials(w.pairpart, p.w_nsplit, wlen, wlen, w.pairsum, 0, stream);
            wide_done = true;
        }
    }
    if constexpr (sizeof(T) == 8) {
        if (p.xpath && !wide_done) {    // narrow fp64 layers: exponent on the FP64 tensor cores, one pass
            a.d0 = 0;
            a.lam_pass = 1;
            rc = mm_pairsx_dispatch<true>(p, a, w.rowfeat, stream);
            if (rc) return rc;
            const long xlen = (long)(p.DOC + 1 + p.Qt) * p.PP;
            launch_reduce_partials(w.pairpart, p.x_nsplit, xlen, (long)(Do < p.DOC ? Do : p.DOC) * p.PP, w.pairsum, 0, stream);
            launch_reduce_partials(w.pairpart + (long)p.DOC * p.PP, p.x_nsplit, xlen, (long)(1 + p.Qt) * p.PP,
                                   w.pairsum + (long)Do * p.PP, 0, stream);
            wide_done = true;
        }
    }
    for (int pass = 0; pass < (wide_done ? 0 : p.npass); pass++) {
        a.d0 = pass * p.DOC;
        a.lam_pass = pass == 0 ? 1 : 0;
        rc = mm_pairs_dispatch<T, true>(p, a, stream);
        if (rc) return rc;
        // fold the row splits: dBp rows of this d-chunk, and (first pass) S0 | S1
        int nd = (Do - a.d0) < p.DOC ? (Do - a.d0) : p.DOC;
        launch_reduce_partials(w.pairpart, p.nsplit, recstride, (long)nd * p.PP, w.pairsum + (long)a.d0 * p.PP, 0, stream);
        if (pass == 0)
            launch_reduce_partials(w.pairpart + (long)p.DOC * p.PP, p.nsplit, recstride, (long)(1 + p.Qt) * p.PP, w.pairsum + (long)Do * p.PP, 0, stream);
    }
    rc = GPB_CHECK_LAUNCH();
    if (rc) return rc;
    // row-wise epilogue: dmx, dvx + row-summed hyper terms
    rc = mm_rows_bwd_dispatch(p, mx, vx, z, ls, sf, A, dm, dv, mout, vacc, w.rowacc, psi1, n, M, Q, Do, dmx,
                              dvx, w.rowpart, stream);
    if (rc) return rc;
    launch_reduce_partials(w.rowpart, p.rows_grid, (long)(2 + Q), (long)(2 + Q), w.rowsum, 0, stream);
    {   // column-wise psi1 part: dA, dZ1
        rc = mm_cols_bwd_dispatch(p, mx, vx, z, ls, A, dm, dv, mout, psi1, n, M, Q, Do, w.colpart, stream);
        if (rc) return rc;
        long len = (long)Do * M + (long)M * Q;
        launch_reduce_partials(w.colpart, p.cols_grid, len, len, w.colsum, 0, stream);
    }
    {
        auto kern = gpb::mm_pair_finish_kernel;
        GPB_LAUNCH(kern, dim3(elementwise_grid((long)Do * M * M)), dim3(256), 0, stream, w.pairsum, p.DOC,
                   Do, z, ls, M, Q, p.PP, dB, w.dZ2, w.dlW);
        auto fin = gpb::mm_final_kernel;
        GPB_LAUNCH(fin, dim3(1), dim3(256), 0, stream, w.colsum, w.rowsum, w.dZ2, w.dlW, ls, M, Q, Do, dA,
                   dzu, dl, dsf2, dvsum);
    }
    return GPB_CHECK_LAUNCH();
}

}  // namespace

extern "C" {

size_t gpb_mm_ws_bytes(int n, int M, int Q, int Do, int backward) {
    if (q_template(Q) < 0 || n < 1 || M < 1 || Do < 1) return 0;
    // the pair padding (and with it the fp64 partial records) depends on the precision: take the max
    MMPlan p8 = mm_plan(8, n, M, Q, Do, backward), p4 = mm_plan(4, n, M, Q, Do, backward);
    size_t b8 = mm_carve<double>(p8, n, M, Q, Do, backward, nullptr, 0).bytes;
    size_t b4 = mm_carve<float>(p4, n, M, Q, Do, backward, nullptr, 0).bytes;
    return b8 > b4 ? b8 : b4;
}

int gpb_mm_fwd(int prec, const double* mx, const double* vx, const double* z, const double* ls,
               const double* sf, const double* A, const double* B, int n, int M, int Q, int Do,
               double* mout, double* vout, double* vacc, double* psi1save, void* ws, size_t ws_bytes,
               void* stream) {
    if (!mx || !vx || !z || !ls || !sf || !A || !B || !mout || !vout || !vacc || !ws) return fail(GPB_ERR_ARG, "mm_fwd: null pointer");
    if (prec == GPB_F64) return mm_fwd_t<double>(mx, vx, z, ls, sf, A, B, n, M, Q, Do, mout, vout, vacc, psi1save, ws, ws_bytes, stream);
    return mm_fwd_t<float>(mx, vx, z, ls, sf, A, B, n, M, Q, Do, mout, vout, vacc, psi1save, ws, ws_bytes, stream);
}

int gpb_mm_bwd(int prec, const double* mx, const double* vx, const double* z, const double* ls,
               const double* sf, const double* A, const double* B, const double* dm, const double* dv,
               const double* mout, const double* vacc, const double* psi1, int n, int M, int Q, int Do, double* dA,
               double* dB, double* dzu, double* dl, double* dsf2, double* dvsum, double* dmx,
               double* dvx, void* ws, size_t ws_bytes, void* stream) {
    if (!mx || !vx || !z || !ls || !sf || !A || !B || !dm || !dv || !mout || !vacc || !psi1 || !dA || !dB || !dzu || !dl ||
        !dsf2 || !dvsum || !dmx || !dvx || !ws)
        return fail(GPB_ERR_ARG, "mm_bwd: null pointer");
    if (prec == GPB_F64)
        return mm_bwd_t<double>(mx, vx, z, ls, sf, A, B, dm, dv, mout, vacc, psi1, n, M, Q, Do, dA, dB, dzu, dl, dsf2, dvsum, dmx, dvx, ws, ws_bytes, stream);
    return mm_bwd_t<float>(mx, vx, z, ls, sf, A, B, dm, dv, mout, vacc, psi1, n, M, Q, Do, dA, dB, dzu, dl, dsf2, dvsum, dmx, dvx, ws, ws_bytes, stream);
}

}  // extern "C"
