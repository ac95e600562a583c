// gpb_capi_det.cu -- deterministic-input layer entry points (a5, a8).
#include "gpb_common.cuh"
#include "gpb_umma.cuh"

namespace {

// ------------------------------- deterministic layer ------------------------------------
struct DetBwdPlan {
    int MP, CWB, RY, gx, gy, rows_per_block, G;
    long rec_len;
};
DetBwdPlan det_bwd_plan(int n, int M, int D, int Do) {
    DetBwdPlan p;
    p.MP = gpb_det_pad_m(M);
    p.CWB = p.MP < 256 ? p.MP : 256;
    p.RY = 256 / p.CWB;
    p.gy = p.MP / p.CWB;
    int want = 2 * sm_count() / p.gy;
    if (want < 1) want = 1;
    int rpb = (int)cdiv(n, want);
    if (rpb < 64) rpb = 64;
    p.rows_per_block = rpb;
    p.gx = (int)cdiv(n, rpb);
    p.G = p.gx * p.RY;
    p.rec_len = (long)p.MP + 2L * p.MP * D + (long)Do * p.MP;
    return p;
}

struct SyrkPlan {
    int MP, nb, nbu, nsplit, rows_per_split;
};
SyrkPlan syrk_plan(int n, int M, int Do) {
    SyrkPlan p;
    p.MP = gpb_det_pad_m(M);
    p.nb = p.MP / 128;
    p.nbu = p.nb * (p.nb + 1) / 2;
    // row splits: an integer number of waves (1 block/SM), up to GPB_SYRK_WAVES, at least 256 rows per
    // block.  Every split writes (and the finish kernel re-reads) a 128 x 128 partial block per output
    // block: at the north-star shape 6 waves cost 230 MB of partial traffic (0.16 ms of a 12 ms step).
    int best = 1;
    for (int waves = GPB_SYRK_WAVES; waves >= 1; waves--) {
        int ns = (int)((long)waves * sm_count() / ((long)p.nbu * Do));
        if (ns >= 1 && cdiv(n, ns) >= 256) { best = ns; break; }
    }
    int rps = (int)cdiv(n, best);
    rps = (int)(cdiv(rps, 16) * 16);
    if (rps < 16) rps = 16;
    p.rows_per_split = rps;
    p.nsplit = (int)cdiv(n, rps);
    return p;
}

template <typename T, int MP>
int det_fwd_launch(const double* x, const double* z, const double* ls, const double* sf,
                   const void* Ap, const void* Bp, int n, int M, int D, int Do, double* mout,
                   double* vout, void* Ksave, void* Tsave, void* stream) {
    typedef gpb::DetCfg<T, MP> C;
    gpb::DetFwdArgs<T> a;
    a.x = x; a.z = z; a.ls = ls; a.sf = sf;
    a.Ap = (const T*)Ap; a.Bp = (const T*)Bp;
    a.n = n; a.M = M; a.D = D; a.Do = Do;
    a.mout = mout; a.vout = vout; a.Ksave = (T*)Ksave; a.Tsave = (T*)Tsave;
    auto kern = gpb::det_fwd_kernel<T, MP>;
    int rc = allow_smem(kern, C::smem_bytes);
    if (rc) return rc;
    int ntiles = (int)cdiv(n, C::TN);
    int grid = ntiles < sm_count() ? ntiles : sm_count();
    prof_begin(0, stream);
    GPB_LAUNCH(kern, dim3(grid), dim3(256), C::smem_bytes, stream, a);
    prof_end(0, stream);
    return GPB_CHECK_LAUNCH();
}

template <typename T>
int det_fwd_t(const double* x, const double* z, const double* ls, const double* sf, const void* Ap,
              const void* Bp, int n, int M, int D, int Do, double* mout, double* vout, void* Ksave,
              void* Tsave, void* stream) {
    switch (gpb_det_pad_m(M)) {
        case 128: return det_fwd_launch<T, 128>(x, z, ls, sf, Ap, Bp, n, M, D, Do, mout, vout, Ksave, Tsave, stream);
        case 256: return det_fwd_launch<T, 256>(x, z, ls, sf, Ap, Bp, n, M, D, Do, mout, vout, Ksave, Tsave, stream);
        case 512: return det_fwd_launch<T, 512>(x, z, ls, sf, Ap, Bp, n, M, D, Do, mout, vout, Ksave, Tsave, stream);
    }
    return fail(GPB_ERR_ARG, "det_fwd: M=%d unsupported (max 512)", M);
}

// fp32-psi mode on tcgen05 (gpb_umma.cuh): 3xTF32 Kfu . B_d with TMEM accumulators
#ifndef GPB_CPU_EMU
template <int MP, int DP>
int det_fwd_umma_launch(const gpb::DetUmmaArgs& a, void* stream) {
    typedef gpb::DetUmmaCfg<MP> C;
    auto kern = gpb::det_fwd_umma_kernel<MP, DP>;
    int rc = allow_smem(kern, C::smem_bytes(DP));
    if (rc) return rc;
    const int ntiles = (int)cdiv(a.n, 128);
    const int grid = ntiles < sm_count() ? ntiles : sm_count();
    prof_begin(0, stream);
    GPB_LAUNCH(kern, dim3(grid), dim3(256), C::smem_bytes(DP), stream, a);
    prof_end(0, stream);
    return GPB_CHECK_LAUNCH();
}
template <int MP>
int det_fwd_umma_dp(const gpb::DetUmmaArgs& a, void* stream) {
    if (a.D <= 4) return det_fwd_umma_launch<MP, 4>(a, stream);
    if (a.D <= 8) return det_fwd_umma_launch<MP, 8>(a, stream);
    if (a.D <= 16) return det_fwd_umma_launch<MP, 16>(a, stream);
    return det_fwd_umma_launch<MP, 32>(a, stream);
}
#endif
int det_tc_dp(int D) { return D <= 4 ? 4 : (D <= 8 ? 8 : (D <= 16 ? 16 : 32)); }

// fp64: tensor-core (DMMA) kernel
template <int MP>
int det_fwd_mma_launch(const double* x, const double* z, const double* ls, const double* sf,
                       const void* Ap, const void* Bp, int n, int M, int D, int Do, double* mout,
                       double* vout, void* Ksave, void* Tsave, void* stream) {
    typedef gpb::DetMmaCfg<MP> C;
    gpb::DetFwdArgs<double> a;
    a.x = x; a.z = z; a.ls = ls; a.sf = sf;
    a.Ap = (const double*)Ap; a.Bp = (const double*)Bp;
    a.n = n; a.M = M; a.D = D; a.Do = Do;
    a.mout = mout; a.vout = vout; a.Ksave = (double*)Ksave; a.Tsave = (double*)Tsave;
    auto kern = gpb::det_fwd_mma_kernel<MP>;
    int rc = allow_smem(kern, C::smem_bytes);
    if (rc) return rc;
    int ntiles = (int)cdiv(n, C::TN);
    int grid = ntiles < sm_count() ? ntiles : sm_count();
    prof_begin(0, stream);
    GPB_LAUNCH(kern, dim3(grid), dim3(256), C::smem_bytes, stream, a);
    prof_end(0, stream);
    return GPB_CHECK_LAUNCH();
}
template <>
int det_fwd_t<double>(const double* x, const double* z, const double* ls, const double* sf, const void* Ap,
                      const void* Bp, int n, int M, int D, int Do, double* mout, double* vout, void* Ksave,
                      void* Tsave, void* stream) {
    switch (gpb_det_pad_m(M)) {
        case 128: return det_fwd_mma_launch<128>(x, z, ls, sf, Ap, Bp, n, M, D, Do, mout, vout, Ksave, Tsave, stream);
        case 256: return det_fwd_mma_launch<256>(x, z, ls, sf, Ap, Bp, n, M, D, Do, mout, vout, Ksave, Tsave, stream);
        case 512: return det_fwd_mma_launch<512>(x, z, ls, sf, Ap, Bp, n, M, D, Do, mout, vout, Ksave, Tsave, stream);
    }
    return fail(GPB_ERR_ARG, "det_fwd: M=%d unsupported (max 512)", M);
}

template <typename T>
int det_bwd_t(const double* x, const double* z, const double* ls, const double* sf, const void* Ap,
              const double* dm, const double* dv, const void* Ksave, const void* Tsave, int n, int M,
              int D, int Do, double* dA, double* dzu, double* dl, double* dsf2, void* ws,
              size_t ws_bytes, void* stream) {
    DetBwdPlan p = det_bwd_plan(n, M, D, Do);
    Carver cv(ws, ws_bytes);
    double* part = (double*)cv.take(sizeof(double) * p.G * p.rec_len);
    double* rec = (double*)cv.take(sizeof(double) * p.rec_len);
    if (!cv.ok()) return fail(GPB_ERR_WS, "det_bwd: workspace %zu < %zu", ws_bytes, cv.off);
    dim3 grid(p.gx, p.gy);
    // Do <= 4 and D <= 16: cp.async ring kernel; otherwise the register-batched generic kernel
#define GPB_BWD2(DP, DOB)                                                                       \
    {                                                                                           \
        if (DOB > 0 && D <= 16) {                                                               \
            const size_t smem = gpb::DetBwdRing<T, (DOB > 0 ? DOB : 1)>::smem_bytes;            \
            if (p.CWB == 128) {                                                                 \
                auto kern = gpb::det_bwd_ring_kernel<T, DP, (DOB > 0 ? DOB : 1), 128>;          \
                int rcs = allow_smem(kern, smem);                                               \
                if (rcs) return rcs;                                                            \
                GPB_LAUNCH(kern, grid, dim3(256), smem, stream, x, z, ls, (const T*)Ap, dm, dv, \
                           (const T*)Ksave, (const T*)Tsave, n, M, p.MP, D, Do,                 \
                           p.rows_per_block, part, p.rec_len);                                  \
            } else {                                                                            \
                auto kern = gpb::det_bwd_ring_kernel<T, DP, (DOB > 0 ? DOB : 1), 256>;          \
                int rcs = allow_smem(kern, smem);                                               \
                if (rcs) return rcs;                                                            \
                GPB_LAUNCH(kern, grid, dim3(256), smem, stream, x, z, ls, (const T*)Ap, dm, dv, \
                           (const T*)Ksave, (const T*)Tsave, n, M, p.MP, D, Do,                 \
                           p.rows_per_block, part, p.rec_len);                                  \
            }                                                                                   \
        } else {                                                                                \
            auto kern = gpb::det_bwd_kernel<T, DP, DOB>;                                        \
            GPB_LAUNCH(kern, grid, dim3(256), 0, stream, x, z, ls, (const T*)Ap, dm, dv,        \
                       (const T*)Ksave, (const T*)Tsave, n, M, p.MP, D, Do, p.rows_per_block,   \
                       part, p.rec_len);                                                        \
        }                                                                                       \
    }
#define GPB_BWD(DP)                                                                             \
    {                                                                                           \
        if (Do == 1) GPB_BWD2(DP, 1) else if (Do == 2) GPB_BWD2(DP, 2) else if (Do <= 4)        \
            GPB_BWD2(DP, 4) else GPB_BWD2(DP, 0)                                                \
    }
    prof_begin(1, stream);
    if (D <= 2) GPB_BWD(2) else if (D <= 4) GPB_BWD(4) else if (D <= 6) GPB_BWD(6)
    else if (D <= 8) GPB_BWD(8) else if (D <= 10) GPB_BWD(10) else if (D <= 12) GPB_BWD(12)
    else GPB_BWD(16)
    prof_end(1, stream);
#undef GPB_BWD
#undef GPB_BWD2
    int rc = GPB_CHECK_LAUNCH();
    if (rc) return rc;
    launch_reduce_partials(part, p.G, p.rec_len, p.rec_len, rec, 0, stream);
    auto fin = gpb::det_bwd_finish_kernel;
    GPB_LAUNCH(fin, dim3(1), dim3(256), 0, stream, rec, sf, M, p.MP, D, Do, dA, dzu, dl, dsf2);
    return GPB_CHECK_LAUNCH();
}

// fp32 on tcgen05 (gpb_umma.cuh): MP <= 256; one CTA per SM over (row splits x Dout)
struct SyrkTcPlan {
    int use, MP, nbu, nsplit, rows_per_split;
};
SyrkTcPlan syrk_tc_plan(int n, int M, int Do) {
    SyrkTcPlan p;
    p.MP = gpb_det_pad_m(M);
    p.nbu = p.MP == 256 ? 3 : 1;
    p.use = 0;
#ifndef GPB_CPU_EMU
    p.use = (GPB_DET_SYRK_TC && (p.MP == 128 || p.MP == 256)) ? 1 : 0;
#endif
    int ns = sm_count() / Do;
    if (ns < 1) ns = 1;
    if ((long)ns * 512 > n) ns = (int)cdiv(n, 512);
    int rps = (int)(cdiv(cdiv(n, ns), 16) * 16);
    p.rows_per_split = rps;
    p.nsplit = (int)cdiv(n, rps);
    return p;
}
#ifndef GPB_CPU_EMU
template <int MP>
int det_syrk_umma_launch(const SyrkTcPlan& p, const float* Ksave, const double* dv, int n, int Do, double* part,
                         void* stream) {
    typedef gpb::SyrkUmmaCfg<MP> C;
    auto kern = gpb::det_syrk_umma_kernel<MP>;
    // (one CTA per SM: the kernel allocates all TMEM columns; at MP = 128 the request is padded accordingly)
    const size_t smem = C::smem_bytes > 120 * 1024 ? C::smem_bytes : 120 * 1024;
    int rc = allow_smem(kern, smem);
    if (rc) return rc;
    gpb::SyrkUmmaArgs a;
    a.Ksave = Ksave; a.dv = dv; a.n = n; a.Do = Do; a.rows_per_split = p.rows_per_split; a.part = part;
    prof_begin(2, stream);
    GPB_LAUNCH(kern, dim3(p.nsplit, Do), dim3(256), smem, stream, a);
    prof_end(2, stream);
    return GPB_CHECK_LAUNCH();
}
#endif

template <typename T>
int det_syrk_t(const void* Ksave, const double* dv, int n, int M, int Do, double* dB, void* ws,
               size_t ws_bytes, void* stream) {
#ifndef GPB_CPU_EMU
    if constexpr (sizeof(T) == 4) {
        SyrkTcPlan tp = syrk_tc_plan(n, M, Do);
        if (tp.use) {
            Carver cvt(ws, ws_bytes);
            double* tpart = (double*)cvt.take(sizeof(double) * (size_t)tp.nsplit * Do * tp.nbu * 128 * 128);
            if (!cvt.ok()) return fail(GPB_ERR_WS, "det_syrk: workspace %zu < %zu", ws_bytes, cvt.off);
            int rct = tp.MP == 256 ? det_syrk_umma_launch<256>(tp, (const float*)Ksave, dv, n, Do, tpart, stream)
                                   : det_syrk_umma_launch<128>(tp, (const float*)Ksave, dv, n, Do, tpart, stream);
            if (rct) return rct;
            auto fint = gpb::det_syrk_finish_kernel;
            GPB_LAUNCH(fint, dim3(elementwise_grid((long)Do * M * M)), dim3(256), 0, stream, tpart, tp.nsplit,
                       tp.MP, M, Do, dB, 1);
            return GPB_CHECK_LAUNCH();
        }
    }
#endif
    SyrkPlan p = syrk_plan(n, M, Do);
    Carver cv(ws, ws_bytes);
    double* part = (double*)cv.take(sizeof(double) * (size_t)p.nsplit * Do * p.nbu * 128 * 128);
    if (!cv.ok()) return fail(GPB_ERR_WS, "det_syrk: workspace %zu < %zu", ws_bytes, cv.off);
    auto kern = gpb::det_syrk_kernel<T>;
    int rc = allow_smem(kern, gpb::SyrkCfg<T>::smem_bytes);
    if (rc) return rc;
    prof_begin(2, stream);
    GPB_LAUNCH(kern, dim3(p.nbu, p.nsplit, Do), dim3(256), gpb::SyrkCfg<T>::smem_bytes, stream,
               (const T*)Ksave, dv, n, p.MP, Do, p.rows_per_split, part);
    prof_end(2, stream);
    rc = GPB_CHECK_LAUNCH();
    if (rc) return rc;
    auto fin = gpb::det_syrk_finish_kernel;
    GPB_LAUNCH(fin, dim3(elementwise_grid((long)Do * M * M)), dim3(256), 0, stream, part, p.nsplit,
               p.MP, M, Do, dB, 0);
    return GPB_CHECK_LAUNCH();
}

// fp64: tensor-core (DMMA) rank update
template <>
int det_syrk_t<double>(const void* Ksave, const double* dv, int n, int M, int Do, double* dB, void* ws,
                       size_t ws_bytes, void* stream) {
    SyrkPlan p = syrk_plan(n, M, Do);
    Carver cv(ws, ws_bytes);
    double* part = (double*)cv.take(sizeof(double) * (size_t)p.nsplit * Do * p.nbu * 128 * 128);
    if (!cv.ok()) return fail(GPB_ERR_WS, "det_syrk: workspace %zu < %zu", ws_bytes, cv.off);
    auto kern = gpb::det_syrk_mma_kernel;
    int rc = allow_smem(kern, gpb::SyrkMmaCfg::smem_bytes);
    if (rc) return rc;
    prof_begin(2, stream);
    GPB_LAUNCH(kern, dim3(p.nbu, p.nsplit, Do), dim3(256), gpb::SyrkMmaCfg::smem_bytes, stream,
               (const double*)Ksave, dv, n, p.MP, Do, p.rows_per_split, part);
    prof_end(2, stream);
    rc = GPB_CHECK_LAUNCH();
    if (rc) return rc;
    auto fin = gpb::det_syrk_finish_kernel;
    GPB_LAUNCH(fin, dim3(elementwise_grid((long)Do * M * M)), dim3(256), 0, stream, part, p.nsplit,
               p.MP, M, Do, dB, 0);
    return GPB_CHECK_LAUNCH();
}

}  // namespace

extern "C" {

int gpb_det_pad_m(int M) {
    if (M < 1) return -1;
    if (M <= 128) return 128;
    if (M <= 256) return 256;
    if (M <= 512) return 512;
    return -1;
}

int gpb_det_pad_operands(int prec, const double* A, const double* B, int M, int Do, void* Ap, void* Bp,
                         void* stream) {
    int MP = gpb_det_pad_m(M);
    if (MP < 0 || !A || !B || !Ap || !Bp || Do < 1) return fail(GPB_ERR_ARG, "det_pad_operands: bad argument");
    int grid = elementwise_grid((long)Do * MP * MP);
    if (prec == GPB_F64) {
        auto kern = gpb::det_pad_kernel<double>;
        GPB_LAUNCH(kern, dim3(grid), dim3(256), 0, stream, A, B, M, MP, Do, (double*)Ap, (double*)Bp);
    } else {
        auto kern = gpb::det_pad_kernel<float>;
        GPB_LAUNCH(kern, dim3(grid), dim3(256), 0, stream, A, B, M, MP, Do, (float*)Ap, (float*)Bp);
    }
    return GPB_CHECK_LAUNCH();
}

int gpb_det_fwd(int prec, const double* x, const double* z, const double* ls, const double* sf,
                const void* Ap, const void* Bp, int n, int M, int D, int Do, double* mout, double* vout,
                void* Ksave, void* Tsave, void* stream) {
    if (!x || !z || !ls || !sf || !Ap || !Bp || !mout || !vout || n < 1 || D < 1 || Do < 1)
        return fail(GPB_ERR_ARG, "det_fwd: bad argument");
    if (D > 32) return fail(GPB_ERR_ARG, "det_fwd: D=%d unsupported (max 32)", D);
    if (prec == GPB_F64) return det_fwd_t<double>(x, z, ls, sf, Ap, Bp, n, M, D, Do, mout, vout, Ksave, Tsave, stream);
    return det_fwd_t<float>(x, z, ls, sf, Ap, Bp, n, M, D, Do, mout, vout, Ksave, Tsave, stream);
}

int gpb_det_tc_available(void) {
#ifndef GPB_CPU_EMU
    return 1;
#else
    return 0;
#endif
}
size_t gpb_det_tc_bu_bytes(int M, int Do) {
    const int MP = gpb_det_pad_m(M);
    return MP < 0 ? 0 : sizeof(float) * 2 * (size_t)Do * MP * MP;
}
size_t gpb_det_tc_zs_bytes(int M, int D) {
    const int MP = gpb_det_pad_m(M);
    return (MP < 0 || D < 1 || D > 32) ? 0 : sizeof(float) * (size_t)MP * det_tc_dp(D);
}
int gpb_det_tc_prep(const void* Bp, const double* z, const double* ls, int M, int D, int Do, void* Bu, void* Zs,
                    void* stream) {
#ifndef GPB_CPU_EMU
    const int MP = gpb_det_pad_m(M);
    if (MP < 0 || !Bp || !z || !ls || !Bu || !Zs || D < 1 || D > 32 || Do < 1)
        return fail(GPB_ERR_ARG, "det_tc_prep: bad argument");
    auto kern = gpb::det_umma_prep_kernel;
    GPB_LAUNCH(kern, dim3(elementwise_grid((long)Do * MP * MP)), dim3(256), 0, stream, (const float*)Bp, z, ls, M, MP,
               D, det_tc_dp(D), Do, (float*)Bu, (float*)Zs);
    return GPB_CHECK_LAUNCH();
#else
    (void)Bp; (void)z; (void)ls; (void)M; (void)D; (void)Do; (void)Bu; (void)Zs; (void)stream;
    return fail(GPB_ERR_ARG, "det_tc_prep: tcgen05 path not available in this build");
#endif
}
int gpb_det_fwd_tc(const double* x, const double* ls, const double* sf, const void* Zs, const void* Ap,
                   const void* Bu, int n, int M, int D, int Do, double* mout, double* vout, void* Ksave,
                   void* Tsave, void* stream) {
#ifndef GPB_CPU_EMU
    if (!x || !ls || !sf || !Zs || !Ap || !Bu || !mout || !vout || !Ksave || !Tsave || n < 1 || D < 1 || D > 32 || Do < 1)
        return fail(GPB_ERR_ARG, "det_fwd_tc: bad argument");
    gpb::DetUmmaArgs a;
    a.x = x; a.ls = ls; a.sf = sf; a.Zs = (const float*)Zs; a.Ap = (const float*)Ap; a.Bu = (const float*)Bu;
    a.n = n; a.M = M; a.D = D; a.Do = Do; a.mout = mout; a.vout = vout; a.Ksave = (float*)Ksave; a.Tsave = (float*)Tsave;
    switch (gpb_det_pad_m(M)) {
        case 128: return det_fwd_umma_dp<128>(a, stream);
        case 256: return det_fwd_umma_dp<256>(a, stream);
        case 512: return det_fwd_umma_dp<512>(a, stream);
    }
    return fail(GPB_ERR_ARG, "det_fwd_tc: M=%d unsupported (max 512)", M);
#else
    (void)x; (void)ls; (void)sf; (void)Zs; (void)Ap; (void)Bu; (void)n; (void)M; (void)D; (void)Do; (void)mout;
    (void)vout; (void)Ksave; (void)Tsave; (void)stream;
    return fail(GPB_ERR_ARG, "det_fwd_tc: tcgen05 path not available in this build");
#endif
}

size_t gpb_det_bwd_ws_bytes(int n, int M, int D, int Do) {
    if (gpb_det_pad_m(M) < 0) return 0;
    DetBwdPlan p = det_bwd_plan(n, M, D, Do);
    return align256(sizeof(double) * p.G * p.rec_len) + align256(sizeof(double) * p.rec_len);
}

int gpb_det_bwd(int prec, const double* x, const double* z, const double* ls, const double* sf,
                const void* Ap, const double* dm, const double* dv, const void* Ksave, const void* Tsave,
                int n, int M, int D, int Do, double* dA, double* dzu, double* dl, double* dsf2, void* ws,
                size_t ws_bytes, void* stream) {
    if (!x || !z || !ls || !sf || !Ap || !dm || !dv || !Ksave || !Tsave || !dA || !dzu || !dl || !dsf2 || !ws)
        return fail(GPB_ERR_ARG, "det_bwd: null pointer");
    if (gpb_det_pad_m(M) < 0 || n < 1 || D < 1 || Do < 1) return fail(GPB_ERR_ARG, "det_bwd: bad size");
    if (prec == GPB_F64)
        return det_bwd_t<double>(x, z, ls, sf, Ap, dm, dv, Ksave, Tsave, n, M, D, Do, dA, dzu, dl, dsf2, ws, ws_bytes, stream);
    return det_bwd_t<float>(x, z, ls, sf, Ap, dm, dv, Ksave, Tsave, n, M, D, Do, dA, dzu, dl, dsf2, ws, ws_bytes, stream);
}

int gpb_det_dx(int prec, const double* x, const double* z, const double* ls, const void* Ap, const double* dm,
               const double* dv, const void* Ksave, const void* Tsave, int n, int M, int D, int Do,
               double* dx, void* stream) {
    if (!x || !z || !ls || !Ap || !dm || !dv || !Ksave || !Tsave || !dx)
        return fail(GPB_ERR_ARG, "det_dx: null pointer");
    const int MP = gpb_det_pad_m(M);
    if (MP < 0 || n < 1 || D < 1 || D > 32 || Do < 1) return fail(GPB_ERR_ARG, "det_dx: bad size");
    const size_t smem = sizeof(double) * ((size_t)M * D + D);
    long blocks = ((long)n + 7) / 8;
    const long cap = 16L * gpb_sm_count();
    const int grid = (int)(blocks < cap ? blocks : cap);
    int rc;
    if (prec == GPB_F64) {
        auto kern = gpb::det_dx_kernel<double>;
        if ((rc = allow_smem(kern, smem))) return rc;
        GPB_LAUNCH(kern, dim3(grid), dim3(256), smem, stream, x, z, ls, (const double*)Ap, dm, dv,
                   (const double*)Ksave, (const double*)Tsave, n, M, MP, D, Do, dx);
    } else {
        auto kern = gpb::det_dx_kernel<float>;
        if ((rc = allow_smem(kern, smem))) return rc;
        GPB_LAUNCH(kern, dim3(grid), dim3(256), smem, stream, x, z, ls, (const float*)Ap, dm, dv,
                   (const float*)Ksave, (const float*)Tsave, n, M, MP, D, Do, dx);
    }
    return GPB_CHECK_LAUNCH();
}

size_t gpb_det_syrk_ws_bytes(int n, int M, int Do) {
    if (gpb_det_pad_m(M) < 0) return 0;
    SyrkPlan p = syrk_plan(n, M, Do);
    SyrkTcPlan tp = syrk_tc_plan(n, M, Do);
    const size_t b0 = sizeof(double) * (size_t)p.nsplit * Do * p.nbu * 128 * 128;
    const size_t b1 = tp.use ? sizeof(double) * (size_t)tp.nsplit * Do * tp.nbu * 128 * 128 : 0;
    return align256(b0 > b1 ? b0 : b1);
}

int gpb_det_syrk(int prec, const void* Ksave, const double* dv, int n, int M, int Do, double* dB,
                 void* ws, size_t ws_bytes, void* stream) {
    if (!Ksave || !dv || !dB || !ws || gpb_det_pad_m(M) < 0 || n < 1 || Do < 1)
        return fail(GPB_ERR_ARG, "det_syrk: bad argument");
    if (prec == GPB_F64) return det_syrk_t<double>(Ksave, dv, n, M, Do, dB, ws, ws_bytes, stream);
    return det_syrk_t<float>(Ksave, dv, n, M, Do, dB, ws, ws_bytes, stream);
}

}  // extern "C"
