// gpb_capi_misc.cu -- shared state + small entry points (kmat, psi_stats, gauss_lik, profiling).
#include "gpb_common.cuh"

char g_gpb_err[512] = "";
long g_gpb_launches = 0;
int g_gpb_prof_on = 0;
#ifndef GPB_CPU_EMU
GpbProfPair g_gpb_prof_pending[4096];
int g_gpb_prof_n = 0;
namespace {
double g_prof_ms[8] = {0};
long g_prof_cnt[8] = {0};
}
#endif

// a14: lik_layers.py:573-627.  out = [ sum quad | sum log|Vy| | dRacc[Do] | dC[Do*Q] ] (unscaled sums)
namespace {
int emis_pad(int x) { return x <= 2 ? 2 : (x <= 4 ? 4 : (x <= 8 ? 8 : -1)); }
int emis_grid(int n) {
    long b = cdiv(n, 128);
    long cap = (long)sm_count() * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
template <int DO, int QT>
int emis_launch(const double* mx, const double* vx, const double* y, const double* C, const double* R,
                double alpha, double scale, int n, int Q, int Do, double* dmx, double* dvx, double* out,
                double* part, void* stream) {
    constexpr int NV = 2 + DO * (1 + QT);
    const int grid = emis_grid(n);
    auto kern = gpb::gauss_emis_kernel<DO, QT>;
    GPB_LAUNCH(kern, dim3(grid), dim3(128), 0, stream, mx, vx, y, C, R, alpha, scale, n, Q, Do, dmx, dvx, part);
    // fold the per-block records, then compact the padded record to [2 + Do + Do*Q]
    double* full = part + (size_t)grid * NV;
    launch_reduce_partials((const double*)part, grid, (long)NV, (long)NV, full, 0, stream);
    auto cmp = gpb::gauss_emis_compact_kernel;
    GPB_LAUNCH(cmp, dim3(1), dim3(128), 0, stream, (const double*)full, DO, QT, Do, Q, out);
    return GPB_CHECK_LAUNCH();
}
}  // namespace


extern "C" {

int gpb_version(void) { return 100; }
const char* gpb_last_error(void) { return g_err; }
int gpb_sm_count(void) { return sm_count(); }
long gpb_launch_count(void) { return g_launches; }
int gpb_prec_bytes(int prec) { return prec == GPB_F32 ? 4 : 8; }


int gpb_kmat(const double* x, const double* z, const double* ls, const double* sf, int n, int M,
             int D, double jitter, double* out, void* stream) {
    if (!x || !z || !ls || !sf || !out || n < 1 || M < 1 || D < 1) return fail(GPB_ERR_ARG, "kmat: bad argument");
    auto kern = gpb::kmat_kernel;
    GPB_LAUNCH(kern, dim3(elementwise_grid((long)n * M)), dim3(256), 0, stream, x, z, ls, sf, n, M, D,
               jitter, out);
    return GPB_CHECK_LAUNCH();
}


int gpb_psi_stats(const double* mx, const double* vx, const double* z, const double* ls,
                  const double* sf, int n, int M, int Q, double* psi1, double* psi2, void* stream) {
    if (!mx || !vx || !z || !ls || !sf || !psi1 || !psi2 || n < 1 || M < 1 || Q < 1)
        return fail(GPB_ERR_ARG, "psi_stats: bad argument");
    auto kern = gpb::psi_stats_kernel;
    GPB_LAUNCH(kern, dim3(elementwise_grid((long)n * M * M)), dim3(256), 0, stream, mx, vx, z, ls, sf, n,
               M, Q, psi1, psi2);
    return GPB_CHECK_LAUNCH();
}

size_t gpb_gauss_lik_ws_bytes(long total) { return align256(sizeof(double) * 2 * (size_t)elementwise_grid(total)); }

int gpb_gauss_lik(const double* m, const double* v, const double* y, const double* sn, double alpha,
                  double scale, long total, int mode, double* dm, double* dv, double* out2, void* ws,
                  size_t ws_bytes, void* stream) {
    if (!m || !v || !y || !sn || !dm || !dv || !out2 || total < 1 || (mode != 0 && mode != 1))
        return fail(GPB_ERR_ARG, "gauss_lik: bad argument");
    int grid = elementwise_grid(total);
    if (ws_bytes < sizeof(double) * 2 * (size_t)grid) return fail(GPB_ERR_WS, "gauss_lik: workspace too small");
    auto kern = gpb::gauss_lik_kernel;
    GPB_LAUNCH(kern, dim3(grid), dim3(256), 0, stream, m, v, y, sn, alpha, scale, total, mode, dm, dv,
               (double*)ws);
    launch_reduce_partials((const double*)ws, grid, 2L, 2L, out2, 0, stream);
    return GPB_CHECK_LAUNCH();
}

int gpb_spd_inverse(const double* A, int batch, int M, double* Ainv, double* logdet, void* stream) {
    if (!A || !Ainv || !logdet || batch < 1 || M < 1) return fail(GPB_ERR_ARG, "spd_inverse: bad argument");
    if (M > 512) return fail(GPB_ERR_ARG, "spd_inverse: M=%d unsupported (max 512)", M);
    if (M <= 256) {      // register-resident variant
        auto k256 = gpb::spd_inverse256_kernel<gpb::kInvCluster>;
        const size_t smem256 = gpb::SpdInv256Cfg<gpb::kInvCluster>::smem_bytes;
        int rc256 = allow_smem(k256, smem256);
        if (rc256) return rc256;
        GPB_LAUNCH(k256, dim3(batch * gpb::kInvCluster), dim3(512), smem256, stream, A, M, Ainv, logdet);
        return GPB_CHECK_LAUNCH();
    }
    auto kern = gpb::spd_inverse_kernel<32>;
    const size_t smem = gpb::SpdInvCfg<32>::smem_bytes(M);
    int rc = allow_smem(kern, smem);
    if (rc) return rc;
    GPB_LAUNCH(kern, dim3(batch * gpb::kInvCluster), dim3(512), smem, stream, A, M, Ainv, logdet);
    return GPB_CHECK_LAUNCH();
}

int gpb_probit_lik(const double* m, const double* v, const double* y, const double* gh_x,
                   const double* gh_w, int ngh, double alpha, double scale, long total, int mode,
                   double* dm, double* dv, double* out2, void* ws, size_t ws_bytes, void* stream) {
    if (!m || !v || !y || !gh_x || !gh_w || !dm || !dv || !out2 || total < 1 || (mode != 0 && mode != 1) ||
        ngh < 1 || ngh > 64)
        return fail(GPB_ERR_ARG, "probit_lik: bad argument");
    int grid = elementwise_grid(total);
    if (ws_bytes < sizeof(double) * 2 * (size_t)grid) return fail(GPB_ERR_WS, "probit_lik: workspace too small");
    auto kern = gpb::probit_lik_kernel;
    GPB_LAUNCH(kern, dim3(grid), dim3(256), 0, stream, m, v, y, gh_x, gh_w, ngh, alpha, scale, total, mode,
               dm, dv, (double*)ws);
    launch_reduce_partials((const double*)ws, grid, 2L, 2L, out2, 0, stream);
    return GPB_CHECK_LAUNCH();
}

size_t gpb_gauss_emis_ws_bytes(int n, int Do, int Q) {
    const int DO = emis_pad(Do), QT = emis_pad(Q);
    if (DO < 0 || QT < 0 || n < 1) return 0;
    const size_t NV = 2 + (size_t)DO * (1 + QT);
    return align256(sizeof(double) * NV * ((size_t)emis_grid(n) + 1));
}

int gpb_gauss_emis(const double* mx, const double* vx, const double* y, const double* C, const double* R,
                   double alpha, double scale, int n, int Q, int Do, double* dmx, double* dvx, double* out,
                   void* ws, size_t ws_bytes, void* stream) {
    if (!mx || !vx || !y || !C || !R || !dmx || !dvx || !out || !ws || n < 1)
        return fail(GPB_ERR_ARG, "gauss_emis: bad argument");
    const int DO = emis_pad(Do), QT = emis_pad(Q);
    if (DO < 0 || QT < 0) return fail(GPB_ERR_ARG, "gauss_emis: Do=%d, Q=%d unsupported (max 8)", Do, Q);
    if (ws_bytes < gpb_gauss_emis_ws_bytes(n, Do, Q)) return fail(GPB_ERR_WS, "gauss_emis: workspace too small");
    double* part = (double*)ws;
#define GPB_EMIS(DOV, QTV) \
    if (DO == DOV && QT == QTV) return emis_launch<DOV, QTV>(mx, vx, y, C, R, alpha, scale, n, Q, Do, dmx, dvx, out, part, stream)
    GPB_EMIS(2, 2); GPB_EMIS(2, 4); GPB_EMIS(2, 8);
    GPB_EMIS(4, 2); GPB_EMIS(4, 4); GPB_EMIS(4, 8);
    GPB_EMIS(8, 2); GPB_EMIS(8, 4); GPB_EMIS(8, 8);
#undef GPB_EMIS
    return fail(GPB_ERR_ARG, "gauss_emis: unreachable");
}

int gpb_gauss_emis_finish(const double* raw, const double* R, double alpha, double scale, long Nb, int Do,
                          int Q, double* fin, void* stream) {
    if (!raw || !R || !fin || Do < 1 || Q < 1) return fail(GPB_ERR_ARG, "gauss_emis_finish: bad argument");
    auto kern = gpb::gauss_emis_finish_kernel;
    GPB_LAUNCH(kern, dim3(1), dim3(64), 0, stream, raw, R, alpha, scale, (double)Nb, Do, Q, fin);
    return GPB_CHECK_LAUNCH();
}

int gpb_profile_enable(int on) {
    g_prof_on = on ? 1 : 0;
    return GPB_OK;
}

int gpb_profile_collect(double* h_ms, long* h_count) {
#ifndef GPB_CPU_EMU
    for (int i = 0; i < g_prof_n; i++) {
        ProfPair& p = g_prof_pending[i];
        float ms = 0;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_prof_ms[p.slot] += ms;
            g_prof_cnt[p.slot]++;
        }
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_prof_n = 0;
    for (int i = 0; i < 8; i++) {
        if (h_ms) h_ms[i] = g_prof_ms[i];
        if (h_count) h_count[i] = g_prof_cnt[i];
        g_prof_ms[i] = 0;
        g_prof_cnt[i] = 0;
    }
#else
    for (int i = 0; i < 8; i++) {
        if (h_ms) h_ms[i] = 0;
        if (h_count) h_count[i] = 0;
    }
#endif
    return GPB_OK;
}

int gpb_fma_peak(int prec, long iters, double* sink, double* h_flops, void* stream) {
    if (!sink || iters < 1) return fail(GPB_ERR_ARG, "fma_peak: bad argument");
    // sink[0] > 0 on entry selects blocks per SM (occupancy experiments); default 8
    int blocks = sm_count() * 8;
    if (h_flops && *h_flops >= 1.0 && *h_flops <= 32.0) blocks = sm_count() * (int)(*h_flops);
    if (prec == GPB_F64) {
        auto kern = gpb::fma_peak_kernel<double>;
        GPB_LAUNCH(kern, dim3(blocks), dim3(256), 0, stream, iters, sink);
    } else {
        auto kern = gpb::fma_peak_kernel<float>;
        GPB_LAUNCH(kern, dim3(blocks), dim3(256), 0, stream, iters, sink);
    }
    if (h_flops) *h_flops = (double)blocks * 256.0 * (double)iters * 8.0 * 2.0;
    return GPB_CHECK_LAUNCH();
}

}  // extern "C"
