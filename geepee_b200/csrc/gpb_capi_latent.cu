// gpb_capi_latent.cu -- C ABI of the elementwise latent-variable kernels (gpb_latent.cuh;
// SURVEY.md section 8 rows a12 / a13).
#include "gpb_common.cuh"
#include "gpb_latent.cuh"

namespace {
int latent_grid(long total) { return elementwise_grid(total); }
}

extern "C" {

size_t gpb_latent_ws_bytes(long total) { return align256(sizeof(double) * 2 * ((size_t)latent_grid(total) + 1)); }

int gpb_lvm_x_fwd(int mode, int nat, const double* x1, const double* x2, const long* sel, long lo, int n, int Q,
                  double prior1, double prior2, double alpha, double* m, double* v, void* stream) {
    if (!x1 || !x2 || !m || !v || n < 1 || Q < 1 || (mode != 0 && mode != 1))
        return fail(GPB_ERR_ARG, "lvm_x_fwd: bad argument");
    gpb::LvmArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = mode; a.nat = nat; a.x1 = x1; a.x2 = x2; a.sel = sel; a.lo = lo; a.n = n; a.Q = Q;
    a.prior1 = prior1; a.prior2 = prior2; a.alpha = alpha; a.o1 = m; a.o2 = v;
    auto kern = gpb::lvm_x_fwd_kernel;
    GPB_LAUNCH(kern, dim3(latent_grid((long)n * Q)), dim3(256), 0, stream, a);
    return GPB_CHECK_LAUNCH();
}

int gpb_lvm_x_bwd(int mode, int nat, const double* x1, const double* x2, const long* sel, long lo, int n, long N,
                  int Q, double prior1, double prior2, double alpha, double s_cav, double s_post,
                  const double* dmx, const double* dvx, double* gx1, double* gx2, double* sums, void* ws,
                  size_t ws_bytes, void* stream) {
    if (!x1 || !x2 || !dmx || !dvx || !gx1 || !gx2 || !sums || n < 1 || Q < 1 || N < n || (mode != 0 && mode != 1))
        return fail(GPB_ERR_ARG, "lvm_x_bwd: bad argument");
    const long total = (long)n * Q;
    if (!ws || ws_bytes < gpb_latent_ws_bytes(total)) return fail(GPB_ERR_WS, "lvm_x_bwd: workspace too small");
    // rows outside this call's selection carry a zero gradient (aep_models.py:803-806)
    if (sel) {
        dev_memset(gx1, sizeof(double) * (size_t)N * Q, stream);
        dev_memset(gx2, sizeof(double) * (size_t)N * Q, stream);
    } else {
        if (lo > 0) {
            dev_memset(gx1, sizeof(double) * (size_t)lo * Q, stream);
            dev_memset(gx2, sizeof(double) * (size_t)lo * Q, stream);
        }
        if (lo + n < N) {
            dev_memset(gx1 + (lo + n) * Q, sizeof(double) * (size_t)(N - lo - n) * Q, stream);
            dev_memset(gx2 + (lo + n) * Q, sizeof(double) * (size_t)(N - lo - n) * Q, stream);
        }
    }
    gpb::LvmArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = mode; a.nat = nat; a.x1 = x1; a.x2 = x2; a.sel = sel; a.lo = lo; a.n = n; a.Q = Q;
    a.prior1 = prior1; a.prior2 = prior2; a.alpha = alpha; a.s_cav = s_cav; a.s_post = s_post;
    a.dmx = dmx; a.dvx = dvx; a.o1 = gx1; a.o2 = gx2; a.part = (double*)ws;
    const int grid = latent_grid(total);
    auto kern = gpb::lvm_x_bwd_kernel;
    GPB_LAUNCH(kern, dim3(grid), dim3(256), 0, stream, a);
    launch_reduce_partials((const double*)ws, grid, 2L, 2L, sums, 0, stream);
    return GPB_CHECK_LAUNCH();
}

int gpb_ssm_cavity(const double* xf1, const double* xf2, long T, int Q, double prior1, double prior2,
                   double alpha, double* cav_m, double* cav_v, void* stream) {
    if (!xf1 || !xf2 || !cav_m || !cav_v || T < 2 || Q < 1) return fail(GPB_ERR_ARG, "ssm_cavity: bad argument");
    gpb::SsmArgs a = {xf1, xf2, T, Q, prior1, prior2, alpha};
    auto kern = gpb::ssm_cavity_kernel;
    GPB_LAUNCH(kern, dim3(latent_grid(T * Q)), dim3(256), 0, stream, a, cav_m, cav_v);
    return GPB_CHECK_LAUNCH();
}

int gpb_ssm_transition(const double* mt, const double* vt, const double* mp, const double* vp, const double* sn,
                       long total, double alpha, double s_dyn, double* dm_layer, double* dvt, double* sums,
                       void* ws, size_t ws_bytes, void* stream) {
    if (!mt || !vt || !mp || !vp || !sn || !dm_layer || !dvt || !sums || total < 1)
        return fail(GPB_ERR_ARG, "ssm_transition: bad argument");
    if (!ws || ws_bytes < gpb_latent_ws_bytes(total)) return fail(GPB_ERR_WS, "ssm_transition: workspace too small");
    const int grid = latent_grid(total);
    auto kern = gpb::ssm_transition_kernel;
    GPB_LAUNCH(kern, dim3(grid), dim3(256), 0, stream, mt, vt, mp, vp, sn, total, alpha, s_dyn, dm_layer, dvt,
               (double*)ws);
    launch_reduce_partials((const double*)ws, grid, 2L, 2L, sums, 0, stream);
    return GPB_CHECK_LAUNCH();
}

int gpb_ssm_sources(const double* xf1, const double* xf2, long T, int Q, double prior1, double prior2, double alpha,
                    const double* prev_dm, const double* prev_dv, long prev_first, long prev_count, int prev_ld,
                    const double* next_dm, const double* next_dv, long next_first, long next_count, int next_ld,
                    const double* up_dm, const double* up_dv, long up_first, long up_count, int up_ld,
                    double* l1, double* l2, void* stream) {
    if (!xf1 || !xf2 || !l1 || !l2 || T < 2 || Q < 1) return fail(GPB_ERR_ARG, "ssm_sources: bad argument");
    gpb::SsmArgs a = {xf1, xf2, T, Q, prior1, prior2, alpha};
    gpb::SsmSrc prev = {prev_dm, prev_dv, prev_first, prev_count, prev_ld, -1.0};
    gpb::SsmSrc next = {next_dm, next_dv, next_first, next_count, next_ld, 1.0};
    gpb::SsmSrc up = {up_dm, up_dv, up_first, up_count, up_ld, 1.0};
    auto kern = gpb::ssm_sources_kernel;
    GPB_LAUNCH(kern, dim3(latent_grid(T * Q)), dim3(256), 0, stream, a, prev, next, up, l1, l2);
    return GPB_CHECK_LAUNCH();
}

int gpb_ssm_xfinal(const double* xf1, const double* xf2, long T, int Q, double prior1, double prior2, double alpha,
                   const double* l1, const double* l2, double* gx1, double* gx2, double* sums, void* ws,
                   size_t ws_bytes, void* stream) {
    if (!xf1 || !xf2 || !l1 || !l2 || !gx1 || !gx2 || !sums || T < 2 || Q < 1)
        return fail(GPB_ERR_ARG, "ssm_xfinal: bad argument");
    if (!ws || ws_bytes < gpb_latent_ws_bytes(T * Q)) return fail(GPB_ERR_WS, "ssm_xfinal: workspace too small");
    gpb::SsmArgs a = {xf1, xf2, T, Q, prior1, prior2, alpha};
    const int grid = latent_grid(T * Q);
    auto kern = gpb::ssm_xfinal_kernel;
    GPB_LAUNCH(kern, dim3(grid), dim3(256), 0, stream, a, l1, l2, gx1, gx2, (double*)ws);
    launch_reduce_partials((const double*)ws, grid, 2L, 2L, sums, 0, stream);
    return GPB_CHECK_LAUNCH();
}

}  // extern "C"
