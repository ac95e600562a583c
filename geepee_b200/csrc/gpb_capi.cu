// gpb_capi.cu -- extern "C" entry points of libgeepee_b200.so (see include/geepee_b200.h).
//
// Host-side launch logic only: tile-shape dispatch, grid sizing in multiples of the SM
// count, workspace carving, deterministic two-stage reductions.  No allocation, no
// synchronisation, no torch types.  The same file is compiled with -DGPB_CPU_EMU by
// tests/emu/build.py, where GPB_LAUNCH runs the kernels on the fiber emulator.
#include "../../include/geepee_b200.h"
#include "gpb_kernels.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace {

char g_err[512] = "";
long g_launches = 0;

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#ifdef GPB_CPU_EMU
}  // namespace
namespace gpb_emu {
void launch_begin(dim3 grid, dim3 block, size_t smem);
bool launch_next_block();
void run_block(void (*tramp)(void*), void* ctx);
}  // namespace gpb_emu
namespace {
template <typename F>
void emu_tramp(void* p) { (*(F*)p)(); }
template <typename F>
void emu_launch(dim3 grid, dim3 block, size_t smem, F f) {
    gpb_emu::launch_begin(grid, block, smem);
    while (gpb_emu::launch_next_block()) gpb_emu::run_block(&emu_tramp<F>, (void*)&f);
}
#define GPB_LAUNCH(kern, grid, block, smem, stream, ...)                          \
    do {                                                                          \
        (void)(stream);                                                           \
        g_launches++;                                                             \
        emu_launch(grid, block, smem, [&]() { kern(__VA_ARGS__); });              \
    } while (0)
#define GPB_CHECK_LAUNCH() GPB_OK
inline int sm_count() { return 2; }
inline void dev_memset(void* p, size_t bytes, void*) { memset(p, 0, bytes); }
template <typename K>
inline int allow_smem(K, size_t) { return GPB_OK; }
#else
#define GPB_LAUNCH(kern, grid, block, smem, stream, ...)                          \
    do {                                                                          \
        g_launches++;                                                             \
        kern<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__);         \
    } while (0)
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GPB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return GPB_OK;
}
#define GPB_CHECK_LAUNCH() check_launch(__func__)
int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
inline void dev_memset(void* p, size_t bytes, void* stream) {
    cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream);
}
template <typename K>
int allow_smem(K kern, size_t bytes) {
    if (bytes <= 48 * 1024) return GPB_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(GPB_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e));
    return GPB_OK;
}
#endif

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline long cdiv(long a, long b) { return (a + b - 1) / b; }

struct Carver {  // bump allocator over the caller's workspace
    char* base;
    size_t off, cap;
    Carver(void* p, size_t c) : base((char*)p), off(0), cap(c) {}
    void* take(size_t bytes) {
        void* r = base ? base + off : nullptr;
        off += align256(bytes);
        return r;
    }
    bool ok() const { return off <= cap; }
};

int elementwise_grid(long total) {
    long b = cdiv(total, 256);
    long cap = (long)sm_count() * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

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
    if (rpb < 32) rpb = 32;
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
    int want = (int)cdiv(2L * sm_count(), (long)p.nbu * Do);
    if (want < 1) want = 1;
    int rps = (int)cdiv(n, want);
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
    GPB_LAUNCH(kern, dim3(grid), dim3(256), C::smem_bytes, stream, a);
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
#define GPB_BWD(DP)                                                                             \
    {                                                                                           \
        auto kern = gpb::det_bwd_kernel<T, DP>;                                                 \
        GPB_LAUNCH(kern, grid, dim3(256), 0, stream, x, z, ls, (const T*)Ap, dm, dv,            \
                   (const T*)Ksave, (const T*)Tsave, n, M, p.MP, D, Do, p.rows_per_block, part, \
                   p.rec_len);                                                                  \
    }
    if (D <= 4) GPB_BWD(4) else if (D <= 8) GPB_BWD(8) else GPB_BWD(16)
#undef GPB_BWD
    int rc = GPB_CHECK_LAUNCH();
    if (rc) return rc;
    auto red = gpb::reduce_partials_kernel;
    GPB_LAUNCH(red, dim3(elementwise_grid(p.rec_len)), dim3(256), 0, stream, part, p.G, p.rec_len,
               p.rec_len, rec, 0);
    auto fin = gpb::det_bwd_finish_kernel;
    GPB_LAUNCH(fin, dim3(1), dim3(256), 0, stream, rec, sf, M, p.MP, D, Do, dA, dzu, dl, dsf2);
    return GPB_CHECK_LAUNCH();
}

template <typename T>
int det_syrk_t(const void* Ksave, const double* dv, int n, int M, int Do, double* dB, void* ws,
               size_t ws_bytes, void* stream) {
    SyrkPlan p = syrk_plan(n, M, Do);
    Carver cv(ws, ws_bytes);
    double* part = (double*)cv.take(sizeof(double) * (size_t)p.nsplit * Do * p.nbu * 128 * 128);
    if (!cv.ok()) return fail(GPB_ERR_WS, "det_syrk: workspace %zu < %zu", ws_bytes, cv.off);
    auto kern = gpb::det_syrk_kernel<T>;
    GPB_LAUNCH(kern, dim3(p.nbu, p.nsplit, Do), dim3(256), 0, stream, (const T*)Ksave, dv, n, p.MP,
               Do, p.rows_per_split, part);
    int rc = GPB_CHECK_LAUNCH();
    if (rc) return rc;
    auto fin = gpb::det_syrk_finish_kernel;
    GPB_LAUNCH(fin, dim3(elementwise_grid((long)Do * M * M)), dim3(256), 0, stream, part, p.nsplit,
               p.MP, M, Do, dB);
    return GPB_CHECK_LAUNCH();
}

// ------------------------------- moment-matched layer -----------------------------------
struct MMPlan {
    int Qt, DOC, RP, PC;
    long P, PP;
    int nchunks, nsplit, rows_per_split, npass;
    int rows_grid, cols_grid, cols_rows_per_block;
    // workspace byte offsets are carved in order by mm_carve
};
int q_template(int Q) {
    if (Q <= 6) return Q;
    if (Q <= 8) return 8;
    if (Q <= 16) return 16;
    return -1;
}
int mm_rp(int Qt, int DOC) {
    int s = 2 * Qt + 2 * DOC + 2;
    return s <= 12 ? 4 : (s <= 26 ? 2 : 1);
}
int mm_rb(int Qt) { return Qt <= 2 ? 8 : (Qt <= 4 ? 4 : (Qt <= 8 ? 2 : 1)); }

MMPlan mm_plan(int n, int M, int Q, int Do) {
    MMPlan p;
    p.Qt = q_template(Q);
    p.DOC = Do == 1 ? 1 : (Do == 2 ? 2 : 4);
    p.npass = (int)cdiv(Do, p.DOC);
    p.RP = mm_rp(p.Qt, p.DOC);
    p.PC = 256 * p.RP;
    p.P = (long)M * (M + 1) / 2;
    p.PP = cdiv(p.P, 1024) * 1024;
    p.nchunks = (int)(p.PP / p.PC);
    int want = (int)cdiv(2L * sm_count(), p.nchunks);
    if (want < 1) want = 1;
    int rb = mm_rb(p.Qt);
    long rps = cdiv(n, want);
    rps = cdiv(rps, rb) * rb;
    if (rps < rb) rps = rb;
    p.rows_per_split = (int)rps;
    p.nsplit = (int)cdiv(n, rps);
    p.rows_grid = (int)cdiv(n, 128);
    if (p.rows_grid > 4 * sm_count()) p.rows_grid = 4 * sm_count();
    if (p.rows_grid < 1) p.rows_grid = 1;
    int cb = 2 * sm_count();
    long crpb = cdiv(n, cb);
    if (crpb < 32) crpb = 32;
    p.cols_rows_per_block = (int)crpb;
    p.cols_grid = (int)cdiv(n, crpb);
    return p;
}

template <typename T>
struct MMWs {
    T *zh, *ep, *bs;
    double *rowacc, *pairpart, *pairsum, *rowpart, *rowsum, *colpart, *colsum, *dZ2, *dlW;
    size_t bytes;
};
template <typename T>
MMWs<T> mm_carve(const MMPlan& p, int n, int M, int Q, int Do, int backward, void* ws, size_t cap) {
    MMWs<T> w;
    Carver cv(ws, cap);
    w.zh = (T*)cv.take(sizeof(T) * p.Qt * p.PP);
    w.ep = (T*)cv.take(sizeof(T) * p.PP);
    w.bs = (T*)cv.take(sizeof(T) * Do * p.PP);
    if (!backward) {
        w.rowacc = (double*)cv.take(sizeof(double) * (size_t)n * Do);
        w.pairpart = w.pairsum = w.rowpart = w.rowsum = w.colpart = w.colsum = w.dZ2 = w.dlW = nullptr;
    } else {
        w.rowacc = (double*)cv.take(sizeof(double) * (size_t)n * (1 + 2 * p.Qt));
        w.pairpart = (double*)cv.take(sizeof(double) * (size_t)p.nsplit * (p.DOC + 1 + p.Qt) * p.PP);
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

template <typename T, int Q, int DOC, bool BWD>
void mm_pairs_launch(const MMPlan& p, const gpb::MMArgs<T>& a, void* stream) {
    auto kern = gpb::mm_pairs_kernel<T, Q, DOC, BWD>;
    GPB_LAUNCH(kern, dim3(p.nchunks, p.nsplit), dim3(256), 0, stream, a);
}
template <typename T, int Q, bool BWD>
int mm_pairs_doc(const MMPlan& p, const gpb::MMArgs<T>& a, void* stream) {
    switch (p.DOC) {
        case 1: mm_pairs_launch<T, Q, 1, BWD>(p, a, stream); return GPB_OK;
        case 2: mm_pairs_launch<T, Q, 2, BWD>(p, a, stream); return GPB_OK;
        case 4: mm_pairs_launch<T, Q, 4, BWD>(p, a, stream); return GPB_OK;
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
             double* vout, void* ws, size_t ws_bytes, void* stream) {
    int rc = mm_check<T>(n, M, Q, Do);
    if (rc) return rc;
    MMPlan p = mm_plan(n, M, Q, Do);
    MMWs<T> w = mm_carve<T>(p, n, M, Q, Do, 0, ws, ws_bytes);
    if (w.bytes > ws_bytes) return fail(GPB_ERR_WS, "mm_fwd: workspace %zu < %zu", ws_bytes, w.bytes);
    auto tab = gpb::mm_pair_table_kernel<T>;
    GPB_LAUNCH(tab, dim3(elementwise_grid(p.PP)), dim3(256), 0, stream, z, ls, sf, B, M, Q, p.Qt, Do,
               p.P, p.PP, w.zh, w.ep, w.bs);
    dev_memset(w.rowacc, sizeof(double) * (size_t)n * Do, stream);
    gpb::MMArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.mx = mx; a.vx = vx; a.ls = ls; a.zh = w.zh; a.ep = w.ep; a.bs = w.bs; a.dv = nullptr;
    a.n = n; a.Qa = Q; a.Do = Do; a.PP = p.PP; a.rows_per_split = p.rows_per_split;
    a.rowacc = w.rowacc; a.pairpart = nullptr; a.full_coef = 0; a.lam_pass = 0;
    for (int pass = 0; pass < p.npass; pass++) {
        a.d0 = pass * p.DOC;
        rc = mm_pairs_dispatch<T, false>(p, a, stream);
        if (rc) return rc;
    }
    rc = GPB_CHECK_LAUNCH();
    if (rc) return rc;
    auto fin = gpb::mm_psi1_fwd_kernel<T>;
    const int nt = 128;
    size_t smem = sizeof(double) * Do * nt + sizeof(T) * ((size_t)M * Q + (size_t)Do * M + 2 * (size_t)Q * nt);
    rc = allow_smem(fin, smem);
    if (rc) return rc;
    GPB_LAUNCH(fin, dim3(p.rows_grid), dim3(nt), smem, stream, mx, vx, z, ls, sf, A, w.rowacc, n, M, Q,
               Do, mout, vout);
    return GPB_CHECK_LAUNCH();
}

template <typename T>
int mm_bwd_t(const double* mx, const double* vx, const double* z, const double* ls, const double* sf,
             const double* A, const double* B, const double* dm, const double* dv,
             const double* mout, int n, int M, int Q, int Do, double* dA, double* dB, double* dzu,
             double* dl, double* dsf2, double* dvsum, double* dmx, double* dvx, void* ws,
             size_t ws_bytes, void* stream) {
    int rc = mm_check<T>(n, M, Q, Do);
    if (rc) return rc;
    MMPlan p = mm_plan(n, M, Q, Do);
    MMWs<T> w = mm_carve<T>(p, n, M, Q, Do, 1, ws, ws_bytes);
    if (w.bytes > ws_bytes) return fail(GPB_ERR_WS, "mm_bwd: workspace %zu < %zu", ws_bytes, w.bytes);
    auto tab = gpb::mm_pair_table_kernel<T>;
    GPB_LAUNCH(tab, dim3(elementwise_grid(p.PP)), dim3(256), 0, stream, z, ls, sf, B, M, Q, p.Qt, Do,
               p.P, p.PP, w.zh, w.ep, w.bs);
    const int NS = 1 + 2 * p.Qt;
    dev_memset(w.rowacc, sizeof(double) * (size_t)n * NS, stream);
    gpb::MMArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.mx = mx; a.vx = vx; a.ls = ls; a.zh = w.zh; a.ep = w.ep; a.bs = w.bs; a.dv = dv;
    a.n = n; a.Qa = Q; a.Do = Do; a.PP = p.PP; a.rows_per_split = p.rows_per_split;
    a.rowacc = w.rowacc; a.pairpart = w.pairpart;
    a.full_coef = p.npass > 1 ? 1 : 0;
    auto red = gpb::reduce_partials_kernel;
    const long recstride = (long)(p.DOC + 1 + p.Qt) * p.PP;
    for (int pass = 0; pass < p.npass; pass++) {
        a.d0 = pass * p.DOC;
        a.lam_pass = pass == 0 ? 1 : 0;
        rc = mm_pairs_dispatch<T, true>(p, a, stream);
        if (rc) return rc;
        // fold the row splits: dBp rows of this d-chunk, and (first pass) S0 | S1
        int nd = (Do - a.d0) < p.DOC ? (Do - a.d0) : p.DOC;
        GPB_LAUNCH(red, dim3(elementwise_grid((long)nd * p.PP)), dim3(256), 0, stream, w.pairpart,
                   p.nsplit, recstride, (long)nd * p.PP, w.pairsum + (long)a.d0 * p.PP, 0);
        if (pass == 0)
            GPB_LAUNCH(red, dim3(elementwise_grid((long)(1 + p.Qt) * p.PP)), dim3(256), 0, stream,
                       w.pairpart + (long)p.DOC * p.PP, p.nsplit, recstride, (long)(1 + p.Qt) * p.PP,
                       w.pairsum + (long)Do * p.PP, 0);
    }
    rc = GPB_CHECK_LAUNCH();
    if (rc) return rc;
    {   // row-wise epilogue: dmx, dvx + row-summed hyper terms
        auto kern = gpb::mm_rows_bwd_kernel<T>;
        const int nt = 128;
        size_t smem = sizeof(double) * (16 + (size_t)(4 * Q + Do) * nt) + sizeof(T) * ((size_t)M * Q + (size_t)Do * M);
        rc = allow_smem(kern, smem);
        if (rc) return rc;
        GPB_LAUNCH(kern, dim3(p.rows_grid), dim3(nt), smem, stream, mx, vx, z, ls, sf, A, dm, dv, mout,
                   w.rowacc, n, M, Q, p.Qt, Do, dmx, dvx, w.rowpart);
        GPB_LAUNCH(red, dim3(1), dim3(256), 0, stream, w.rowpart, p.rows_grid, (long)(2 + Q),
                   (long)(2 + Q), w.rowsum, 0);
    }
    {   // column-wise psi1 part: dA, dZ1
        auto kern = gpb::mm_cols_bwd_kernel<T>;
        const int nt = 128;
        size_t smem = sizeof(double) * ((size_t)32 * (2 * Q + 1 + Do) + (size_t)(2 * Q + 2 * Do) * nt);
        rc = allow_smem(kern, smem);
        if (rc) return rc;
        GPB_LAUNCH(kern, dim3(p.cols_grid), dim3(nt), smem, stream, mx, vx, z, ls, sf, A, dm, dv, mout, n,
                   M, Q, Do, p.cols_rows_per_block, w.colpart);
        long len = (long)Do * M + (long)M * Q;
        GPB_LAUNCH(red, dim3(elementwise_grid(len)), dim3(256), 0, stream, w.colpart, p.cols_grid, len,
                   len, w.colsum, 0);
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

// =========================================================================================
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
    auto red = gpb::reduce_partials_kernel;
    GPB_LAUNCH(red, dim3(1), dim3(256), 0, stream, (const double*)ws, grid, 2L, 2L, out2, 0);
    return GPB_CHECK_LAUNCH();
}

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

size_t gpb_det_syrk_ws_bytes(int n, int M, int Do) {
    if (gpb_det_pad_m(M) < 0) return 0;
    SyrkPlan p = syrk_plan(n, M, Do);
    return align256(sizeof(double) * (size_t)p.nsplit * Do * p.nbu * 128 * 128);
}

int gpb_det_syrk(int prec, const void* Ksave, const double* dv, int n, int M, int Do, double* dB,
                 void* ws, size_t ws_bytes, void* stream) {
    if (!Ksave || !dv || !dB || !ws || gpb_det_pad_m(M) < 0 || n < 1 || Do < 1)
        return fail(GPB_ERR_ARG, "det_syrk: bad argument");
    if (prec == GPB_F64) return det_syrk_t<double>(Ksave, dv, n, M, Do, dB, ws, ws_bytes, stream);
    return det_syrk_t<float>(Ksave, dv, n, M, Do, dB, ws, ws_bytes, stream);
}

size_t gpb_mm_ws_bytes(int n, int M, int Q, int Do, int backward) {
    if (q_template(Q) < 0 || n < 1 || M < 1 || Do < 1) return 0;
    MMPlan p = mm_plan(n, M, Q, Do);
    return mm_carve<double>(p, n, M, Q, Do, backward, nullptr, 0).bytes;  // fp64 sizing covers fp32
}

int gpb_mm_fwd(int prec, const double* mx, const double* vx, const double* z, const double* ls,
               const double* sf, const double* A, const double* B, int n, int M, int Q, int Do,
               double* mout, double* vout, void* ws, size_t ws_bytes, void* stream) {
    if (!mx || !vx || !z || !ls || !sf || !A || !B || !mout || !vout || !ws) return fail(GPB_ERR_ARG, "mm_fwd: null pointer");
    if (prec == GPB_F64) return mm_fwd_t<double>(mx, vx, z, ls, sf, A, B, n, M, Q, Do, mout, vout, ws, ws_bytes, stream);
    return mm_fwd_t<float>(mx, vx, z, ls, sf, A, B, n, M, Q, Do, mout, vout, ws, ws_bytes, stream);
}

int gpb_mm_bwd(int prec, const double* mx, const double* vx, const double* z, const double* ls,
               const double* sf, const double* A, const double* B, const double* dm, const double* dv,
               const double* mout, int n, int M, int Q, int Do, double* dA, double* dB, double* dzu,
               double* dl, double* dsf2, double* dvsum, double* dmx, double* dvx, void* ws,
               size_t ws_bytes, void* stream) {
    if (!mx || !vx || !z || !ls || !sf || !A || !B || !dm || !dv || !mout || !dA || !dB || !dzu || !dl ||
        !dsf2 || !dvsum || !dmx || !dvx || !ws)
        return fail(GPB_ERR_ARG, "mm_bwd: null pointer");
    if (prec == GPB_F64)
        return mm_bwd_t<double>(mx, vx, z, ls, sf, A, B, dm, dv, mout, n, M, Q, Do, dA, dB, dzu, dl, dsf2, dvsum, dmx, dvx, ws, ws_bytes, stream);
    return mm_bwd_t<float>(mx, vx, z, ls, sf, A, B, dm, dv, mout, n, M, Q, Do, dA, dB, dzu, dl, dsf2, dvsum, dmx, dvx, ws, ws_bytes, stream);
}

int gpb_fma_peak(int prec, long iters, double* sink, double* h_flops, void* stream) {
    if (!sink || iters < 1) return fail(GPB_ERR_ARG, "fma_peak: bad argument");
    int blocks = sm_count() * 8;
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
