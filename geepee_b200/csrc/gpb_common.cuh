// gpb_common.cuh -- shared host-side helpers of the C-ABI translation units
// (gpb_capi_misc.cu, gpb_capi_det.cu, gpb_capi_mm.cu; see include/geepee_b200.h).
//
// Host-side launch logic only: tile-shape dispatch, grid sizing in multiples of the SM
// count, workspace carving, deterministic two-stage reductions.  No allocation, no
// synchronisation, no torch types.  The same file is compiled with -DGPB_CPU_EMU by
// tests/emu/build.py, where GPB_LAUNCH runs the kernels on the fiber emulator.
#pragma once
// Development knobs of the fp64 pair kernels (A/B builds, geepee_b200/build.py); the defaults are the
// product configuration.  GPB_MM_RP64: cap on the pairs a thread owns (0 = automatic);
// GPB_EXP_REP: replicas of the exp table in shared memory; GPB_MM_NR_FWD: rows per loop trip of
// the forward kernel.
#ifndef GPB_MM_RP64
#define GPB_MM_RP64 0
#endif
// largest per-pair state size (2Q + 2DOC + 2 values) for which a thread owns 4 pairs
#ifndef GPB_MM_RP4_MAX
#define GPB_MM_RP4_MAX 18
#endif
// (a second __launch_bounds__ argument is deliberately NOT exposed: `(256, 1)` let ptxas take 176
//  registers for the backward kernel -> one CTA per SM, +12 % time; `(256, 3)` is the rejected
//  3-CTA experiment of DESIGN.md section 7)
#ifndef GPB_EXP_REP
#define GPB_EXP_REP 16
#endif
// bit-field argument reduction of the fp64 exp in the pair kernels (gpb_kernels.cuh, ExpBits)
#ifndef GPB_EXP_BITS
#define GPB_EXP_BITS 1
#endif
// narrow fp64 moment-matched layers: pair kernels with the exponent on the FP64 tensor cores (gpb_pairsx.cuh)
// mm_pairs_tc_kernel, single-output passes: how many of every 4 exponentials are evaluated on the FMA pipe
// instead of the SFU
#ifndef GPB_MM_TC_POLY
#define GPB_MM_TC_POLY 1
#endif
// fp32 dB rank update of the deterministic layer on tcgen05 (gpb_umma.cuh)
#ifndef GPB_DET_SYRK_TC
#define GPB_DET_SYRK_TC 1
#endif
// fp32 forward of narrow moment-matched layers: exponent GEMM on tcgen05 (gpb_umma.cuh)
#ifndef GPB_MM_TC
#define GPB_MM_TC 1
#endif
#ifndef GPB_MM_XPATH
#define GPB_MM_XPATH 1
#endif
// ... and their backward twin.  Measured on the B200 (tools/kbench.py mm, n = 32768): forward 1.368 vs 1.460 ms
// (M=256, Q=2, Do=2), backward 3.43 vs 2.85 ms -- a block of the tensor-exponent kernel covers 128 pairs instead
// of 1024, so the per-tile barrier, the row staging and the cross-warp row sums are amortised over 8x fewer
// pairs, which costs the backward (4 + 2Q row sums per row) more than the tensor instruction saves.  Off.
#ifndef GPB_MM_XPATH_BWD
#define GPB_MM_XPATH_BWD 0
#endif
#ifndef GPB_SYRK_WAVES
#define GPB_SYRK_WAVES 2
#endif
#ifndef GPB_MM_NR_FWD
#define GPB_MM_NR_FWD 2
#endif
// warps per CTA of the wide-layer tensor-core backward (8 or 16; measured equal on the B200: 14.17 vs 14.04 ms at the cfg2 shape)
#ifndef GPB_MM_WIDE_WARPS
#define GPB_MM_WIDE_WARPS 8
#endif
#include "../../include/geepee_b200.h"
#include "gpb_kernels.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

// shared state lives in gpb_capi_misc.cu
extern char g_gpb_err[512];
extern long g_gpb_launches;
extern int g_gpb_prof_on;
#define g_err g_gpb_err
#define g_launches g_gpb_launches
#define g_prof_on g_gpb_prof_on

namespace {

inline int fail(int code, const char* fmt, ...) {
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
inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GPB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return GPB_OK;
}
#define GPB_CHECK_LAUNCH() check_launch(__func__)
inline int sm_count() {
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

// ---- optional per-kernel device timing (CUDA events on the launching stream) -------------
// slots: 0 det_fwd, 1 det_bwd, 2 det_syrk, 3 mm_pairs fwd, 4 mm_pairs bwd, 5 mm_rows_bwd,
//        6 mm_cols_bwd, 7 unused
#ifdef GPB_CPU_EMU
inline void prof_begin(int, void*) {}
inline void prof_end(int, void*) {}
#else
}  // namespace
struct GpbProfPair { cudaEvent_t a, b; int slot; };
extern GpbProfPair g_gpb_prof_pending[4096];
extern int g_gpb_prof_n;
namespace {
#define g_prof_pending g_gpb_prof_pending
#define g_prof_n g_gpb_prof_n
typedef GpbProfPair ProfPair;
inline void prof_begin(int slot, void* stream) {
    if (!g_prof_on || g_prof_n >= 4096) return;
    ProfPair& p = g_prof_pending[g_prof_n];
    p.slot = slot;
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
    cudaEventRecord(p.a, (cudaStream_t)stream);
}
inline void prof_end(int slot, void* stream) {
    if (!g_prof_on || g_prof_n >= 4096) return;
    (void)slot;
    cudaEventRecord(g_prof_pending[g_prof_n].b, (cudaStream_t)stream);
    g_prof_n++;
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

inline int elementwise_grid(long total) {
    long b = cdiv(total, 256);
    long cap = (long)sm_count() * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// second stage of the deterministic two-stage reductions: out[i] (+)= sum_g part[g*gstride + i]
inline void launch_reduce_partials(const double* part, int G, long gstride, long len, double* out,
                                   int accumulate, void* stream) {
    if (G >= 64) {      // a thread per output would walk G partials serially (1184 after gauss_lik: 85 us)
        auto red = gpb::reduce_partials_warp_kernel;
        long blocks = cdiv(len, 8);
        const long cap = (long)sm_count() * 8;
        GPB_LAUNCH(red, dim3((unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks))), dim3(256), 0, stream,
                   part, G, gstride, len, out, accumulate);
    } else {
        auto red = gpb::reduce_partials_kernel;
        GPB_LAUNCH(red, dim3(elementwise_grid(len)), dim3(256), 0, stream, part, G, gstride, len, out, accumulate);
    }
}

}  // namespace
