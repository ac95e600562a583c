// gpb_capi_tail.cu -- the replicated O(Dout M^3) tail as a program of batched fp64 primitives
// (include/geepee_b200.h: GpbTailOp, gpb_tail_exec, gpb_tail_gather; kernels in gpb_tail.cuh).
#include "gpb_common.cuh"
#include "gpb_tail.cuh"

namespace {

gpb::TailOp to_dev_op(const GpbTailOp& h) {
    gpb::TailOp o;
    o.kind = h.kind; o.flags = h.flags; o.batch = h.batch; o.m = h.m; o.n = h.n; o.k = h.k;
    for (int i = 0; i < 6; i++) { o.src[i] = h.src[i]; o.sstride[i] = h.sstride[i]; o.ld[i] = h.ld[i]; }
    for (int i = 0; i < 8; i++) o.coef[i] = h.coef[i];
    o.dst = h.dst; o.dstride = h.dstride; o.ldd = h.ldd;
    return o;
}

template <int BM, int BN, int WM, int WN, bool VEC>
int launch_gemm(const gpb::TailOp& o, void* stream) {
    typedef gpb::TailGemmCfg<BM, BN, WM, WN> C;
    auto kern = gpb::tail_gemm_kernel<BM, BN, WM, WN, VEC>;
    int rc = allow_smem(kern, C::smem_bytes);
    if (rc != GPB_OK) return rc;
    dim3 grid((unsigned)cdiv(o.n, BN), (unsigned)cdiv(o.m, BM), (unsigned)o.batch);
    GPB_LAUNCH(kern, grid, dim3(C::NT), C::smem_bytes, stream, o);
    return GPB_OK;
}

int run_op(const GpbTailOp& h, void* stream) {
    if (!h.dst || h.batch < 1) return fail(GPB_ERR_ARG, "tail op %d: bad argument", h.kind);
    gpb::TailOp o = to_dev_op(h);
    switch (h.kind) {
    case GPB_TOP_GEMM: {
        if (!h.src[0] || !h.src[1] || h.m < 1 || h.n < 1 || h.k < 1) return fail(GPB_ERR_ARG, "tail gemm: bad argument");
        // small problems: 32x32 tiles so that more SMs take part (the tail is latency bound)
        const long tiles64 = cdiv(h.m, 64) * cdiv(h.n, 64) * (long)h.batch;
        // 16-byte staging needs even leading dimensions / batch strides and 16-byte aligned bases
        const bool vec = ((h.ld[0] | h.ld[1]) & 1) == 0 && ((h.sstride[0] | h.sstride[1]) & 1) == 0 &&
                         (((size_t)h.src[0] | (size_t)h.src[1]) & 15) == 0;
        const bool big = tiles64 >= sm_count() || (h.m > 32 && h.n > 32 && tiles64 * 4 > 6L * sm_count());
        if (big) return vec ? launch_gemm<64, 64, 2, 4, true>(o, stream) : launch_gemm<64, 64, 2, 4, false>(o, stream);
        return vec ? launch_gemm<32, 32, 2, 2, true>(o, stream) : launch_gemm<32, 32, 2, 2, false>(o, stream);
    }
    case GPB_TOP_LINCOMB: {
        if (h.m < 1 || h.n < 1) return fail(GPB_ERR_ARG, "tail lincomb: bad argument");
        if ((h.src[4] == nullptr) != (h.src[5] == nullptr)) return fail(GPB_ERR_ARG, "tail lincomb: outer product needs u and v");
        const long total = (long)h.m * h.n * (((h.flags >> 8) & 1) ? 1 : h.batch);
        auto kern = gpb::tail_lincomb_kernel;
        GPB_LAUNCH(kern, dim3(elementwise_grid(total)), dim3(256), 0, stream, o);
        return GPB_OK;
    }
    case GPB_TOP_MATVEC: {
        if (h.m < 1 || h.k < 1) return fail(GPB_ERR_ARG, "tail matvec: bad argument");
        auto kern = gpb::tail_matvec_kernel;
        const long warps = (long)h.batch * h.m;
        long blocks = cdiv(warps, 8);
        const long cap = (long)sm_count() * 8;
        GPB_LAUNCH(kern, dim3((unsigned)(blocks > cap ? cap : blocks)), dim3(256), 0, stream, o);
        return GPB_OK;
    }
    case GPB_TOP_DOTS: {
        auto kern = gpb::tail_dots_kernel;
        GPB_LAUNCH(kern, dim3(1), dim3(1024), 0, stream, o);
        return GPB_OK;
    }
    case GPB_TOP_UNPACK_R: {
        if (!h.src[0] || h.m < 1) return fail(GPB_ERR_ARG, "tail unpack_r: bad argument");
        auto kern = gpb::tail_unpack_r_kernel;
        GPB_LAUNCH(kern, dim3(elementwise_grid((long)h.batch * h.m * h.m)), dim3(256), 0, stream, o);
        return GPB_OK;
    }
    case GPB_TOP_PACK_R: {
        if (!h.src[0] || !h.src[1] || h.m < 1) return fail(GPB_ERR_ARG, "tail pack_r: bad argument");
        auto kern = gpb::tail_pack_r_kernel;
        GPB_LAUNCH(kern, dim3(elementwise_grid((long)h.batch * h.m * (h.m + 1) / 2)), dim3(256), 0, stream, o);
        return GPB_OK;
    }
    case GPB_TOP_KHYPER: {
        for (int i = 0; i < 6; i++)
            if (!h.src[i]) return fail(GPB_ERR_ARG, "tail khyper: null operand %d", i);
        if (h.m < 1 || h.k < 1 || h.k > 32) return fail(GPB_ERR_ARG, "tail khyper: D=%d unsupported (1..32)", h.k);
        // scratch for the per-block shares of the D + 1 scalar sums: behind the output record
        // (the caller sizes dst as gpb_tail_khyper_out_len(M, D) doubles)
        const int nb = (int)cdiv(h.m, 8);
        double* part = h.dst + 1 + h.k + (long)h.m * h.k;
        int dmax;
        if (h.k <= 4) { dmax = 4; auto kern = gpb::tail_khyper_kernel<4>; GPB_LAUNCH(kern, dim3(nb), dim3(256), 0, stream, o, part); }
        else if (h.k <= 8) { dmax = 8; auto kern = gpb::tail_khyper_kernel<8>; GPB_LAUNCH(kern, dim3(nb), dim3(256), 0, stream, o, part); }
        else if (h.k <= 16) { dmax = 16; auto kern = gpb::tail_khyper_kernel<16>; GPB_LAUNCH(kern, dim3(nb), dim3(256), 0, stream, o, part); }
        else { dmax = 32; auto kern = gpb::tail_khyper_kernel<32>; GPB_LAUNCH(kern, dim3(nb), dim3(256), 0, stream, o, part); }
        auto fin = gpb::tail_khyper_finish_kernel;
        GPB_LAUNCH(fin, dim3(1), dim3(64), 0, stream, o, (const double*)part, nb, dmax);
        return GPB_OK;
    }
    case GPB_TOP_SUM: {
        if (!h.src[0] || !h.src[1] || h.sstride[0] < 0) return fail(GPB_ERR_ARG, "tail sum: bad argument");
        const long count = h.sstride[0];
        long blocks = cdiv(count > 0 ? count : 1, 4096);
        if (blocks > 1024) blocks = 1024;
        double* part = const_cast<double*>(h.src[1]);
        auto kern = gpb::tail_sum_partial_kernel;
        GPB_LAUNCH(kern, dim3((unsigned)blocks), dim3(256), 0, stream, h.src[0], count, part);
        launch_reduce_partials((const double*)part, (int)blocks, 1L, 1L, h.dst, 0, stream);
        return GPB_OK;
    }
    default:
        return fail(GPB_ERR_ARG, "tail op: unknown kind %d", h.kind);
    }
}

}  // namespace

extern "C" {

int gpb_tail_exec(const GpbTailOp* h_ops, int n_ops, void* stream) {
    if (!h_ops || n_ops < 0) return fail(GPB_ERR_ARG, "tail_exec: bad argument");
    for (int i = 0; i < n_ops; i++) {
        int rc = run_op(h_ops[i], stream);
        if (rc != GPB_OK) return rc;
    }
    return GPB_CHECK_LAUNCH();
}

long gpb_tail_khyper_out_len(int M, int D) {
    return 1 + D + (long)M * D + cdiv(M, 8) * 33;
}

int gpb_tail_gather(int n, const double* const* h_srcs, const long* h_counts, double scale, double* dst,
                    void* stream) {
    if (n < 0 || (n > 0 && (!h_srcs || !h_counts)) || !dst) return fail(GPB_ERR_ARG, "tail_gather: bad argument");
    long done = 0;
    for (int i0 = 0; i0 < n; i0 += gpb::GatherArgs::MAXN) {
        gpb::GatherArgs a;
        a.n = (n - i0) < gpb::GatherArgs::MAXN ? (n - i0) : gpb::GatherArgs::MAXN;
        a.off[0] = 0;
        for (int i = 0; i < a.n; i++) {
            if (!h_srcs[i0 + i] || h_counts[i0 + i] < 0) return fail(GPB_ERR_ARG, "tail_gather: bad source %d", i0 + i);
            a.src[i] = h_srcs[i0 + i];
            a.off[i + 1] = a.off[i] + h_counts[i0 + i];
        }
        a.scale = scale;
        a.dst = dst + done;
        if (a.off[a.n] > 0) {
            auto kern = gpb::tail_gather_kernel;
            GPB_LAUNCH(kern, dim3(elementwise_grid(a.off[a.n])), dim3(256), 0, stream, a);
        }
        done += a.off[a.n];
    }
    return GPB_CHECK_LAUNCH();
}

int gpb_tail_copy(int n, const double* const* h_srcs, double* const* h_dsts, const long* h_counts,
                  void* stream) {
    if (n < 0 || (n > 0 && (!h_srcs || !h_dsts || !h_counts))) return fail(GPB_ERR_ARG, "tail_copy: bad argument");
    for (int i0 = 0; i0 < n; i0 += gpb::CopyArgs::MAXN) {
        gpb::CopyArgs a;
        a.n = (n - i0) < gpb::CopyArgs::MAXN ? (n - i0) : gpb::CopyArgs::MAXN;
        a.off[0] = 0;
        for (int i = 0; i < a.n; i++) {
            if (!h_srcs[i0 + i] || !h_dsts[i0 + i] || h_counts[i0 + i] < 0)
                return fail(GPB_ERR_ARG, "tail_copy: bad entry %d", i0 + i);
            a.src[i] = h_srcs[i0 + i];
            a.dst[i] = h_dsts[i0 + i];
            a.off[i + 1] = a.off[i] + h_counts[i0 + i];
        }
        if (a.off[a.n] > 0) {
            auto kern = gpb::tail_copy_kernel;
            GPB_LAUNCH(kern, dim3(elementwise_grid(a.off[a.n])), dim3(256), 0, stream, a);
        }
    }
    return GPB_CHECK_LAUNCH();
}

/* Device-side address of page-locked (pinned, mapped) host memory, so that a copy KERNEL can read it directly:
 * small per-step uploads (the parameter vector) then bypass the H2D copy engine, whose FIFO would otherwise put them
 * behind any large input copy that is in flight.  Returns GPB_ERR_ARG if the memory is not mapped. */
int gpb_host_device_ptr(const void* host_ptr, void** dev_ptr) {
    if (!host_ptr || !dev_ptr) return fail(GPB_ERR_ARG, "host_device_ptr: null pointer");
#ifndef GPB_CPU_EMU
    void* d = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&d, const_cast<void*>(host_ptr), 0);
    if (e != cudaSuccess || !d) {
        (void)cudaGetLastError();      // not sticky: clear it
        return fail(GPB_ERR_ARG, "host_device_ptr: memory is not mapped pinned host memory (%s)",
                    cudaGetErrorString(e));
    }
    *dev_ptr = d;
#else
    *dev_ptr = const_cast<void*>(host_ptr);
#endif
    return GPB_OK;
}

}  // extern "C"
