// gpb_rt.cuh -- thin runtime shim under the geepee_b200 kernels.
//
// Product build (nvcc, sm_100a): everything below maps 1:1 onto CUDA built-ins
// and inline PTX (cp.async).  Test build (-DGPB_CPU_EMU, g++ only, used by
// tests/emu/ -- never by the product): the same kernel source is compiled for the
// host and run by a fiber scheduler that executes every CUDA thread of a block
// as a coroutine, so indexing / reduction / barrier logic can be checked
// against the oracle in the GPU-less build container.
#pragma once

#ifndef GPB_CPU_EMU
// ============================ CUDA ========================================
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

#define GPB_DEVICE __device__ __forceinline__
#define GPB_MEMBER __device__ __forceinline__
#define GPB_KERNEL static __global__
#define GPB_SHARED __shared__
#define GPB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define GPB_LAUNCH_BOUNDS(n) __launch_bounds__(n)
#define GPB_LAUNCH_BOUNDS2(n, m) __launch_bounds__(n, m)
#define GPB_UNROLL _Pragma("unroll")
#define GPB_UNROLL_N(n) _Pragma(GPB_STR(unroll n))
#define GPB_STR(x) #x
#define GPB_ALIGN16 __align__(16)

namespace gpb {

GPB_DEVICE void sync_threads() { __syncthreads(); }
// thread-block cluster (sm_90+): hardware barrier over the CTAs of a cluster; global-memory
// writes made before it are visible to the whole cluster after it
#define GPB_CLUSTER(n) __cluster_dims__(n, 1, 1)
GPB_DEVICE void cluster_sync() {
    __threadfence();
    cooperative_groups::this_cluster().sync();
}
// the cluster barrier alone: arrive has release and wait has acquire semantics at cluster scope, which already
// orders the global-memory accesses of the cluster's threads
GPB_DEVICE void cluster_barrier() { cooperative_groups::this_cluster().sync(); }
GPB_DEVICE int cluster_rank() { return (int)cooperative_groups::this_cluster().block_rank(); }
GPB_DEVICE void sync_warp() { __syncwarp(); }
// named barrier over the first `nthreads` threads of the block (a multiple of 32; all of them call it)
template <int ID, int NTHREADS>
GPB_DEVICE void sync_group() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory"); }
GPB_DEVICE double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
GPB_DEVICE float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
GPB_DEVICE void atomic_add(double* p, double v) { atomicAdd(p, v); }
// FP64 tensor-core MMA (DMMA.8x8x4): C[8x8] += A[8x4] * B[4x8].  Fragments, with g = lane / 4 and
// t = lane % 4:  a = A[g][t],  b = B[t][g],  c0 = C[g][2t],  c1 = C[g][2t+1].
GPB_DEVICE void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 16-byte async global->shared copy (LDGSTS).  Both pointers 16B aligned.
GPB_DEVICE void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
// same, but writes 16 zero bytes when !valid (src-size 0: nothing is read from gmem_src)
GPB_DEVICE void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(n));
}
// 16-byte copy of which only the first `nbytes` (0, 8 or 16) are read; the rest is zero filled
GPB_DEVICE void cp_async16_bytes(void* smem_dst, const void* gmem_src, int nbytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(nbytes));
}
// 8-byte variant (element-granular staging of small per-row records), zero-fill when !valid
GPB_DEVICE void cp_async8_zfill(void* smem_dst, const void* gmem_src, bool valid) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int n = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem_src), "r"(n));
}
GPB_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
GPB_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

GPB_DEVICE double fast_exp(double x) { return exp(x); }
GPB_DEVICE float fast_exp(float x) { return __expf(x); }
GPB_DEVICE double ldg(const double* p) { return __ldg(p); }
GPB_DEVICE float ldg(const float* p) { return __ldg(p); }

}  // namespace gpb

#else
// ============================ CPU emulation ===============================
#include <math.h>
#include <stdint.h>
#include <string.h>

#define GPB_DEVICE static inline
#define GPB_MEMBER inline
#define GPB_KERNEL static
#define GPB_SHARED static
#define GPB_DYN_SMEM(name) unsigned char* name = gpb_emu::dyn_smem
#define GPB_LAUNCH_BOUNDS(n)
#define GPB_LAUNCH_BOUNDS2(n, m)
#define GPB_UNROLL
#define GPB_UNROLL_N(n)
#define GPB_ALIGN16 __attribute__((aligned(16)))
#define __restrict__

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

namespace gpb_emu {
extern dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern unsigned char* dyn_smem;
void barrier();                       // yield until every fiber of the block arrives
void group_barrier(int nthreads);     // yield until the first `nthreads` fibers arrive
double shfl_xor_f64(double v, int m);  // warp exchange through a mailbox
double shfl_idx_f64(double v, int src); // read `v` of lane `src`
void dmma_f64(double a, double b, double* d0, double* d1);  // one 8x8x4 MMA: returns A.B for this lane
}  // namespace gpb_emu

#define threadIdx (gpb_emu::t_threadIdx)
#define blockIdx (gpb_emu::t_blockIdx)
#define blockDim (gpb_emu::t_blockDim)
#define gridDim (gpb_emu::t_gridDim)

namespace gpb {

static inline void sync_threads() { gpb_emu::barrier(); }
// the emulator runs blocks one after another: cluster kernels are built with clusters of ONE block
#define GPB_CLUSTER(n)
static inline void cluster_sync() { gpb_emu::barrier(); }
static inline void cluster_barrier() { gpb_emu::barrier(); }
static inline int cluster_rank() { return 0; }
static inline void sync_warp() { (void)gpb_emu::shfl_xor_f64(0.0, 0); }   // warp rendezvous
template <int ID, int NTHREADS>
static inline void sync_group() { gpb_emu::group_barrier(NTHREADS); }
static inline double shfl_xor(double v, int m) { return gpb_emu::shfl_xor_f64(v, m); }
static inline float shfl_xor(float v, int m) { return (float)gpb_emu::shfl_xor_f64((double)v, m); }
static inline void atomic_add(double* p, double v) { *p += v; }
// emulated DMMA: same fragment layout as the PTX instruction, operands exchanged lane to lane
static inline void dmma(double& c0, double& c1, double a, double b) {
    double d0, d1;
    gpb_emu::dmma_f64(a, b, &d0, &d1);
    c0 += d0;
    c1 += d1;
}
static inline void cp_async16(void* d, const void* s) { memcpy(d, s, 16); }
static inline void cp_async16_zfill(void* d, const void* s, bool valid) {
    if (valid) memcpy(d, s, 16); else memset(d, 0, 16);
}
static inline void cp_async8_zfill(void* d, const void* s, bool valid) {
    if (valid) memcpy(d, s, 8); else memset(d, 0, 8);
}
static inline void cp_async16_bytes(void* d, const void* s, int nbytes) {
    memset(d, 0, 16);
    if (nbytes > 0) memcpy(d, s, nbytes);
}
static inline void cp_async_commit() {}
template <int N>
static inline void cp_async_wait() {}
static inline double fast_exp(double x) { return exp(x); }
static inline float fast_exp(float x) { return expf(x); }
static inline double ldg(const double* p) { return *p; }
static inline float ldg(const float* p) { return *p; }

}  // namespace gpb
#endif

namespace gpb {

// butterfly all-reduce over the 32 lanes of a warp
template <typename T>
GPB_DEVICE T warp_sum(T v) {
    GPB_UNROLL
    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor(v, m);
    return v;
}

}  // namespace gpb
