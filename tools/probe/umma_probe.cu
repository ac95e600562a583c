// tcgen05 (UMMA) probe: C[128 x N] = A[128 x K] * B[N x K]^T, TF32 operands, fp32 accumulation in TMEM.
// Operands are written to shared memory by ordinary stores in the canonical K-major, no-swizzle layout
//   element (r, k) at  (k / 4) * (R * 16) + r * 16 + (k % 4) * 4  bytes      (R = rows of the operand)
// i.e. "core matrices" of 8 rows x 16 bytes, SBO = 128 B between 8-row groups, LBO = R * 16 B between
// the two 16-byte K chunks of one K = 8 instruction.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu ; run: ./umma_probe [N] [K] [mode]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);               // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
    return d;                                               // layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128) umma_probe(const float* __restrict__ A, const float* __restrict__ B,
                                                  float* __restrict__ C, int N, int K, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* sA = (float*)smem;                       // [K/4][128][4]
    float* sB = sA + 128 * K;                       // [K/4][N][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    // mode 0: K-major (above).  mode 1 / 2: MN-major, no swizzle: core matrix = 8 k x 16 bytes (4 rows),
    //   element (r, k) at ((k / 8) * (R / 4) + r / 4) * 128 + (k % 8) * 16 + (r % 4) * 4 bytes;
    //   mode 1 passes the 128-byte distance between 4-row groups as SBO, mode 2 as LBO.
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        if (mode == 0) sA[(k / 4) * 128 * 4 + r * 4 + (k % 4)] = A[r * K + k];
        else sA[((k / 8) * 32 + r / 4) * 32 + (k % 8) * 4 + (r % 4)] = A[r * K + k];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        if (mode == 0) sB[(k / 4) * N * 4 + r * 4 + (k % 4)] = B[r * K + k];
        else sB[((k / 8) * (N / 4) + r / 4) * 32 + (k % 8) * 4 + (r % 4)] = B[r * K + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");        // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        if (mode) idesc |= (1u << 15) | (1u << 16);            // a_major = b_major = MN
        for (int ks = 0; ks < K / 8; ks++) {
            uint64_t da = make_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
            uint64_t db = make_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
            if (mode == 1) {
                da = make_desc(smem_u32(sA) + ks * 32 * 128, 32 * 128, 128);
                db = make_desc(smem_u32(sB) + ks * (N / 4) * 128, (N / 4) * 128, 128);
            } else if (mode == 2) {
                da = make_desc(smem_u32(sA) + ks * 32 * 128, 128, 32 * 128);
                db = make_desc(smem_u32(sB) + ks * (N / 4) * 128, 128, (N / 4) * 128);
            }
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    // epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31; thread = row
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) C[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase));
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 64, K = argc > 2 ? atoi(argv[2]) : 32, mode = argc > 3 ? atoi(argv[3]) : 0;
    std::vector<float> A(128 * K), B((size_t)N * K), C((size_t)128 * N);
    srand(1);
    for (auto& x : A) x = (float)(rand() % 17 - 8) / 8.0f;        // exactly representable in TF32
    for (auto& x : B) x = (float)(rand() % 13 - 6) / 4.0f;
    float *dA, *dB, *dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, C.size() * 4);
    const size_t smem = (size_t)(128 + N) * K * 4;
    cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_probe<<<1, 128, smem>>>(dA, dB, dC, N, K, mode);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    int bad = 0;
    for (int r = 0; r < 128; r++)
        for (int c = 0; c < N; c++) {
            double ref = 0;
            for (int k = 0; k < K; k++) ref += (double)A[r * K + k] * B[c * K + k];
            double d = fabs(ref - C[(size_t)r * N + c]);
            if (d > worst) worst = d;
            if (d > 1e-4 && bad < 8) { printf("  mismatch r=%d c=%d got %g ref %g\n", r, c, C[(size_t)r * N + c], ref); bad++; }
        }
    printf("mode=%d N=%d K=%d worst abs err %.3g -> %s\n", mode, N, K, worst, worst < 1e-4 ? "OK" : "FAIL");
    return worst < 1e-4 ? 0 : 1;
}
