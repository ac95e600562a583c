// Dependent-issue latencies on sm_100a that bound the serial pivot chain of the SPD inversion:
// DFMA / DMUL / MUFU.RCP64H chains, shared-memory round trip, named barrier over 4 warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_kernel(double* out, long long* cyc, double x0) {
    __shared__ double sh[256];
    const int tid = threadIdx.x;
    constexpr int N = 512;
    double a = x0 + tid * 1e-9, b = 0.999999, c = 1e-7;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = fma(a, b, c);
    t1 = clock64();
    if (tid == 0) cyc[0] = (t1 - t0);
    // 2. dependent DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = a * b;
    t1 = clock64();
    if (tid == 0) cyc[1] = (t1 - t0);
    // 3. dependent MUFU.RCP64H + 2 DFMA (one Newton step on the high-word estimate)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
        a = r + 1.5;
    }
    t1 = clock64();
    if (tid == 0) cyc[2] = (t1 - t0);
    // 4. shared-memory round trip: store, load dependent
    sh[tid] = a;
    __syncthreads();
    t0 = clock64();
    int idx = tid;
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        double v = sh[idx];
        idx = (int)(v * 0.0) + ((idx + 32) & 255);      // address depends on the loaded value
    }
    t1 = clock64();
    if (tid == 0) cyc[3] = (t1 - t0);
    a += idx;
    // 5. named barrier over the block's warps
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
    t1 = clock64();
    if (tid == 0) cyc[4] = (t1 - t0);
    // 6. store -> barrier -> load chain (one pivot's data hand-over)
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        sh[(tid + i) & 255] = a;
        asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
        a = sh[(tid + i + 33) & 255] + 1.0;
    }
    t1 = clock64();
    if (tid == 0) cyc[5] = (t1 - t0);
    // 7. FFMA chain for comparison
    float f = (float)a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fmaf(f, 0.99999f, 1e-7f);
    t1 = clock64();
    if (tid == 0) cyc[6] = (t1 - t0);
    out[tid] = a + f;
    if (tid == 0) cyc[7] = N;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 4096 * sizeof(double));
    cudaMallocManaged(&cyc, 8 * sizeof(long long));
    const char* names[7] = {"DFMA dependent", "DMUL dependent", "MUFU.RCP64H + DADD dependent", "LDS dependent (address)",
                            "bar.sync (named)", "STS -> bar.sync -> LDS -> DADD", "FFMA dependent"};
    for (int nt = 32; nt <= 128; nt *= 4) {
        lat_kernel<<<1, nt>>>(out, cyc, 1.0);
        cudaDeviceSynchronize();
        lat_kernel<<<1, nt>>>(out, cyc, 1.0);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
        printf("threads per block %d\n", nt);
        for (int i = 0; i < 7; i++) printf("  %-34s %7.1f cycles / iteration\n", names[i], (double)cyc[i] / (double)cyc[7]);
    }
    return 0;
}
