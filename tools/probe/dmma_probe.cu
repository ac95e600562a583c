// dmma_probe.cu -- development microbenchmark: sustained rate of mma.sync.m8n8k4.f64 (DMMA) on one
// SM vs the FP64 FMA pipe.   nvcc -arch=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) probe(long iters, double* sink, const double* src) {
    double c[NACC][2], a[4], b[4];
    for (int i = 0; i < NACC; i++) { c[i][0] = src[threadIdx.x + i]; c[i][1] = src[threadIdx.x + 32 + i]; }
    for (int i = 0; i < 4; i++) { a[i] = src[threadIdx.x + 64 + i] + 1e-3; b[i] = src[threadIdx.x + 96 + i] + 1e-3; }
    for (long it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a[i & 3], b[(i >> 2) & 3]);
    }
    double r = 0;
    for (int i = 0; i < NACC; i++) r += c[i][0] + c[i][1];
    if (r == 1.2345) sink[blockIdx.x] = r;
}

template <int NACC>
void run(int sms, double* sink, double* src, double ghz) {
    const long iters = 4000;
    for (int bps : {1, 2, 4}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe<NACC><<<sms * bps, 256>>>(iters, sink, src);
        cudaDeviceSynchronize();
        float best = 1e9f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            probe<NACC><<<sms * bps, 256>>>(iters, sink, src);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double flops = (double)sms * bps * 8 * iters * NACC * 512.0;   // 8*8*4*2 per warp DMMA
        printf("DMMA m8n8k4 x%2d accumulators, blocks/SM=%d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at %.3f GHz)\n",
               NACC, bps, best, flops / (best * 1e-3) / 1e12, flops / 2 / (best * 1e-3) / (ghz * 1e9) / sms, ghz);
    }
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    double *sink, *src;
    cudaMalloc(&sink, sizeof(double) * sms * 64);
    cudaMalloc(&src, sizeof(double) * 1024);
    cudaMemset(src, 0, sizeof(double) * 1024);
    run<4>(sms, sink, src, khz * 1e-6);
    run<8>(sms, sink, src, khz * 1e-6);
    run<16>(sms, sink, src, khz * 1e-6);
    return 0;
}
