// dp_probe.cu -- development microbenchmark (not part of the library): what does the FP64 pipe of
// one SM sustain for instruction streams that look like the pair kernel (distinct register
// operands, DADD/DMUL/DFMA mixes, few resident warps)?   nvcc -arch=sm_100a -O3 -o dp_probe dp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void __launch_bounds__(256) probe(long iters, double* sink, const double* src) {
    double a[8], b[8], c[8];
    int iv[8];
    __shared__ int sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = (i * 7) & 1023;
    for (int i = 0; i < 8; i++) iv[i] = threadIdx.x * 3 + i;
    __syncthreads();
    for (int i = 0; i < 8; i++) {
        a[i] = src[threadIdx.x + i];
        b[i] = src[threadIdx.x + 8 + i] * 1e-3 + 0.999;
        c[i] = src[threadIdx.x + 16 + i] * 1e-7;
    }
    for (long it = 0; it < iters; it++) {
        if (V == 0) {            // 8 chains, shared multiplier/addend
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b[0] + c[0];
        } else if (V == 1) {     // 8 chains, three distinct register operands each
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b[i] + c[i];
        } else if (V == 2) {     // distinct operands, operands rotate between chains
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[(i + 3) & 7] * b[(i + 5) & 7] + c[i];
        } else if (V == 3) {     // add / mul / fma mix 1:1:2, distinct operands
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
                a[i] = a[i] + c[i];
                a[i + 1] = a[i + 1] * b[i + 1];
                a[i + 2] = a[i + 2] * b[i + 2] + c[i + 2];
                a[i + 3] = a[i + 3] * b[i + 3] + c[i + 3];
            }
        } else if (V == 4) {     // DADD only
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] + c[i];
        } else if (V == 5) {     // DMUL only
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b[i];
        } else if (V == 6) {     // 4 chains only (ILP 4), distinct operands
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = a[i] * b[i] + c[i];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = a[i] * b[i + 4] + c[i + 4];
        } else if (V == 7) {     // 2 chains only (ILP 2)
#pragma unroll
            for (int i = 0; i < 4; i++) { a[0] = a[0] * b[i] + c[i]; a[1] = a[1] * b[i + 4] + c[i + 4]; }
        } else if (V == 8) {     // shared multiplier only
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b[0] + c[i];
        } else if (V == 9) {     // shared addend only
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b[i] + c[0];
        } else if (V == 10) {    // accumulate form, shared multiplicand: a[i] += b0 * c[i]
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = b[0] * c[i] + a[i];
        } else if (V == 11) {    // accumulate form, operand shared by PAIRS of consecutive FMAs
#pragma unroll
            for (int i = 0; i < 8; i += 2) { a[i] = b[i] * c[i] + a[i]; a[i + 1] = b[i] * c[i + 1] + a[i + 1]; }
        } else if (V == 12) {    // accumulate form, operand shared by groups of 4
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
                a[i] = b[i] * c[i] + a[i]; a[i + 1] = b[i] * c[i + 1] + a[i + 1];
                a[i + 2] = b[i] * c[i + 2] + a[i + 2]; a[i + 3] = b[i] * c[i + 3] + a[i + 3];
            }
        } else if (V == 13) {    // constant-bank addend: a = a*b + K
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b[i] + 1.0000001e-7;
        } else if (V == 14) {    // accumulate form, all distinct: a[i] += b[i]*c[i]
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = b[i] * c[i] + a[i];
        } else if (V == 16 || V == 17 || V == 18 || V == 19) {
            // fast-path DFMA (shared operands) + one non-DP instruction per DFMA:
            // 16: IMAD (FMA-heavy pipe)  17: LOP3 (ALU)  18: LDS  19: two IMADs per DFMA
#pragma unroll
            for (int i = 0; i < 8; i++) {
                a[i] = a[i] * b[0] + c[0];
                if (V == 16) iv[i] = iv[i] * iv[(i + 1) & 7] + 12345;
                if (V == 19) { iv[i] = iv[i] * iv[(i + 1) & 7] + 12345; iv[(i + 3) & 7] = iv[(i + 3) & 7] * iv[i] + 77; }
                if (V == 17) iv[i] = (iv[i] & iv[(i + 1) & 7]) ^ 0x5a5a5a5a;
                if (V == 18) iv[i] = sm[(iv[i] + threadIdx.x) & 1023];
            }
        } else if (V == 20) {    // conversions only: 8 x (F2F.F32.F64 + F2F.F64.F32) per iteration
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = (double)((float)a[i] * 1.0000001f);
        } else if (V == 21) {    // fast DFMA + one F64->F32->F64 round trip per DFMA
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = a[i] * b[0] + c[0]; b[i] = (double)((float)b[i] * 1.0000001f); }
        } else if (V == 22) {    // 8 fast DFMA + two round trips (the ratio of a table exp with an fp32 tail)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b[0] + c[0];
            c[1] = (double)((float)c[1] * 1.0000001f);
            c[2] = (double)((float)c[2] * 1.0000001f);
        } else if (V == 15) {    // 2x4 outer product: a[i*4+j] += b[i]*c[j]
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) a[i * 4 + j] = b[i] * c[j] + a[i * 4 + j];
        }
    }
    double r = 0;
    for (int i = 0; i < 8; i++) r += a[i] + b[i] + c[i] + (double)iv[i];
    if (r == 1.2345) sink[blockIdx.x] = r;
}

template <int V>
void run(const char* name, int sms, double* sink, double* src, double clk_ghz) {
    const long iters = 20000;
    for (int bps : {1, 4}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe<V><<<sms * bps, 256>>>(iters, sink, src);
        cudaDeviceSynchronize();
        float best = 1e9f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            probe<V><<<sms * bps, 256>>>(iters, sink, src);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double winst = (double)bps * 8 /*warps*/ * iters * 8.0;           // warp instr per SM
        double per_clk = winst / (best * 1e-3 * clk_ghz * 1e9);
        printf("%-28s blocks/SM=%d warps/SMSP=%2d  %.3f ms  %.3f DP warp-instr/clk/SM (peak 2.0) -> %.0f%%\n",
               name, bps, bps * 2, best, per_clk, 100 * per_clk / 2.0);
    }
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    double ghz = khz * 1e-6;
    printf("SMs %d, clock %.3f GHz (nominal max; percentages assume it)\n", sms, ghz);
    double *sink, *src;
    cudaMalloc(&sink, sizeof(double) * sms * 64);
    cudaMalloc(&src, sizeof(double) * 1024);
    cudaMemset(src, 0, sizeof(double) * 1024);
    run<0>("fma shared operands", sms, sink, src, ghz);
    run<1>("fma distinct operands", sms, sink, src, ghz);
    run<2>("fma rotating operands", sms, sink, src, ghz);
    run<3>("add:mul:fma 1:1:2", sms, sink, src, ghz);
    run<4>("dadd only", sms, sink, src, ghz);
    run<5>("dmul only", sms, sink, src, ghz);
    run<6>("fma ILP4", sms, sink, src, ghz);
    run<7>("fma ILP2", sms, sink, src, ghz);
    run<8>("a=a*B0+c[i] shared mult", sms, sink, src, ghz);
    run<9>("a=a*b[i]+C0 shared addend", sms, sink, src, ghz);
    run<10>("a+=B0*c[i] shared A", sms, sink, src, ghz);
    run<11>("a+=b*c shared by pairs", sms, sink, src, ghz);
    run<12>("a+=b*c shared by 4", sms, sink, src, ghz);
    run<13>("a=a*b+const", sms, sink, src, ghz);
    run<14>("a+=b[i]*c[i] distinct", sms, sink, src, ghz);
    run<15>("2x4 outer product", sms, sink, src, ghz);
    run<16>("fast DFMA + 1 IMAD each", sms, sink, src, ghz);
    run<19>("fast DFMA + 2 IMAD each", sms, sink, src, ghz);
    run<17>("fast DFMA + 1 LOP3 each", sms, sink, src, ghz);
    run<18>("fast DFMA + 1 LDS each", sms, sink, src, ghz);
    run<20>("8 F64->F32->F64 only (as 8)", sms, sink, src, ghz);
    run<21>("fast DFMA + 1 cvt pair each", sms, sink, src, ghz);
    run<22>("8 fast DFMA + 2 cvt pairs", sms, sink, src, ghz);
    return 0;
}
