// Micro-benchmark: per-SM throughput of the FP64 operations the smoothing kernels are made of
// (B200, sm_100a).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false fp64_rates.cu -o fp64_rates
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP> __global__ void k(double *out, double a0, double b0)
{
    double x[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        x[i] = a0 + i * 1e-3 + threadIdx.x * 1e-6;
        f[i] = (float)x[i];
    }
    for (int it = 0; it < ITERS; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            if (OP == 0) x[i] = x[i] + b0;                          // DADD
            if (OP == 1) x[i] = x[i] * b0;                          // DMUL
            if (OP == 2) x[i] = fma(x[i], b0, a0);                  // DFMA
            if (OP == 3) x[i] = x[i] / b0;                          // IEEE division
            if (OP == 4) x[i] = sqrt(x[i]) + a0;                    // IEEE sqrt (+1 add)
            if (OP == 5) { f[i] = (float)x[i]; x[i] = x[i] + (double)f[i]; }  // F2F both ways + DADD
            if (OP == 6) f[i] = fmaf(f[i], 1.0001f, 0.5f);          // FFMA
            if (OP == 7) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i])); x[i] = r + a0; }
            if (OP == 8) { f[i] = (float)x[i]; x[i] = x[i] + b0; }  // one F2F.F32.F64 + DADD
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s += x[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char *name, double opsPerIter)
{
    const int blocks = 148 * 8, threads = 256;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 1.5, 1.000001);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 1.5, 1.000001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double)blocks * threads * ITERS * 8 * opsPerIter;
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-34s %8.3f ms  %8.2f Gop/s  %6.2f op/clk/SM (at %d MHz)\n", name, ms, total / ms * 1e-6,
           total / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}

int main()
{
    run<0>("DADD", 1);
    run<1>("DMUL", 1);
    run<2>("DFMA", 1);
    run<3>("IEEE f64 division", 1);
    run<4>("IEEE f64 sqrt (+DADD)", 1);
    run<5>("F2F f64->f32->f64 (+DADD)", 1);
    run<8>("F2F f64->f32 (+DADD)", 1);
    run<6>("FFMA", 1);
    run<7>("rsqrt.approx.f64 (+DADD)", 1);
    return 0;
}
