// peaks.cu -- FP64 peak microbenchmarks for the roofline denominators of the assembly kernels ("measure
// achieved peaks on the box").  Two dependency-free register loops: DMMA (mma.sync.m8n8k4.f64, the FP64 tensor pipe
// the block Gauss-Jordan runs on) and scalar DFMA.  One persistent wave: 148 SMs x resident CTAs.
#include "plan.hpp"

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double *sink) {
    double c[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) { c[t][0] = threadIdx.x * 1e-9; c[t][1] = t * 1e-9; }
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 8; ++t) pb_dmma(c[t], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
    if (s == 123.456) sink[0] = s;  // never true: keeps the loop alive
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double *sink) {
    double x[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) x[t] = threadIdx.x * 1e-9 + t;
    const double a = 1.0 - 1e-12, b = 1e-13;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 8; ++t) x[t] = fma(x[t], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += x[t];
    if (s == 123.456) sink[0] = s;
}

// kind 0: DMMA (512 flops per warp instruction), kind 1: DFMA (2 flops per thread instruction).  Best of 5 launches.
extern "C" int pb_fp64_peak(int kind, double *tflops) {
    if (!tflops || kind < 0 || kind > 1) return pb_fail_(PB_EINVAL, "bad arguments");
    DevBuf sink;
    CUDA_TRY(sink.ensure(8));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 1 << 15, block = 256, grid = kSMs * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, 0));
        if (kind == 0) dmma_peak_kernel<<<grid, block>>>(iters, sink.as<double>());
        else dfma_peak_kernel<<<grid, block>>>(iters, sink.as<double>());
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        pb_count_launch_();
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double per_thread_instr = (double)iters * 8;
        const double flops = kind == 0 ? per_thread_instr * (grid * (double)block / 32) * 512.0
                                       : per_thread_instr * (grid * (double)block) * 2.0;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return PB_OK;
}
