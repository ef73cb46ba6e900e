// Microbenchmarks for the roofline denominators MEASURED_PEAKS.json does not carry
// (SURVEY.md 8d): FP32 FMA-pipe peak and L2 read bandwidth.
#include <cuda_runtime.h>
#include <stdint.h>
#include "vkrt_internal.h"

namespace vkrt {

// 8 independent FFMA chains per thread, 3-register form (a = a*b + c)
__global__ void __launch_bounds__(256) k_ffma(float *out, int iters, float b, float c)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = __fmaf_rn(a0, b, c); a1 = __fmaf_rn(a1, b, c); a2 = __fmaf_rn(a2, b, c); a3 = __fmaf_rn(a3, b, c);
            a4 = __fmaf_rn(a4, b, c); a5 = __fmaf_rn(a5, b, c); a6 = __fmaf_rn(a6, b, c); a7 = __fmaf_rn(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

cudaError_t measure_fp32_peak(float *tflops)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    float *out = nullptr;
    cudaError_t e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 0.f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_ffma<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        const float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops = best;
    return e;
}

// every thread streams float4s from a 32 MiB buffer (L2-resident on B200's 126 MB L2) many times
__global__ void __launch_bounds__(256) k_l2read(const float4 *__restrict__ buf, size_t n, int reps, float *out)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
            const float4 v = __ldcg(buf + i);
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) out[0] = acc;
}

cudaError_t measure_l2_bandwidth(float *gbs)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t bytes = (size_t)32 << 20, n = bytes / sizeof(float4);
    float4 *buf = nullptr; float *out = nullptr;
    cudaError_t e = cudaMalloc(&buf, bytes);
    if (e != cudaSuccess) return e;
    cudaMalloc(&out, 4);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20;
    float best = 0.f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_l2read<<<sms * 8, 256>>>(buf, n, reps, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const float g = (float)((double)bytes * reps / (ms * 1e-3) / 1e9);
        if (rep > 0 && g > best) best = g;
    }
    e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf); cudaFree(out);
    *gbs = best;
    return e;
}

} // namespace vkrt
