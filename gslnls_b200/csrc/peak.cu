// peak.cu -- measurement hooks for bench.py's roofline denominators (not on the product path):
//   * FP64-pipe issue peak of this GPU, as dependent-free DFMA chains and as mma.sync.m8n8k4.f64 (DMMA):
//     the bound of the tiled pass kernel (K1b).  MEASURED_PEAKS.json carries HBM and bf16 only.
//   * plain read bandwidth of a buffer (sum of doubles, 16-byte loads): what a pass kernel could at best reach.
// Timed with CUDA events on a private stream, best of `reps`.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "../../include/gslnls_b200.h"

namespace gslnls {
void set_error(const std::string &s);

template <int ILP>
__global__ void __launch_bounds__(256) peak_dfma(double *out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
        v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i)
        s += v[i];
    if (s == 12345.678)
        out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) peak_dmma(double *out, int iters, double a0, double b0)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
        c[i][0] = c[i][1] = 0.0;
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i)
        s += c[i][0] + c[i][1];
    if (s == 12345.678)
        out[0] = s;
}

__global__ void __launch_bounds__(512) peak_read(const double2 *in, size_t n2, double *out)
{
    double s = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        const double2 a = __ldcs(in + i), b = __ldcs(in + i + stride), c = __ldcs(in + i + 2 * stride),
                      d = __ldcs(in + i + 3 * stride);
        s += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
    }
    for (; i < n2; i += stride) {
        const double2 a = __ldcs(in + i);
        s += a.x + a.y;
    }
    if (s == 12345.678)
        out[0] = s;
}
} // namespace gslnls
using namespace gslnls;

#define CKP(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                            \
            return GSLNLS_ECUDA;                                                                       \
        }                                                                                              \
    } while (0)

extern "C" GSLNLS_API int gslnls_measure_fp64_peak(int device, double *dfma_tflops, double *dmma_tflops)
{
    if (gslnls_device_count() <= device) {
        set_error("no usable CUDA device");
        return GSLNLS_ENODEVICE;
    }
    CKP(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKP(cudaGetDeviceProperties(&prop, device));
    cudaStream_t st;
    cudaEvent_t e0, e1;
    CKP(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CKP(cudaEventCreate(&e0));
    CKP(cudaEventCreate(&e1));
    double *out = nullptr;
    CKP(cudaMalloc(&out, 8));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double best[2] = {0.0, 0.0};
    for (int which = 0; which < 2; ++which)
        for (int rep = 0; rep < 6; ++rep) {
            CKP(cudaEventRecord(e0, st));
            if (which == 0)
                peak_dfma<8><<<blocks, threads, 0, st>>>(out, iters, 0.999999, 1e-9);
            else
                peak_dmma<8><<<blocks, threads, 0, st>>>(out, iters, 0.5, 0.25);
            CKP(cudaEventRecord(e1, st));
            CKP(cudaEventSynchronize(e1));
            float ms = 0.f;
            CKP(cudaEventElapsedTime(&ms, e0, e1));
            // DFMA: 2 flop per thread-instruction; DMMA m8n8k4: 8*8*4*2 = 512 flop per warp-instruction
            const double flop = which == 0 ? 2.0 * 8 * iters * (double)blocks * threads
                                           : 512.0 * 8 * iters * (double)blocks * (threads / 32);
            if (rep > 0)
                best[which] = std::max(best[which], flop / (ms * 1e-3) / 1e12);
        }
    if (dfma_tflops)
        *dfma_tflops = best[0];
    if (dmma_tflops)
        *dmma_tflops = best[1];
    cudaFree(out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return GSLNLS_SUCCESS;
}

extern "C" GSLNLS_API int gslnls_measure_read_bandwidth(int device, size_t bytes, double *gb_per_s)
{
    if (!gb_per_s || bytes < 1024)
        return GSLNLS_EINVAL;
    if (gslnls_device_count() <= device) {
        set_error("no usable CUDA device");
        return GSLNLS_ENODEVICE;
    }
    CKP(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKP(cudaGetDeviceProperties(&prop, device));
    cudaStream_t st;
    cudaEvent_t e0, e1;
    CKP(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CKP(cudaEventCreate(&e0));
    CKP(cudaEventCreate(&e1));
    double *buf = nullptr, *out = nullptr;
    CKP(cudaMalloc(&buf, bytes));
    CKP(cudaMalloc(&out, 8));
    CKP(cudaMemsetAsync(buf, 0, bytes, st));
    double best = 0.0;
    for (int rep = 0; rep < 8; ++rep) {
        CKP(cudaEventRecord(e0, st));
        peak_read<<<prop.multiProcessorCount * 4, 512, 0, st>>>((const double2 *)buf, bytes / 16, out);
        CKP(cudaEventRecord(e1, st));
        CKP(cudaEventSynchronize(e1));
        float ms = 0.f;
        CKP(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 1)
            best = std::max(best, (double)bytes / (ms * 1e-3) / 1e9);
    }
    *gb_per_s = best;
    cudaFree(buf);
    cudaFree(out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return GSLNLS_SUCCESS;
}
