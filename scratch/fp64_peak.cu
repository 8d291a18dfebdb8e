// scratch microbenchmark: FP64 DFMA and DMMA (mma.sync f64) peak on this GPU
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void dmma884_kernel(double *out, int iters, double a0, double b0)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

// m16n8k4: A 2 regs, B 1 reg, C 4 regs ; m16n8k8: A 4, B 2, C 4 ; m16n8k16: A 8, B 4, C 4
template <int NACC>
__global__ void dmma16816_kernel(double *out, int iters, double a0, double b0)
{
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void dmma1684_kernel(double *out, int iters, double a0, double b0)
{
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0; }
    double a[2] = {a0 + threadIdx.x * 1e-9, a0 + 1}, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// mixed: DMMA + independent DFMA in the same warp: do the pipes overlap?
template <int NACC, int NF>
__global__ void mixed_kernel(double *out, int iters, double a0, double b0)
{
    double c[NACC][2], v[NF];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
    for (int i = 0; i < NF; ++i) v[i] = threadIdx.x + i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
        for (int i = 0; i < NF; ++i) v[i] = fma(v[i], a0, b0);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;
}

template <class F>
float time_it(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double *out; cudaMalloc(&out, 64);
    const int iters = 20000;
    for (int wpb : {4, 8, 16, 32}) {
        const int threads = wpb * 32, blocks = sms;
        float ms = time_it([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 8 * iters * (double)threads * blocks;
        printf("DFMA ilp8  warps/SM %2d: %.2f TFLOP/s\n", wpb, fl / ms / 1e9);
        ms = time_it([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 512.0 * 8 * iters * (double)wpb * blocks;
        printf("DMMA m8n8k4 x8acc  warps/SM %2d: %.2f TFLOP/s\n", wpb, fl / ms / 1e9);
        ms = time_it([&] { dmma1684_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 1024.0 * 4 * iters * (double)wpb * blocks;
        printf("DMMA m16n8k4 x4acc warps/SM %2d: %.2f TFLOP/s\n", wpb, fl / ms / 1e9);
        ms = time_it([&] { dmma16816_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 4096.0 * 4 * iters * (double)wpb * blocks;
        printf("DMMA m16n8k16 x4acc warps/SM %2d: %.2f TFLOP/s\n", wpb, fl / ms / 1e9);
        ms = time_it([&] { mixed_kernel<8, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = (512.0 * 8 * wpb + 2.0 * 8 * threads) * iters * (double)blocks;
        printf("mixed 8 DMMA + 8 DFMA  warps/SM %2d: %.2f TFLOP/s total (%.3f ms)\n", wpb, fl / ms / 1e9, ms);
    }
    return 0;
}
