// Microbenchmark: sustained FFMA / DFMA / MUFU-double rates on this GPU (design input for the fp64 core).
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
    T x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = x[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ddiv_kernel(double* out, int iters, double a) {
    double x = 1.0 + threadIdx.x, y = 2.0 + threadIdx.x;
    for (int it = 0; it < iters; ++it) { x = a / x + 1.0; y = a / y + 2.0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}
__global__ void datan_kernel(double* out, int iters, double a) {
    double x = 0.1 + 1e-3 * threadIdx.x;
    for (int it = 0; it < iters; ++it) x = atan2(x, a) + 0.3;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    void* buf; cudaMalloc(&buf, (size_t)blocks * threads * 8);
    float ms;
    ms = timeit([&] { fma_kernel<float, 8><<<blocks, threads>>>((float*)buf, iters, 1.0001f, 0.5f); });
    printf("FFMA  : %.2f TFLOP/s\n", 2.0 * blocks * threads * (double)iters * 8 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { fma_kernel<double, 8><<<blocks, threads>>>((double*)buf, iters, 1.0001, 0.5); });
    printf("DFMA  : %.2f TFLOP/s  (%.3f ms)\n", 2.0 * blocks * threads * (double)iters * 8 / (ms * 1e-3) / 1e12, ms);
    ms = timeit([&] { fma_kernel<double, 1><<<blocks, threads>>>((double*)buf, iters, 1.0001, 0.5); });
    printf("DFMA ilp1 (8 warps/SMSP): %.2f TFLOP/s\n", 2.0 * blocks * threads * (double)iters / (ms * 1e-3) / 1e12);
    ms = timeit([&] { ddiv_kernel<<<blocks, threads>>>((double*)buf, 512, 3.0); });
    printf("DDIV  : %.2f G div/s\n", 2.0 * blocks * threads * 512.0 / (ms * 1e-3) / 1e9);
    ms = timeit([&] { datan_kernel<<<blocks, threads>>>((double*)buf, 256, 3.0); });
    printf("atan2(double): %.2f G/s\n", 1.0 * blocks * threads * 256.0 / (ms * 1e-3) / 1e9);
    return 0;
}
