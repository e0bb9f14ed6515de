// dependent-chain latency of DFMA/DADD/FFMA and F2F conversions (one warp), in SM clocks
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void chain(T* out, long long* cyc, int n, T a, T b) {
    T x = (T)threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x = x * a + b; x = x * a + b; x = x * a + b; x = x * a + b; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_add(double* out, long long* cyc, int n, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x = x + b; x = x + b; x = x + b; x = x + b; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_cvt(double* out, long long* cyc, int n, double b) {
    double x = threadIdx.x + 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { float f = (float)x; x = (double)f + b; f = (float)x; x = (double)f + b; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    long long h; const int n = 4096;
    chain<double><<<1, 32>>>(d, c, n, 1.0000001, 0.5); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency: %.1f cycles\n", (double)h / (4.0 * n));
    chain<float><<<1, 32>>>((float*)d, c, n, 1.0000001f, 0.5f); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("FFMA dependent latency: %.1f cycles\n", (double)h / (4.0 * n));
    chain_add<<<1, 32>>>(d, c, n, 0.5); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DADD dependent latency: %.1f cycles\n", (double)h / (4.0 * n));
    chain_cvt<<<1, 32>>>(d, c, n, 0.5); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("F2F(d->f)+F2F(f->d)+DADD chain: %.1f cycles per triple\n", (double)h / (2.0 * n));
    return 0;
}
