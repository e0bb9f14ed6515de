// Issue rate of the legacy warp-level mma.sync path on this GPU (tools, not product): how many m16n8k8 TF32 / m16n8k16 F16
// instructions one SM retires per clock with W warps resident.  Decides between 3xTF32 and a 2-term F16 split for the policy MLP.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND> __global__ void rate(float* out, int iters, long long* cyc) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-9f;
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    const char* names[3] = {"tf32 m16n8k8", "f16 m16n8k16", "bf16 m16n8k16"};
    for (int kind = 0; kind < 3; ++kind)
        for (int warps = 4; warps <= 32; warps *= 2) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) rate<0><<<148, warps * 32>>>(out, iters, cyc);
                else if (kind == 1) rate<1><<<148, warps * 32>>>(out, iters, cyc);
                else rate<2><<<148, warps * 32>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double per_sm_clk = (double)iters * 8 * warps / (double)h;
            const double flops = per_sm_clk * (kind == 0 ? 2048 : 4096);
            printf("{\"kind\": \"%s\", \"warps_per_sm\": %d, \"mma_per_sm_clk\": %.4f, \"cycles_per_mma_per_smsp\": %.2f, \"flop_per_sm_clk\": %.0f}\n",
                   names[kind], warps, per_sm_clk, 4.0 / per_sm_clk, flops);
        }
    return 0;
}
