// mma.sync throughput on sm_100a: m16n8k8 tf32 and m16n8k16 bf16, fp32 accumulate
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int KIND, int ILP>
__global__ void k(float *out, int iters) {
    float c[ILP][4];
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f900000u, 0x3fa00000u, 0x3fb00000u}, b[2] = {0x3f800000u, 0x3f880000u + threadIdx.x};
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int ILP>
void run(const char *name, int warps) {
    float *out; cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, ILP><<<148, warps * 32>>>(out, 10);
    cudaEventRecord(e0);
    k<KIND, ILP><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = 148.0 * warps * iters * ILP;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%-34s warps/SM %2d: %.2f cycles per MMA per SM sub-partition, %.1f TFLOP/s\n", name, warps, cyc / (mmas / 148 / 4), mmas * (KIND == 0 ? 1024 : 2048) * 2 / ms / 1e9);
    cudaFree(out);
}
int main() {
    run<0, 8>("m16n8k8 tf32 ILP8", 4);
    run<0, 8>("m16n8k8 tf32 ILP8", 16);
    run<1, 8>("m16n8k16 bf16 ILP8", 4);
    run<1, 8>("m16n8k16 bf16 ILP8", 16);
    return 0;
}
