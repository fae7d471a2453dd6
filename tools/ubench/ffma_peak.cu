// Pure register FFMA / FFMA2 throughput on sm_100a (no memory): the real ceiling of the conv inner loop.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void ffma2_acc(u64 &acc, u64 a, u64 b) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
template <int KIND>  // 0: FFMA2 scalar-broadcast x, pair w; 1: FFMA2 pair x pair; 2: scalar FFMA
__global__ void k(float *out, int iters, float seed) {
    u64 acc[16];
    float facc[32];
    float xs[8];
    u64 w[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) facc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) xs[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = pack2(seed * i, seed + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            if (KIND == 0) {
                const u64 xx = pack2(xs[ci], xs[ci]);
#pragma unroll
                for (int a = 0; a < 16; ++a) ffma2_acc(acc[a], xx, w[a & 3]);
            } else if (KIND == 1) {
                const u64 xx = pack2(xs[ci], xs[(ci + 1) & 7]);
#pragma unroll
                for (int a = 0; a < 16; ++a) ffma2_acc(acc[a], xx, w[a & 3]);
            } else {
                float wf[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(wf[2 * i]), "=f"(wf[2 * i + 1]) : "l"(w[i]));
#pragma unroll
                for (int a = 0; a < 32; ++a) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(facc[a]) : "f"(xs[ci]), "f"(wf[a & 7]));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[i])); s += a + b; }
#pragma unroll
    for (int i = 0; i < 32; ++i) s += facc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND>
void run(const char *name, int warps) {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<148 * 4, warps * 8>>>(out, 10, 1.f);
    cudaEventRecord(e0);
    k<KIND><<<148 * 4, warps * 8>>>(out, iters, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * 4 * warps * 8 * (double)iters * 8 * 32;
    printf("%-36s warps/SM %2d: %6.2f TFLOP/s (%.1f%% of 74.4)\n", name, warps, 2 * fma / ms / 1e9, 2 * fma / ms / 1e9 / 74.4 * 100);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16}) { run<0>("FFMA2 scalar-x pair-w", w); run<1>("FFMA2 pair-x pair-w", w); run<2>("FFMA scalar", w); }
    return 0;
}
