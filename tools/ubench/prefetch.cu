// Does prefetch.global.L1 / .L2 hide latency on sm_100a?  One warp per SM, dependent use right after each load.
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>  // 0 none, 1 prefetch.L1, 2 prefetch.L2, 3 real load into rotating regs ("touch")
__global__ void k(const float *__restrict__ x, size_t stride_f, int iters, int dist, float *out, long long *cyc) {
    const float *p = x + (size_t)blockIdx.x * stride_f * (iters + dist + 8) + threadIdx.x * 8;
    float acc = 0.f;
    float t0 = 0.f;
    long long c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        const float *q = p + (size_t)i * stride_f;
        if (MODE == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(q + (size_t)dist * stride_f));
        if (MODE == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + (size_t)dist * stride_f));
        if (MODE == 3) {
            float d;
            asm volatile("ld.global.f32 %0, [%1];" : "=f"(d) : "l"(q + (size_t)dist * stride_f));
            t0 += d * 0.f;  // consumed one iteration later at the earliest (scoreboard), keeps the load alive
        }
        float v[8];
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(q));
        acc += v[0] + v[7];
    }
    long long c1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + t0;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
int main() {
    const int nb = 148, iters = 2000, dist = 8;
    const size_t stride_f = 32 * 8;  // one warp-step = 1 KB contiguous (32 rows x 32 B)
    size_t n = (size_t)nb * stride_f * (iters + dist + 8);
    float *x, *out; long long *cyc;
    cudaMalloc(&x, n * 4); cudaMemset(x, 0, n * 4);
    cudaMalloc(&out, nb * 32 * 4); cudaMalloc(&cyc, nb * 8);
    // flush buffer > L2
    char *fl; cudaMalloc(&fl, 512 << 20);
    auto run = [&](auto kern, const char *name) {
        cudaMemset(fl, 1, 512 << 20);
        kern<<<nb, 32>>>(x, stride_f, iters, dist, out, cyc);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < nb; ++i) s += h[i];
        printf("%-28s %8.1f cycles / iteration   (%s)\n", name, s / nb / iters, cudaGetErrorString(cudaGetLastError()));
    };
    run(k<0>, "no prefetch (DRAM)");
    run(k<1>, "prefetch.global.L1 dist 8");
    run(k<2>, "prefetch.global.L2 dist 8");
    run(k<3>, "touch load dist 8");
    // L2-resident: run twice without flushing (array 148*2016*1KB = 305 MB > L2, so use a smaller one)
    return 0;
}
