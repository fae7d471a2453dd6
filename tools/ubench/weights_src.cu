// microbenchmark: weight-operand source for lane=row FFMA2 conv inner loop
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void ffma2_acc(u64 &acc, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__constant__ __align__(16) float cw[8][27 * 64];
__device__ __forceinline__ void gather8(const float *p, float (&v)[8]) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
template <int RPT, int SRC>  // SRC 0: smem LDS.128, 1: constant (uniform regs)
__global__ void __launch_bounds__(128) kern(const float *__restrict__ x, const float *__restrict__ wg, float *__restrict__ y, int n) {
    __shared__ __align__(16) float s_w[27 * 64];
    const int g = blockIdx.y;
    if (SRC == 0) {
        for (int i = threadIdx.x; i < 27 * 64; i += 128) s_w[i] = wg[g * 27 * 64 + i];
        __syncthreads();
    }
    int row[RPT];
    u64 acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        row[r] = blockIdx.x * 128 * RPT + r * 128 + threadIdx.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0;
    }
    const float *xg = x + (size_t)g * n * 8;
#pragma unroll 1
    for (int k = 0; k < 27; ++k) {
        float xv[RPT][8];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            int nb = row[r] + (k - 13) * 3;
            nb = nb < 0 ? 0 : (nb >= n ? n - 1 : nb);
            gather8(xg + (size_t)nb * 8, xv[r]);
        }
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            u64 wq[4];
            if (SRC == 0) {
                const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(s_w + k * 64 + ci * 8);
                const ulonglong2 t0 = w2[0], t1 = w2[1];
                wq[0] = t0.x, wq[1] = t0.y, wq[2] = t1.x, wq[3] = t1.y;
            } else {
                const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(&cw[g][k * 64 + ci * 8]);
                const ulonglong2 t0 = w2[0], t1 = w2[1];
                wq[0] = t0.x, wq[1] = t0.y, wq[2] = t1.x, wq[3] = t1.y;
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const u64 xx = pack2(xv[r][ci], xv[r][ci]);
#pragma unroll
                for (int q = 0; q < 4; ++q) ffma2_acc(acc[r][q], xx, wq[q]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
        if (row[r] < n)
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<u64 *>(y + ((size_t)g * n + row[r]) * 8)[q] = acc[r][q];
}
template <int RPT, int SRC>
void run(const float *x, const float *w, float *y, int n, const char *name) {
    dim3 grid((n + 128 * RPT - 1) / (128 * RPT), 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<RPT, SRC><<<grid, 128>>>(x, w, y, n);
    cudaEventRecord(e0);
    const int it = 10;
    for (int i = 0; i < it; ++i) kern<RPT, SRC><<<grid, 128>>>(x, w, y, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= it;
    double macs = (double)n * 8 * 27 * 64;
    printf("%-28s %8.1f us  %6.2f TFLOP/s (%4.1f%% of 74.4)  err=%s\n", name, ms * 1e3, 2 * macs / ms / 1e9, 2 * macs / ms / 1e9 / 74.4 * 100, cudaGetErrorString(cudaGetLastError()));
}

// variant S: the tile's input rows are staged in shared memory first (stand-in for a TMA bulk copy), gathers are LDS.128 x2
template <int RPT, int PAD_KB>
__global__ void __launch_bounds__(128) kern_s(const float *__restrict__ x, const float *__restrict__ wg, float *__restrict__ y, int n) {
    extern __shared__ __align__(16) float s_all[];
    float *s_w = s_all;                 // 27*64
    float *s_x = s_all + 27 * 64;       // (128*RPT + 128) rows x 8
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < 27 * 64; i += 128) s_w[i] = wg[g * 27 * 64 + i];
    const float *xg = x + (size_t)g * n * 8;
    const int t0 = blockIdx.x * 128 * RPT - 64;
    const int rows_s = 128 * RPT + 128;
    for (int i = threadIdx.x; i < rows_s * 2; i += 128) {
        int rr = t0 + (i >> 1);
        rr = rr < 0 ? 0 : (rr >= n ? n - 1 : rr);
        reinterpret_cast<float4 *>(s_x)[i] = reinterpret_cast<const float4 *>(xg + (size_t)rr * 8)[i & 1];
    }
    __syncthreads();
    int row[RPT];
    u64 acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        row[r] = blockIdx.x * 128 * RPT + r * 128 + threadIdx.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0;
    }
#pragma unroll 1
    for (int k = 0; k < 27; ++k) {
        float xv[RPT][8];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int loc = row[r] + (k - 13) * 3 - t0;
            const float4 a = reinterpret_cast<const float4 *>(s_x)[loc * 2], b = reinterpret_cast<const float4 *>(s_x)[loc * 2 + 1];
            xv[r][0] = a.x, xv[r][1] = a.y, xv[r][2] = a.z, xv[r][3] = a.w, xv[r][4] = b.x, xv[r][5] = b.y, xv[r][6] = b.z, xv[r][7] = b.w;
        }
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(s_w + k * 64 + ci * 8);
            const ulonglong2 t0w = w2[0], t1w = w2[1];
            u64 wq[4] = {t0w.x, t0w.y, t1w.x, t1w.y};
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const u64 xx = pack2(xv[r][ci], xv[r][ci]);
#pragma unroll
                for (int q = 0; q < 4; ++q) ffma2_acc(acc[r][q], xx, wq[q]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
        if (row[r] < n)
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<u64 *>(y + ((size_t)g * n + row[r]) * 8)[q] = acc[r][q];
}
template <int RPT, int PAD_KB>
void run_s(const float *x, const float *w, float *y, int n, const char *name) {
    dim3 grid((n + 128 * RPT - 1) / (128 * RPT), 8);
    size_t smem = (size_t)PAD_KB * 1024;
    cudaFuncSetAttribute(kern_s<RPT, PAD_KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern_s<RPT, PAD_KB><<<grid, 128, smem>>>(x, w, y, n);
    cudaEventRecord(e0);
    const int it = 10;
    for (int i = 0; i < it; ++i) kern_s<RPT, PAD_KB><<<grid, 128, smem>>>(x, w, y, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= it;
    double macs = (double)n * 8 * 27 * 64;
    printf("%-28s %8.1f us  %6.2f TFLOP/s (%4.1f%% of 74.4)  err=%s\n", name, ms * 1e3, 2 * macs / ms / 1e9, 2 * macs / ms / 1e9 / 74.4 * 100, cudaGetErrorString(cudaGetLastError()));
}

// variant P: smem weights, neighbour rows of offset k+1 loaded into a second register buffer before the FMAs of offset k
template <int RPT>
__global__ void __launch_bounds__(128) kern_p(const float *__restrict__ x, const float *__restrict__ wg, float *__restrict__ y, int n) {
    __shared__ __align__(16) float s_w[27 * 64];
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < 27 * 64; i += 128) s_w[i] = wg[g * 27 * 64 + i];
    __syncthreads();
    int row[RPT];
    u64 acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        row[r] = blockIdx.x * 128 * RPT + r * 128 + threadIdx.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0;
    }
    const float *xg = x + (size_t)g * n * 8;
    auto load = [&](int k, float (&xv)[RPT][8]) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            int nb = row[r] + (k - 13) * 3;
            nb = nb < 0 ? 0 : (nb >= n ? n - 1 : nb);
            gather8(xg + (size_t)nb * 8, xv[r]);
        }
    };
    auto fma = [&](int k, const float (&xv)[RPT][8]) {
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(s_w + k * 64 + ci * 8);
            const ulonglong2 t0 = w2[0], t1 = w2[1];
            u64 wq[4] = {t0.x, t0.y, t1.x, t1.y};
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const u64 xx = pack2(xv[r][ci], xv[r][ci]);
#pragma unroll
                for (int q = 0; q < 4; ++q) ffma2_acc(acc[r][q], xx, wq[q]);
            }
        }
    };
    float xa[RPT][8], xb[RPT][8];
    load(0, xa);
#pragma unroll 1
    for (int k = 0; k < 26; k += 2) {
        load(k + 1, xb);
        fma(k, xa);
        load(k + 2, xa);
        fma(k + 1, xb);
    }
    fma(26, xa);
#pragma unroll
    for (int r = 0; r < RPT; ++r)
        if (row[r] < n)
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<u64 *>(y + ((size_t)g * n + row[r]) * 8)[q] = acc[r][q];
}
template <int RPT>
void run_p(const float *x, const float *w, float *y, int n, const char *name) {
    dim3 grid((n + 128 * RPT - 1) / (128 * RPT), 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern_p<RPT><<<grid, 128>>>(x, w, y, n);
    cudaEventRecord(e0);
    const int it = 10;
    for (int i = 0; i < it; ++i) kern_p<RPT><<<grid, 128>>>(x, w, y, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= it;
    double macs = (double)n * 8 * 27 * 64;
    printf("%-28s %8.1f us  %6.2f TFLOP/s (%4.1f%% of 74.4)  err=%s\n", name, ms * 1e3, 2 * macs / ms / 1e9, 2 * macs / ms / 1e9 / 74.4 * 100, cudaGetErrorString(cudaGetLastError()));
}

template <int RPT, int PAD_KB>
__global__ void __launch_bounds__(128) kern_sp(const float *__restrict__ x, const float *__restrict__ wg, float *__restrict__ y, int n) {
    extern __shared__ __align__(16) float s_all[];
    float *s_w = s_all;
    const int rows_s = 128 * RPT + 128;
    float4 *s_x0 = reinterpret_cast<float4 *>(s_all + 27 * 64);
    float4 *s_x1 = s_x0 + rows_s;
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < 27 * 64; i += 128) s_w[i] = wg[g * 27 * 64 + i];
    const float *xg = x + (size_t)g * n * 8;
    const int t0 = blockIdx.x * 128 * RPT - 64;
    for (int i = threadIdx.x; i < rows_s * 2; i += 128) {
        int rr = t0 + (i >> 1);
        rr = rr < 0 ? 0 : (rr >= n ? n - 1 : rr);
        const float4 v = reinterpret_cast<const float4 *>(xg + (size_t)rr * 8)[i & 1];
        if (i & 1) s_x1[i >> 1] = v; else s_x0[i >> 1] = v;
    }
    __syncthreads();
    int row[RPT];
    u64 acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        row[r] = blockIdx.x * 128 * RPT + r * 128 + threadIdx.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0;
    }
#pragma unroll 1
    for (int k = 0; k < 27; ++k) {
        float xv[RPT][8];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int loc = row[r] + (k - 13) * 3 - t0;
            const float4 a = s_x0[loc], b = s_x1[loc];
            xv[r][0] = a.x, xv[r][1] = a.y, xv[r][2] = a.z, xv[r][3] = a.w, xv[r][4] = b.x, xv[r][5] = b.y, xv[r][6] = b.z, xv[r][7] = b.w;
        }
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(s_w + k * 64 + ci * 8);
            const ulonglong2 t0w = w2[0], t1w = w2[1];
            u64 wq[4] = {t0w.x, t0w.y, t1w.x, t1w.y};
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const u64 xx = pack2(xv[r][ci], xv[r][ci]);
#pragma unroll
                for (int q = 0; q < 4; ++q) ffma2_acc(acc[r][q], xx, wq[q]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
        if (row[r] < n)
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<u64 *>(y + ((size_t)g * n + row[r]) * 8)[q] = acc[r][q];
}
template <int RPT, int PAD_KB>
void run_sp(const float *x, const float *w, float *y, int n, const char *name) {
    dim3 grid((n + 128 * RPT - 1) / (128 * RPT), 8);
    size_t smem = (size_t)PAD_KB * 1024;
    cudaFuncSetAttribute(kern_sp<RPT, PAD_KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern_sp<RPT, PAD_KB><<<grid, 128, smem>>>(x, w, y, n);
    cudaEventRecord(e0);
    const int it = 10;
    for (int i = 0; i < it; ++i) kern_sp<RPT, PAD_KB><<<grid, 128, smem>>>(x, w, y, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= it;
    double macs = (double)n * 8 * 27 * 64;
    printf("%-28s %8.1f us  %6.2f TFLOP/s (%4.1f%% of 74.4)  err=%s\n", name, ms * 1e3, 2 * macs / ms / 1e9, 2 * macs / ms / 1e9 / 74.4 * 100, cudaGetErrorString(cudaGetLastError()));
}

// variant CP: constant-bank weights + register double-buffered gathers
template <int RPT>
__global__ void __launch_bounds__(128) kern_cp(const float *__restrict__ x, float *__restrict__ y, int n) {
    const int g = blockIdx.y;
    int row[RPT];
    u64 acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        row[r] = blockIdx.x * 128 * RPT + r * 128 + threadIdx.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0;
    }
    const float *xg = x + (size_t)g * n * 8;
    auto load = [&](int k, float (&xv)[RPT][8]) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            int nb = row[r] + (k - 13) * 3;
            nb = nb < 0 ? 0 : (nb >= n ? n - 1 : nb);
            gather8(xg + (size_t)nb * 8, xv[r]);
        }
    };
    auto fma = [&](int k, const float (&xv)[RPT][8]) {
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(&cw[g][k * 64 + ci * 8]);
            const ulonglong2 t0 = w2[0], t1 = w2[1];
            u64 wq[4] = {t0.x, t0.y, t1.x, t1.y};
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const u64 xx = pack2(xv[r][ci], xv[r][ci]);
#pragma unroll
                for (int q = 0; q < 4; ++q) ffma2_acc(acc[r][q], xx, wq[q]);
            }
        }
    };
    float xa[RPT][8], xb[RPT][8];
    load(0, xa);
#pragma unroll 1
    for (int k = 0; k < 26; k += 2) {
        load(k + 1, xb);
        fma(k, xa);
        load(k + 2, xa);
        fma(k + 1, xb);
    }
    fma(26, xa);
#pragma unroll
    for (int r = 0; r < RPT; ++r)
        if (row[r] < n)
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<u64 *>(y + ((size_t)g * n + row[r]) * 8)[q] = acc[r][q];
}
template <int RPT>
void run_cp(const float *x, float *y, int n, const char *name) {
    dim3 grid((n + 128 * RPT - 1) / (128 * RPT), 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern_cp<RPT><<<grid, 128>>>(x, y, n);
    cudaEventRecord(e0);
    const int it = 10;
    for (int i = 0; i < it; ++i) kern_cp<RPT><<<grid, 128>>>(x, y, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= it;
    double macs = (double)n * 8 * 27 * 64;
    printf("%-28s %8.1f us  %6.2f TFLOP/s (%4.1f%% of 74.4)  err=%s\n", name, ms * 1e3, 2 * macs / ms / 1e9, 2 * macs / ms / 1e9 / 74.4 * 100, cudaGetErrorString(cudaGetLastError()));
}

// variant H: half of the weights (ci < SPLIT) from shared memory, the rest from the constant bank
template <int RPT, int SPLIT>
__global__ void __launch_bounds__(128) kern_h(const float *__restrict__ x, const float *__restrict__ wg, float *__restrict__ y, int n) {
    __shared__ __align__(16) float s_w[27 * 64];
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < 27 * 64; i += 128) s_w[i] = wg[g * 27 * 64 + i];
    __syncthreads();
    int row[RPT];
    u64 acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        row[r] = blockIdx.x * 128 * RPT + r * 128 + threadIdx.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0;
    }
    const float *xg = x + (size_t)g * n * 8;
#pragma unroll 1
    for (int k = 0; k < 27; ++k) {
        float xv[RPT][8];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            int nb = row[r] + (k - 13) * 3;
            nb = nb < 0 ? 0 : (nb >= n ? n - 1 : nb);
            gather8(xg + (size_t)nb * 8, xv[r]);
        }
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            u64 wq[4];
            if (ci < SPLIT) {
                const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(s_w + k * 64 + ci * 8);
                const ulonglong2 t0 = w2[0], t1 = w2[1];
                wq[0] = t0.x, wq[1] = t0.y, wq[2] = t1.x, wq[3] = t1.y;
            } else {
                const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(&cw[g][k * 64 + ci * 8]);
                const ulonglong2 t0 = w2[0], t1 = w2[1];
                wq[0] = t0.x, wq[1] = t0.y, wq[2] = t1.x, wq[3] = t1.y;
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const u64 xx = pack2(xv[r][ci], xv[r][ci]);
#pragma unroll
                for (int q = 0; q < 4; ++q) ffma2_acc(acc[r][q], xx, wq[q]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
        if (row[r] < n)
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<u64 *>(y + ((size_t)g * n + row[r]) * 8)[q] = acc[r][q];
}
template <int RPT, int SPLIT>
void run_h(const float *x, const float *w, float *y, int n, const char *name) {
    dim3 grid((n + 128 * RPT - 1) / (128 * RPT), 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern_h<RPT, SPLIT><<<grid, 128>>>(x, w, y, n);
    cudaEventRecord(e0);
    const int it = 10;
    for (int i = 0; i < it; ++i) kern_h<RPT, SPLIT><<<grid, 128>>>(x, w, y, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= it;
    double macs = (double)n * 8 * 27 * 64;
    printf("%-28s %8.1f us  %6.2f TFLOP/s (%4.1f%% of 74.4)  err=%s\n", name, ms * 1e3, 2 * macs / ms / 1e9, 2 * macs / ms / 1e9 / 74.4 * 100, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int n = 277000;
    float *x, *y, *w;
    cudaMalloc(&x, (size_t)n * 8 * 8 * 4), cudaMalloc(&y, (size_t)n * 8 * 8 * 4), cudaMalloc(&w, 8 * 27 * 64 * 4);
    std::vector<float> h((size_t)n * 64, 0.5f), hw(8 * 27 * 64, 0.01f);
    cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(cw, w, hw.size() * 4, 0, cudaMemcpyDeviceToDevice);
    run<1, 0>(x, w, y, n, "smem  RPT=1");
    run<2, 0>(x, w, y, n, "smem  RPT=2");
    run<4, 0>(x, w, y, n, "smem  RPT=4");
    run<8, 0>(x, w, y, n, "smem  RPT=8");
    run<8, 1>(x, w, y, n, "const RPT=8");
    run_p<2>(x, w, y, n, "prefetch RPT=2");
    run_p<4>(x, w, y, n, "prefetch RPT=4");
    run_sp<4, 32>(x, w, y, n, "planar staged RPT=4 32K(7blk)");
    run_sp<4, 56>(x, w, y, n, "planar staged RPT=4 56K(4blk)");
    run_sp<4, 75>(x, w, y, n, "planar staged RPT=4 75K(3blk)");
    run_sp<8, 75>(x, w, y, n, "planar staged RPT=8 75K(3blk)");
    run_s<4, 32>(x, w, y, n, "staged RPT=4 smem32K(7blk)");
    run_s<4, 56>(x, w, y, n, "staged RPT=4 smem56K(4blk)");
    run_s<4, 75>(x, w, y, n, "staged RPT=4 smem75K(3blk)");
    run_s<4, 110>(x, w, y, n, "staged RPT=4 smem110K(2blk)");
    run_s<2, 32>(x, w, y, n, "staged RPT=2 smem32K");
    run_cp<2>(x, y, n, "const+prefetch RPT=2");
    run_cp<4>(x, y, n, "const+prefetch RPT=4");
    run_h<4, 2>(x, w, y, n, "hybrid 2 smem + 6 const RPT=4");
    run_h<4, 4>(x, w, y, n, "hybrid 4 smem + 4 const RPT=4");
    run<1, 1>(x, w, y, n, "const RPT=1");
    run<2, 1>(x, w, y, n, "const RPT=2");
    run<4, 1>(x, w, y, n, "const RPT=4");
    return 0;
}
