// cp.async.bulk (1-D TMA, UBLKCP) global -> shared: latency of one transfer as a function of size and piece count,
// from L2 and from DRAM, and the aggregate bandwidth when every SM streams tiles through a double buffer.
// Answers: how far ahead must a producer issue the staged neighbour ranges of conv27 / bwd_w (net_kernels.cuh)?
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void *d, const void *s, uint32_t n, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}

// one block, one thread: `reps` transfers of `bytes` split into `pieces` copies, each waited for before the next
__global__ void lat_kernel(const char *src, size_t stride, int bytes, int pieces, int reps, long long *cyc, float *sink) {
    extern __shared__ __align__(128) char buf[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence();
        long long tot = 0;
        float acc = 0.f;
        const int pb = bytes / pieces;
        for (int r = 0; r < reps; ++r) {
            const char *s = src + (size_t)r * stride;
            const long long t0 = clock64();
            mbar_expect(&bar, (uint32_t)(pb * pieces));
            for (int p = 0; p < pieces; ++p) bulk(buf + p * pb, s + (size_t)p * pb, pb, &bar);
            mbar_wait(&bar, r & 1);
            tot += clock64() - t0;
            acc += reinterpret_cast<float *>(buf)[0];
        }
        cyc[blockIdx.x] = tot;
        sink[blockIdx.x] = acc;
    }
}

// every block streams `tiles` tiles of `bytes` through a 2-stage ring; thread 0 issues, all threads read one word
__global__ void bw_kernel(const char *src, int bytes, int tiles, float *sink) {
    extern __shared__ __align__(128) char buf[];
    __shared__ __align__(8) uint64_t bar[2];
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1), mbar_init(&bar[1], 1);
        mbar_fence();
    }
    __syncthreads();
    const char *s = src + (size_t)blockIdx.x * tiles * bytes;
    if (threadIdx.x == 0) {
        mbar_expect(&bar[0], bytes);
        bulk(buf, s, bytes, &bar[0]);
    }
    float acc = 0.f;
    for (int t = 0; t < tiles; ++t) {
        const int st = t & 1;
        if (threadIdx.x == 0 && t + 1 < tiles) {
            mbar_expect(&bar[st ^ 1], bytes);
            bulk(buf + (st ^ 1) * bytes, s + (size_t)(t + 1) * bytes, bytes, &bar[st ^ 1]);
        }
        mbar_wait(&bar[st], (t >> 1) & 1);
        acc += reinterpret_cast<float *>(buf + st * bytes)[threadIdx.x];
        __syncthreads();
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    const size_t N = 1ull << 30;
    char *src, *fl;
    long long *cyc;
    float *sink;
    cudaMalloc(&src, N), cudaMemset(src, 1, N);
    cudaMalloc(&fl, 512 << 20);
    cudaMalloc(&cyc, 1024 * 8), cudaMalloc(&sink, 1 << 22);
    cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10);
    cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("# SM clock (attr) %.0f MHz\n", clk / 1e3);
    printf("# latency of one transfer (cycles), one thread of one block, waits for each before the next\n");
    printf("%8s %7s %12s %12s\n", "bytes", "pieces", "L2-resident", "DRAM");
    const int reps = 64;
    for (int bytes : {1024, 4096, 8192, 16384, 32768, 65536}) {
        for (int pieces : {1, 3, 13}) {
            if (bytes / pieces % 16) {
                // round the piece down to 16 bytes
            }
            const int pb = bytes / pieces / 16 * 16, b = pb * pieces;
            long long h = 0;
            double res[2];
            for (int dram = 0; dram < 2; ++dram) {
                // L2-resident: the same `reps` tiles were touched by a previous identical run; DRAM: flush in between
                lat_kernel<<<1, 32, b>>>(src, 1 << 20, b, pieces, reps, cyc, sink);
                cudaDeviceSynchronize();
                if (dram) cudaMemset(fl, dram, 512 << 20);
                lat_kernel<<<1, 32, b>>>(src, 1 << 20, b, pieces, reps, cyc, sink);
                cudaDeviceSynchronize();
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
                res[dram] = (double)h / reps;
            }
            printf("%8d %7d %12.0f %12.0f\n", b, pieces, res[0], res[1]);
        }
    }
    printf("# aggregate bandwidth: grid blocks x 128 threads, 2-stage ring of `bytes` tiles, 64 tiles per block\n");
    printf("%8s %8s %10s %10s\n", "bytes", "blocks", "ms", "GB/s");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int bytes : {8192, 32768}) {
        for (int bps : {1, 2, 3}) {
            const int blocks = 148 * bps, tiles = 64;
            if ((size_t)blocks * tiles * bytes > N) continue;
            for (int w = 0; w < 2; ++w) {
                cudaMemset(fl, w, 512 << 20);
                cudaEventRecord(e0);
                bw_kernel<<<blocks, 128, 2 * bytes>>>(src, bytes, tiles, sink);
                cudaEventRecord(e1);
                cudaDeviceSynchronize();
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%8d %8d %10.3f %10.1f   (%s)\n", bytes, blocks, ms, (double)blocks * tiles * bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
