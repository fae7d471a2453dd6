// cp.async.bulk.tensor.2d ... tile::gather4 (sm_100a): four 32-byte rows of an [R][8] fp32 tensor per instruction.
// Question (VERDICT r1, item 2 ii): can the TMA engine gather the neighbour rows of conv27 / bwd_w into shared memory
// faster than the LSU does with LDG.128 pairs?  Rows follow the kernel map's shape: runs of three consecutive rows
// (a column's dz = -1, 0, +1) at a pseudo-random run start.
//   * latency of one gather4 (one thread, waited for);
//   * issue cost: cycles per gather4 when one thread / one warp issues a tile's worth back to back;
//   * aggregate rows/s with every SM gathering, against an LDG.128 gather of the same rows into shared memory.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);          \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void gather4(void *dst, const CUtensorMap *tm, int col, int r0, int r1, int r2, int r3, uint64_t *b) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(s32(dst)),
        "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(s32(b))
        : "memory");
}

// row index of neighbour slot j of tile-local row i: runs of three consecutive rows
__device__ __forceinline__ int nbr_row(int base, int i, int j, int R) {
    const unsigned h = (unsigned)(base + i) * 2654435761u + (unsigned)(j / 3) * 40503u;
    const int start = (int)((base + i + (int)(h % 4096u) - 2048 + R) % (R - 3));
    return start + j % 3;
}

constexpr int ROWB = 32;   // bytes per row

// correctness + latency: one thread, one gather4 at a time
__global__ void lat_kernel(const __grid_constant__ CUtensorMap tm, const float *src, int R, int reps, long long *cyc, int *bad) {
    __shared__ __align__(128) float buf[32];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence();
        long long tot = 0;
        int nbad = 0;
        for (int r = 0; r < reps; ++r) {
            int rows[4];
            for (int q = 0; q < 4; ++q) rows[q] = nbr_row(blockIdx.x * 977 + r * 131, q, q * 3, R);
            const long long t0 = clock64();
            mbar_expect(&bar, 4 * ROWB);
            gather4(buf, &tm, 0, rows[0], rows[1], rows[2], rows[3], &bar);
            mbar_wait(&bar, r & 1);
            tot += clock64() - t0;
            for (int q = 0; q < 4; ++q)
                for (int c = 0; c < 8; ++c) nbad += buf[q * 8 + c] != src[(size_t)rows[q] * 8 + c];
        }
        cyc[blockIdx.x] = tot;
        bad[blockIdx.x] = nbad;
    }
}

// One tile = TR rows x NJ neighbour slots.  `issuers` threads issue the gather4s of a tile (slot-major, 4 consecutive
// tile rows of one slot per instruction) into a 2-stage ring; the whole block then reads the tile once from shared memory.
template <int TR, int NJ>
__global__ void tma_gather_kernel(const __grid_constant__ CUtensorMap tm, int R, int tiles, int issuers, long long *issue_cyc, float *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[2];
    constexpr int TILEB = TR * NJ * ROWB, NG = TR * NJ / 4;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1), mbar_init(&bar[1], 1);
        mbar_fence();
    }
    __syncthreads();
    long long icyc = 0;
    auto issue = [&](int t) {
        const int st = t & 1;
        if (threadIdx.x == 0) mbar_expect(&bar[st], TILEB);
        __syncwarp();
        const long long t0 = clock64();
        const int base = (blockIdx.x * tiles + t) * TR;
        for (int gi = threadIdx.x; gi < NG && (int)threadIdx.x < issuers; gi += issuers) {
            const int j = gi / (TR / 4), i = (gi % (TR / 4)) * 4;
            gather4(smem + st * TILEB + gi * 4 * ROWB, &tm, 0, nbr_row(base, i, j, R), nbr_row(base, i + 1, j, R), nbr_row(base, i + 2, j, R),
                    nbr_row(base, i + 3, j, R), &bar[st]);
        }
        icyc += clock64() - t0;
    };
    if (threadIdx.x < 32) issue(0);
    float acc = 0.f;
    for (int t = 0; t < tiles; ++t) {
        const int st = t & 1;
        if (t + 1 < tiles && threadIdx.x < 32) issue(t + 1);
        mbar_wait(&bar[st], (t >> 1) & 1);
        const float4 *p = reinterpret_cast<const float4 *>(smem + st * TILEB);
        for (int i = threadIdx.x; i < TILEB / 16; i += blockDim.x) acc += p[i].x;
        __syncthreads();
    }
    if (threadIdx.x == 0) issue_cyc[blockIdx.x] = icyc;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// the same rows through the LSU: every thread gathers rows with two LDG.128, stores them to shared memory, the block reads
template <int TR, int NJ>
__global__ void ldg_gather_kernel(const float *src, int R, int tiles, float *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TILEB = TR * NJ * ROWB;
    float acc = 0.f;
    for (int t = 0; t < tiles; ++t) {
        const int base = (blockIdx.x * tiles + t) * TR;
        for (int e = threadIdx.x; e < TR * NJ; e += blockDim.x) {
            const int j = e / TR, i = e % TR;
            const float4 *q = reinterpret_cast<const float4 *>(src + (size_t)nbr_row(base, i, j, R) * 8);
            const float4 a = __ldg(q), b = __ldg(q + 1);
            float4 *d = reinterpret_cast<float4 *>(smem + (size_t)e * ROWB);
            d[0] = a, d[1] = b;
        }
        __syncthreads();
        const float4 *p = reinterpret_cast<const float4 *>(smem);
        for (int i = threadIdx.x; i < TILEB / 16; i += blockDim.x) acc += p[i].x;
        __syncthreads();
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int R = 277 * 1024;   // rows of the finest scale of a loot frame
    std::vector<float> h((size_t)R * 8);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000003);
    float *d;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn) return printf("no cuTensorMapEncodeTiled\n"), 1;
    CUtensorMap tm;
    const cuuint64_t dims[2] = {8, (cuuint64_t)R}, strides[1] = {32};
    const cuuint32_t box[2] = {8, 1}, es[2] = {1, 1};
    const CUresult cr = ((EncodeTiled)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr), 1;
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("# SM clock (attr) %d MHz; tensor [%d][8] fp32, box {8,1}\n", clk / 1000, R);

    long long *cyc;
    int *bad;
    float *sink;
    CK(cudaMalloc(&cyc, 1024 * 8));
    CK(cudaMalloc(&bad, 1024 * 4));
    CK(cudaMalloc(&sink, 1024 * 512 * 4));
    {
        const int reps = 64;
        lat_kernel<<<4, 32>>>(tm, d, R, reps, cyc, bad);
        CK(cudaDeviceSynchronize());
        long long hc[4];
        int hb[4];
        CK(cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hb, bad, sizeof(hb), cudaMemcpyDeviceToHost));
        printf("# one gather4 (4 x 32 B), issued and waited for by one thread: %lld cycles; mismatching words: %d\n", hc[0] / reps, hb[0] + hb[1] + hb[2] + hb[3]);
    }
    constexpr int TR = 64, NJ = 27;   // a 64-row tile: 27 x 64 x 32 B = 54 KB per stage
    const int tiles = 32, smem = 2 * TR * NJ * ROWB;
    CK(cudaFuncSetAttribute(tma_gather_kernel<TR, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(ldg_gather_kernel<TR, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem / 2));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("# every block gathers %d tiles of %d rows x %d slots (runs of 3 consecutive rows), 256 threads\n", tiles, TR, NJ);
    printf("%-28s %7s %9s %12s %10s %22s\n", "variant", "blocks", "ms", "Mrows/s", "GB/s", "issue cycles/gather4");
    for (int blocks : {148, 296}) {
        for (int issuers : {1, 32}) {
            tma_gather_kernel<TR, NJ><<<blocks, 256, smem>>>(tm, R, tiles, issuers, cyc, sink);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            tma_gather_kernel<TR, NJ><<<blocks, 256, smem>>>(tm, R, tiles, issuers, cyc, sink);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            long long hc;
            CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
            const double rows = (double)blocks * tiles * TR * NJ;
            char name[64];
            snprintf(name, sizeof(name), "gather4, %d issuing thread%s", issuers, issuers > 1 ? "s" : "");
            const double per = (double)hc / ((double)tiles * (TR * NJ / 4) / issuers);
            printf("%-28s %7d %9.3f %12.1f %10.1f %22.1f\n", name, blocks, ms, rows / ms * 1e-3, rows * ROWB / ms * 1e-6, per);
        }
        ldg_gather_kernel<TR, NJ><<<blocks, 256, smem / 2>>>(d, R, tiles, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        ldg_gather_kernel<TR, NJ><<<blocks, 256, smem / 2>>>(d, R, tiles, sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double rows = (double)blocks * tiles * TR * NJ;
        printf("%-28s %7d %9.3f %12.1f %10.1f %22s\n", "LDG.128 x2 -> STS, 256 thr", blocks, ms, rows / ms * 1e-3, rows * ROWB / ms * 1e-6, "-");
    }
    return 0;
}
