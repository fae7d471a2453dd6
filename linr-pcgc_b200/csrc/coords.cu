// Coordinate stage: lexicographic sort/unique, octree down/up, open-addressing hash, 27-neighbour kernel map.
// Integer work, HBM/L2 bound; bit-exact against oracle/linr_oracle.py (which is pinned to the reference's
// own Python: models/sort_functions.py, models/quantize_functions.py, models/module_utils.py:86-318).
#include <cub/cub.cuh>

#include "common.cuh"
#include "prof.cuh"

using linr::K_COORD;
using linr::ProfScope;

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ uint64_t ckey(int x, int y, int z, int bits) {
    return ((uint64_t)(uint32_t)x << (2 * bits)) | ((uint64_t)(uint32_t)y << bits) | (uint64_t)(uint32_t)z;
}
__device__ __forceinline__ void cunpack(uint64_t k, int bits, int &x, int &y, int &z) {
    const uint64_t m = (1ull << bits) - 1;
    x = (int)((k >> (2 * bits)) & m);
    y = (int)((k >> bits) & m);
    z = (int)(k & m);
}

__global__ void pack_kernel(const int32_t *__restrict__ xyz, int64_t n, int bits, int shift, uint64_t *__restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ckey(xyz[3 * i] >> shift, xyz[3 * i + 1] >> shift, xyz[3 * i + 2] >> shift, bits);
}

// n may live on the device (result of a unique); n_host_bound is the launch bound.
__global__ void unpack_kernel(const uint64_t *__restrict__ keys, const int64_t *__restrict__ d_n, int64_t n_bound, int bits,
                              int32_t *__restrict__ xyz) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t n = d_n ? *d_n : n_bound;
    if (i >= n) return;
    int x, y, z;
    cunpack(keys[i], bits, x, y, z);
    xyz[3 * i] = x;
    xyz[3 * i + 1] = y;
    xyz[3 * i + 2] = z;
}

__global__ void min3_kernel(const int32_t *__restrict__ xyz, int64_t n, int32_t *__restrict__ out) {
    // single block, deterministic (integer min is order independent anyway)
    int mx = INT32_MAX, my = INT32_MAX, mz = INT32_MAX;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        mx = min(mx, xyz[3 * i]);
        my = min(my, xyz[3 * i + 1]);
        mz = min(mz, xyz[3 * i + 2]);
    }
    for (int o = 16; o; o >>= 1) {
        mx = min(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        my = min(my, __shfl_xor_sync(0xffffffffu, my, o));
        mz = min(mz, __shfl_xor_sync(0xffffffffu, mz, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out + 0, mx);
        atomicMin(out + 1, my);
        atomicMin(out + 2, mz);
    }
}
__global__ void fill3_kernel(int32_t *p, int v) {
    if (threadIdx.x < 3) p[threadIdx.x] = v;
}
__global__ void sub3_kernel(const int32_t *__restrict__ in, int64_t n, const int32_t *__restrict__ mn, int32_t *__restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= 3 * n) return;
    out[i] = in[i] - mn[i % 3];
}

// occ[parent] |= 1 << octant for every child; parent row by binary search in the sorted unique parent keys.
__global__ void occ_kernel(const int32_t *__restrict__ child, int64_t nc, int bits, const uint64_t *__restrict__ pkeys,
                           const int64_t *__restrict__ d_np, uint32_t *__restrict__ occ_words) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const int x = child[3 * i], y = child[3 * i + 1], z = child[3 * i + 2];
    const uint64_t pk = ckey(x >> 1, y >> 1, z >> 1, bits);
    int64_t lo = 0, hi = *d_np;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (pkeys[mid] < pk) lo = mid + 1;
        else hi = mid;
    }
    const int oct = ((x & 1) << 2) | ((y & 1) << 1) | (z & 1);  // models/module_utils.py:93
    atomicOr(occ_words + (lo >> 2), (1u << oct) << (8 * (lo & 3)));
}

__global__ void popc_kernel(const uint8_t *__restrict__ occ, int64_t n, int64_t *__restrict__ cnt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = __popc((unsigned)occ[i]);
    if (i == n) cnt[i] = 0;
}

__global__ void expand_kernel(const int32_t *__restrict__ parent, const uint8_t *__restrict__ occ,
                              const int64_t *__restrict__ off, int64_t n, int bits, uint64_t *__restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = parent[3 * i] * 2, y = parent[3 * i + 1] * 2, z = parent[3 * i + 2] * 2;
    unsigned o = occ[i];
    int64_t w = off[i];
    for (int j = 0; j < 8; ++j)
        if (o >> j & 1) keys[w++] = ckey(x + (j >> 2 & 1), y + (j >> 1 & 1), z + (j & 1), bits);
}

// ---- hash -------------------------------------------------------------------------------------
constexpr uint64_t HEMPTY = ~0ull;
struct HashSlot {
    unsigned long long key;
    int32_t row;
    int32_t pad;
};

__device__ __forceinline__ uint64_t hkey(int s, int x, int y, int z) {
    return ((uint64_t)s << 60) | ((uint64_t)(uint32_t)x << 40) | ((uint64_t)(uint32_t)y << 20) | (uint64_t)(uint32_t)z;
}
__device__ __forceinline__ uint64_t hmix(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

__global__ void hash_clear_kernel(HashSlot *t, int64_t cap) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < cap) {
        t[i].key = HEMPTY;
        t[i].row = -1;
        t[i].pad = 0;
    }
}

__global__ void hash_insert_kernel(const int32_t *__restrict__ xyz, const uint8_t *__restrict__ scale, int64_t n,
                                   HashSlot *__restrict__ t, int64_t cap) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = hkey(scale ? scale[i] : 0, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    uint64_t s = hmix(k) & (cap - 1);
    for (;;) {
        unsigned long long old = atomicCAS(&t[s].key, (unsigned long long)HEMPTY, (unsigned long long)k);
        if (old == HEMPTY || old == k) {
            t[s].row = (int32_t)i;  // rows are unique per (scale, coord): a single writer
            return;
        }
        s = (s + 1) & (cap - 1);
    }
}

__device__ __forceinline__ int hash_find(const HashSlot *__restrict__ t, int64_t cap, int sc, int x, int y, int z) {
    if ((x | y | z) < 0 || x >= (1 << 20) || y >= (1 << 20) || z >= (1 << 20)) return -1;
    const uint64_t k = hkey(sc, x, y, z);
    uint64_t s = hmix(k) & (cap - 1);
    for (;;) {
        const uint64_t cur = t[s].key;
        if (cur == k) return t[s].row;
        if (cur == HEMPTY) return -1;
        s = (s + 1) & (cap - 1);
    }
}

__global__ void hash_lookup_kernel(const int32_t *__restrict__ q, const uint8_t *__restrict__ qs, int64_t nq,
                                   const HashSlot *__restrict__ t, int64_t cap, int32_t *__restrict__ rows) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nq) return;
    rows[i] = hash_find(t, cap, qs ? qs[i] : 0, q[3 * i], q[3 * i + 1], q[3 * i + 2]);
}

// One thread per row: 27 probes -> dense table (optional), compact anchors + mask, 7 face bits.
__global__ void nbr_kernel(const int32_t *__restrict__ xyz, const uint8_t *__restrict__ scale, int64_t n,
                           const HashSlot *__restrict__ t, int64_t cap, int32_t *__restrict__ nbr27,
                           int32_t *__restrict__ anchor, int64_t ld, uint32_t *__restrict__ mask, uint8_t *__restrict__ nbr7) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const int sc = scale ? scale[i] : 0;
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        const int dx = c % 3 - 1, dy = c / 3 - 1;
        int a = 0;
        bool have = false;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int k = c + 9 * j;
            const int r = (k == 13) ? (int)i : hash_find(t, cap, sc, x + dx, y + dy, z + j - 1);
            if (nbr27) nbr27[i * 27 + k] = r;
            if (r >= 0) {
                m |= 1u << (3 * c + j);
                if (!have) {
                    a = r;
                    have = true;
                }
            }
        }
        if (anchor) anchor[c * ld + i] = a;
    }
    if (mask) mask[i] = m;
    if (nbr7) {
        // offsets_ini order (main.py:24): self, -x, +x, -y, +y, -z, +z  ->  k = 13, 12, 14, 10, 16, 4, 22
        // mask bit of k: 3*(k%9) + k/9
        auto bit = [&](int k) { return (m >> (3 * (k % 9) + k / 9)) & 1u; };
        nbr7[i] = (uint8_t)(bit(13) | bit(12) << 1 | bit(14) << 2 | bit(10) << 3 | bit(16) << 4 | bit(4) << 5 | bit(22) << 6);
    }
}


// Per 128-row tile and per dx in {-1,0,+1}: the row range [lo, hi) that holds every neighbour of the tile's rows
// with that dx.  Rows are x-major sorted, so o -> row(C[o] + delta) is monotone for a fixed delta: the neighbours of
// a tile of consecutive rows sit in three nearly contiguous row ranges, which the conv / weight-gradient kernels
// stage into shared memory with bulk (TMA) copies instead of gathering them line by line through L1.
__global__ void __launch_bounds__(128) tile_range_kernel(const int32_t *__restrict__ anchor, int64_t ld,
                                                         const uint32_t *__restrict__ mask, int64_t n, int32_t *__restrict__ rng) {
    __shared__ int s_lo[4][3], s_hi[4][3];
    const int64_t row = blockIdx.x * 128ll + threadIdx.x;
    int lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {0, 0, 0};
    if (row < n) {
        const uint32_t m = mask[row];
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const uint32_t m3 = (m >> (3 * c)) & 7u;
            if (m3) {
                const int a = anchor[c * ld + row];
                lo[c % 3] = min(lo[c % 3], a);
                hi[c % 3] = max(hi[c % 3], a + __popc(m3));
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) s_lo[threadIdx.x >> 5][d] = lo[d], s_hi[threadIdx.x >> 5][d] = hi[d];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        int l = s_lo[0][d], h = s_hi[0][d];
        for (int w = 1; w < 4; ++w) l = min(l, s_lo[w][d]), h = max(h, s_hi[w][d]);
        if (h <= l) l = 0, h = 0;
        rng[blockIdx.x * 6 + 2 * d] = l;
        rng[blockIdx.x * 6 + 2 * d + 1] = h;
    }
}

// Pair lists: for every 256-row tile and every kernel offset k the (output row, neighbour row) pairs that exist:
// entry = (row - tile_base) << 24 | neighbour_row; cnt[t * 32 + k] entries.  Storage: 27 * 256 entries per tile; the
// lists of the offsets of half h (LINR_BW3_SLOTS, common.cuh) lie back to back from entry h * 14 * 256 of the tile's
// storage, in table order, every list starting on a multiple of 4 entries -- what one block of the weight-gradient
// kernel needs of a tile is ONE contiguous run.  That kernel walks a list 32 pairs at a time,
// one pair per lane, and reads both rows (32 bytes each) from shared memory with two 16-byte loads per row.  A quarter
// warp of such loads is conflict-free when each aligned group of FOUR lanes holds rows that differ modulo 4 (lanes
// 0..3 of a quarter read one half of their rows, lanes 4..7 the other, see RawRow in net_kernels.cuh).  In row order
// the rows of a list have gaps and 1.63 x the ideal number of wavefronts; so the ORDER inside a list is chosen here,
// once per frame: entries are classed by (row mod 4, neighbour mod 4); every cyclic diagonal
// {(a, a + d mod 4), a = 0..3} yields min-count groups of four entries with distinct rows AND distinct neighbours
// (neighbour - row is constant along a column run, so most entries sit on one diagonal); what is left over follows in
// row order.  1.2 x ideal for both operands.  One warp per (tile, column c) builds the three lists of its column
// (k = c + 9 j) with ballots only, so a list is a deterministic function of the kernel map.
__constant__ int8_t c_pair_slot[2][BW3_NW] = LINR_BW3_SLOTS;
struct PairPos {
    int8_t v[27];   // half << 4 | index inside the half, of offset k
};
constexpr PairPos make_pair_pos() {
    constexpr int8_t t[2][BW3_NW] = LINR_BW3_SLOTS;
    PairPos p{};
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < BW3_NW; ++i)
            if (t[h][i] < 27) p.v[t[h][i]] = (int8_t)(h << 4 | i);
    return p;
}
__constant__ PairPos c_pair_pos_tab = make_pair_pos();
__global__ void __launch_bounds__(288) pair_list_kernel(const int32_t *__restrict__ anchor, int64_t ld, const uint32_t *__restrict__ mask,
                                                        int64_t n, int32_t *__restrict__ cnt, uint32_t *__restrict__ list) {
    __shared__ uint32_t s_asc[9][3][256];   // the lists in row order
    __shared__ uint8_t s_rank[9][256];      // rank of an entry inside its class
    __shared__ int s_cnt[27];
    const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;   // 9 warps: one per (dx,dy) column
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t tile = blockIdx.x, base = tile * 256;
    int count[3] = {0, 0, 0};
    for (int st = 0; st < 8; ++st) {
        const int rl = st * 32 + lane;
        const int64_t row = base + rl;
        uint32_t m3 = 0;
        int an = 0;
        if (row < n) {
            m3 = (mask[row] >> (3 * c)) & 7u;
            if (m3) an = anchor[c * ld + row];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const bool has = (m3 >> j) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, has);
            if (has) {
                const int nb = an + __popc(m3 & ((1u << j) - 1u));
                s_asc[c][j][count[j] + __popc(bal & lt)] = ((uint32_t)rl << 24) | (uint32_t)nb;
            }
            count[j] += __popc(bal);
        }
    }
    if (lane < 3) s_cnt[c + 9 * lane] = lane == 0 ? count[0] : (lane == 1 ? count[1] : count[2]);
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
        const int nj = count[j];
        const uint32_t *asc = s_asc[c][j];
        // where list k starts: after the lists that precede it in its half (each rounded up to 4 entries)
        const int k = c + 9 * j, half = c_pair_pos_tab.v[k] >> 4, idx = c_pair_pos_tab.v[k] & 15;
        int before = (lane < idx) ? ((s_cnt[c_pair_slot[half][lane]] + 3) & ~3) : 0;
#pragma unroll
        for (int o = 8; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        before = __shfl_sync(0xffffffffu, before, 0);
        uint32_t *out = list + tile * PAIR_TILE_ENTRIES + half * PAIR_HALF_ENTRIES + before;
        // pass 1: class counts (warp-uniform registers) and the rank of every entry inside its class
        int ncls[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) ncls[q] = 0;
        for (int i0 = 0; i0 < nj; i0 += 32) {
            const bool on = i0 + lane < nj;
            const uint32_t e = on ? asc[i0 + lane] : 0u;
            const int cls = on ? (int)(((e >> 24) & 3u) * 4u + (e & 3u)) : -1;
            int rk = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const uint32_t bal = __ballot_sync(0xffffffffu, cls == q);
                if (cls == q) rk = ncls[q] + __popc(bal & lt);
                ncls[q] += __popc(bal);
            }
            if (on) s_rank[c][i0 + lane] = (uint8_t)rk;   // < 256: a class with 256 entries would need 1024 rows
        }
        // groups of four per diagonal d: g[d] = min over a of count(a, a + d); diagonals laid out one after the other
        int g[4], gb[5];
        gb[0] = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            g[d] = 256;
#pragma unroll
            for (int a = 0; a < 4; ++a) g[d] = min(g[d], ncls[a * 4 + ((a + d) & 3)]);
            gb[d + 1] = gb[d] + g[d];
        }
        __syncwarp();
        // pass 2: grouped entries to their slot, the rest behind them in row order
        int tail = 4 * gb[4];
        for (int i0 = 0; i0 < nj; i0 += 32) {
            const bool on = i0 + lane < nj;
            const uint32_t e = on ? asc[i0 + lane] : 0u;
            const int a = (int)((e >> 24) & 3u), d = (int)((e - (e >> 24)) & 3u);
            const int rk = on ? (int)s_rank[c][i0 + lane] : 0;
            int gd = g[0], gbd = gb[0];
#pragma unroll
            for (int q = 1; q < 4; ++q)
                if (d == q) gd = g[q], gbd = gb[q];
            const bool grouped = on && rk < gd;
            const uint32_t rest = __ballot_sync(0xffffffffu, on && !grouped);
            if (grouped) out[4 * (gbd + rk) + a] = e;
            else if (on) out[tail + __popc(rest & lt)] = e;
            tail += __popc(rest);
        }
        __syncwarp();
    }
    if (lane < 3) cnt[tile * 32 + c + 9 * lane] = lane == 0 ? count[0] : (lane == 1 ? count[1] : count[2]);
    if (c == 0 && lane >= 27) cnt[tile * 32 + lane] = 0;
}

struct SortWs {
    uint64_t *a, *b;
    int64_t *cnt;
    void *cub;
    size_t cub_bytes;
};

size_t cub_bytes_for(int64_t n) {
    size_t s1 = 0, s2 = 0, s3 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, s1, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)n, 0, 64);
    cub::DeviceSelect::Unique(nullptr, s2, (uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t *)nullptr, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, s3, (int64_t *)nullptr, (int64_t *)nullptr, (int)n + 1);
    size_t m = s1 > s2 ? s1 : s2;
    return (m > s3 ? m : s3) + 256;
}

bool carve(void *ws, size_t bytes, int64_t n, SortWs &w) {
    WsCursor c(ws, bytes);
    w.a = c.take<uint64_t>(n + 1);
    w.b = c.take<uint64_t>(n + 1);
    w.cnt = c.take<int64_t>(4);
    w.cub_bytes = cub_bytes_for(n);
    w.cub = c.take<char>(w.cub_bytes);
    return c.ok;
}

inline int grid_for(int64_t n) { return (int)ceil_div64(n > 0 ? n : 1, TPB); }

}  // namespace

extern "C" {

size_t linr_coord_ws_bytes(int64_t n) {
    if (n < 1) n = 1;
    return 2 * align_up((n + 1) * 8, 256) + 256 + cub_bytes_for(n) + 1024;
}

int linr_coord_min_sub(const int32_t *d_in, int64_t n, int32_t *d_out, int32_t *d_min, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    LINR_REQUIRE(n > 0, "linr_coord_min_sub: empty input");
    {
        ProfScope prof(K_COORD, 1, s);
        fill3_kernel<<<1, 32, 0, s>>>(d_min, INT32_MAX);
    }
    int g = (int)(ceil_div64(n, TPB) < 1024 ? ceil_div64(n, TPB) : 1024);
    {
        ProfScope prof(K_COORD, 1, s);
        min3_kernel<<<g, TPB, 0, s>>>(d_in, n, d_min);
    }
    {
        ProfScope prof(K_COORD, 1, s);
        sub3_kernel<<<grid_for(3 * n), TPB, 0, s>>>(d_in, n, d_min, d_out);
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

static int sort_impl(const int32_t *d_in, int64_t n, int bits, int shift, bool uniq, int32_t *d_out, int64_t *d_n_out,
                     uint64_t **keys_out, void *d_ws, size_t ws_bytes, cudaStream_t s) {
    LINR_REQUIRE(bits >= 1 && bits <= 20, "coordinate bit width %d out of range [1,20]", bits);
    LINR_REQUIRE(n >= 0 && n < (1ll << 31), "row count out of range");
    SortWs w;
    if (!carve(d_ws, ws_bytes, n, w)) {
        linr_set_error("workspace too small: have %zu need %zu", ws_bytes, linr_coord_ws_bytes(n));
        return LINR_ENOMEM;
    }
    if (n == 0) {
        if (d_n_out) LINR_CHECK_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(int64_t), s));
        return LINR_OK;
    }
    {
        ProfScope prof(K_COORD, 1, s);
        pack_kernel<<<grid_for(n), TPB, 0, s>>>(d_in, n, bits, shift, w.a);
    }
    size_t cb = w.cub_bytes;
    LINR_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(w.cub, cb, w.a, w.b, (int)n, 0, 3 * bits, s));
    uint64_t *res = w.b;
    if (uniq) {
        cb = w.cub_bytes;
        LINR_CHECK_CUDA(cub::DeviceSelect::Unique(w.cub, cb, w.b, w.a, d_n_out, (int)n, s));
        res = w.a;
    }
    if (d_out) {
        ProfScope prof(K_COORD, 1, s);
        unpack_kernel<<<grid_for(n), TPB, 0, s>>>(res, uniq ? d_n_out : nullptr, n, bits, d_out);
    }
    if (keys_out) *keys_out = res;
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_coord_sort_unique(const int32_t *d_in, int64_t n, int bits, int32_t *d_out, int64_t *d_n_out, void *d_ws,
                           size_t ws_bytes, void *stream) {
    return sort_impl(d_in, n, bits, 0, true, d_out, d_n_out, nullptr, d_ws, ws_bytes, (cudaStream_t)stream);
}

int linr_coord_sort(const int32_t *d_in, int64_t n, int bits, int32_t *d_out, void *d_ws, size_t ws_bytes, void *stream) {
    return sort_impl(d_in, n, bits, 0, false, d_out, nullptr, nullptr, d_ws, ws_bytes, (cudaStream_t)stream);
}

int linr_octree_down(const int32_t *d_child, int64_t nc, int bits, int32_t *d_parent, uint8_t *d_occ, int64_t *d_np,
                     void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    LINR_REQUIRE(nc > 0, "linr_octree_down: empty level");
    uint64_t *pkeys = nullptr;
    int rc = sort_impl(d_child, nc, bits, 1, true, d_parent, d_np, &pkeys, d_ws, ws_bytes, s);
    if (rc) return rc;
    LINR_CHECK_CUDA(cudaMemsetAsync(d_occ, 0, align_up((size_t)nc, 4), s));
    {
        ProfScope prof(K_COORD, 1, s);
        occ_kernel<<<grid_for(nc), TPB, 0, s>>>(d_child, nc, bits, pkeys, d_np, (uint32_t *)d_occ);
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_octree_up_count(const uint8_t *d_occ, int64_t n, int64_t *d_off, void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    LINR_REQUIRE(n > 0 && n < (1ll << 31) - 1, "linr_octree_up_count: bad n");
    SortWs w;
    if (!carve(d_ws, ws_bytes, n + 1, w)) {
        linr_set_error("workspace too small");
        return LINR_ENOMEM;
    }
    int64_t *cnt = (int64_t *)w.a;
    {
        ProfScope prof(K_COORD, 1, s);
        popc_kernel<<<grid_for(n + 1), TPB, 0, s>>>(d_occ, n, cnt);
    }
    size_t cb = w.cub_bytes;
    LINR_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w.cub, cb, cnt, d_off, (int)(n + 1), s));
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_octree_up_expand(const int32_t *d_parent, const uint8_t *d_occ, const int64_t *d_off, int64_t n, int64_t n_child,
                          int bits, int32_t *d_child, void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    LINR_REQUIRE(bits >= 1 && bits <= 20, "bits out of range");
    LINR_REQUIRE(n_child >= 0 && n_child <= 8 * n, "n_child out of range");
    if (n_child == 0) return LINR_OK;
    SortWs w;
    if (!carve(d_ws, ws_bytes, n_child, w)) {
        linr_set_error("workspace too small: have %zu need %zu", ws_bytes, linr_coord_ws_bytes(n_child));
        return LINR_ENOMEM;
    }
    {
        ProfScope prof(K_COORD, 1, s);
        expand_kernel<<<grid_for(n), TPB, 0, s>>>(d_parent, d_occ, d_off, n, bits, w.a);
    }
    size_t cb = w.cub_bytes;
    LINR_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(w.cub, cb, w.a, w.b, (int)n_child, 0, 3 * bits, s));
    {
        ProfScope prof(K_COORD, 1, s);
        unpack_kernel<<<grid_for(n_child), TPB, 0, s>>>(w.b, nullptr, n_child, bits, d_child);
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

size_t linr_hash_bytes(int64_t cap) { return (size_t)cap * sizeof(HashSlot); }

int linr_hash_build(const int32_t *d_xyz, const uint8_t *d_scale, int64_t n, void *d_table, int64_t cap, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    LINR_REQUIRE(cap >= 2 * n && (cap & (cap - 1)) == 0, "hash capacity must be a power of two >= 2n");
    {
        ProfScope prof(K_COORD, 1, s);
        hash_clear_kernel<<<grid_for(cap), TPB, 0, s>>>((HashSlot *)d_table, cap);
    }
    if (n > 0) {
        ProfScope prof(K_COORD, 1, s);
        hash_insert_kernel<<<grid_for(n), TPB, 0, s>>>(d_xyz, d_scale, n, (HashSlot *)d_table, cap);
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_nbr_build(const int32_t *d_xyz, const uint8_t *d_scale, int64_t n, const void *d_table, int64_t cap,
                   int32_t *d_nbr27, int32_t *d_anchor, int64_t ld, uint32_t *d_mask, uint8_t *d_nbr7, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    LINR_REQUIRE(!d_anchor || ld >= n, "anchor leading dimension smaller than n");
    if (n == 0) return LINR_OK;
    {
        ProfScope prof(K_COORD, 1, s);
        nbr_kernel<<<(int)ceil_div64(n, 128), 128, 0, s>>>(d_xyz, d_scale, n, (const HashSlot *)d_table, cap, d_nbr27, d_anchor,
                                                            ld, d_mask, d_nbr7);
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_tile_ranges(const linr_rows *rows, int32_t *d_rng, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    LINR_REQUIRE(rows && d_rng, "linr_tile_ranges: null argument");
    LINR_REQUIRE(rows->ld >= rows->n_rows, "anchor leading dimension smaller than n_rows");
    if (rows->n_rows <= 0) return LINR_OK;
    {
        ProfScope prof(K_COORD, 1, s);
        tile_range_kernel<<<(int)ceil_div64(rows->n_rows, 128), 128, 0, s>>>(rows->d_anchor, rows->ld, rows->d_mask, rows->n_rows, d_rng);
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

void linr_pair_list_order(int32_t *order28) {
    constexpr int8_t t[2][BW3_NW] = LINR_BW3_SLOTS;
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < BW3_NW; ++i) order28[h * BW3_NW + i] = t[h][i];
}

int linr_pair_lists(const linr_rows *rows, int32_t *d_cnt, uint32_t *d_list, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    LINR_REQUIRE(rows && d_cnt && d_list, "linr_pair_lists: null argument");
    LINR_REQUIRE(rows->ld >= rows->n_rows, "anchor leading dimension smaller than n_rows");
    LINR_REQUIRE(rows->n_rows < (1ll << 24), "linr_pair_lists: entries hold 24-bit row indices (n_rows < 16,777,216)");
    if (rows->n_rows <= 0) return LINR_OK;
    {
        ProfScope prof(K_COORD, 1, s);
        pair_list_kernel<<<(int)ceil_div64(rows->n_rows, 256), 288, 0, s>>>(rows->d_anchor, rows->ld, rows->d_mask, rows->n_rows, d_cnt, d_list);
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_hash_lookup(const int32_t *d_q, const uint8_t *d_qs, int64_t nq, const void *d_table, int64_t cap,
                     int32_t *d_rows, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (nq == 0) return LINR_OK;
    {
        ProfScope prof(K_COORD, 1, s);
        hash_lookup_kernel<<<grid_for(nq), TPB, 0, s>>>(d_q, d_qs, nq, (const HashSlot *)d_table, cap, d_rows);
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

}  // extern "C"
