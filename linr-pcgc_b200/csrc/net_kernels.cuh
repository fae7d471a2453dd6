// Device kernels of the occupancy network (fp32, deterministic).  See DESIGN.md "Kernels".
//
// Thread mapping (forward / grad-input): ONE THREAD PER OUTPUT ROW; the per-row arithmetic order is
// fixed (offset column c = 0..8, then dz = -1,0,+1, then input channel) and does not depend on the
// grid, on grouping or on which stage/scale is being processed -> encoder (batched, teacher forced)
// and decoder (sequential) produce bit-identical probabilities.
// Weight gradients (conv, MLP heads, SCE): LANE = ROW with lane-private sums over a row chunk, one
// transposing warp butterfly per chunk, one partial vector per chunk, summed later in chunk order
// (no floating-point atomics anywhere -> bitwise reproducible run to run).
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace linr {

constexpr int MAXG = 8;  // max groups per launch (7 LDFE blocks / 8 heads)

struct Tens {
    float *p;
    int64_t gs;  // group stride (floats)
    int ld;      // row stride (floats)
    int off;     // first column
};
__device__ __forceinline__ float *tptr(const Tens &t, int g, int64_t row) { return t.p + g * t.gs + row * (int64_t)t.ld + t.off; }

struct RowMap {
    const int32_t *anchor;
    int64_t ld;
    const uint32_t *mask;
    int64_t n_rows;
    const int32_t *tile_rng;    // [ceil(n_rows/128)][6] neighbour row ranges per 128-row tile (or null)
    const int32_t *pair_cnt;    // [ceil(n_rows/256)][32] and
    const uint32_t *pair_list;  // [ceil(n_rows/256)][27][256]: (row - tile base) << 24 | neighbour row (or null)
};

template <int C>
__device__ __forceinline__ void load_row(const float *p, float (&v)[C]) {
    static_assert(C % 4 == 0, "rows are float4 multiples");
#pragma unroll
    for (int i = 0; i < C; i += 4) {
        const float4 q = *reinterpret_cast<const float4 *>(p + i);
        v[i] = q.x, v[i + 1] = q.y, v[i + 2] = q.z, v[i + 3] = q.w;
    }
}
// Gathered feature rows (read-only in the kernel that gathers them): one 256-bit load per 8-channel row
// (LDG.E.256 on sm_100a) halves the L1 wavefronts of two strided 128-bit loads.
template <int C>
__device__ __forceinline__ void gather_row(const float *p, float (&v)[C]) {
    if constexpr (C == 8) {
        asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(p));
    } else {
        const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
    }
}
template <int C>
__device__ __forceinline__ void store_row(float *p, const float (&v)[C]) {
#pragma unroll
    for (int i = 0; i < C; i += 4) *reinterpret_cast<float4 *>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

// ------------------------------------------------------------------------------------------------
// 3x3x3 sparse convolution, output-stationary gather.
//   MODE 0: float input [rows,CIN];  MODE 1: input = low `cin` bits of occ[row] (teacher-forced
//   sibling occupancy, values {0,1}); MODE 2: float input + fused head (MLP 8->24->1, sigmoid, bits, CDF).
// ------------------------------------------------------------------------------------------------
struct ConvArgs {
    RowMap map;
    const float *params;
    int w_off[MAXG], b_off[MAXG];  // per group offsets into params; b_off < 0: no bias
    const float *bias_direct;      // single-layer API: bias pointer (overrides b_off)
    int flip;                      // weights are W_f[27][COUT][CIN] of the forward conv; use W'[k][ci][co] = W_f[26-k][co][ci]
    Tens x;
    const uint8_t *occ;
    int cin_base, cin_step;  // MODE 1: cin = cin_base + g * cin_step
    Tens y, res, rmask;      // res.p / rmask.p may be null
    int relu, accum;
    // MODE 2 (head)
    int w1_off[MAXG], b1_off[MAXG], w2_off[MAXG], b2_off[MAXG];
    int stage_base;      // group g predicts octant stage_base + g
    int stage_out_base;  // outputs are written at stage index (stage - stage_out_base)
    float *probs;         // [8,n_rows] or null
    uint16_t *cdf;        // [8,n_rows] or null
    float *dz;            // [8,n_rows] or null (train)
    float dz_scale;       // loss_scale / ln2
    float *bits_partial;  // [gridDim.y * gridDim.x] or null
    int64_t out_ld;       // row count used as the stage stride of probs/cdf/dz
    int bank_off[MAXG];   // CW variant: float offset of each group's staged weights in the constant bank
    // fused kernel_size-1 conv of the Inception block (MODE 0/1 only), same arithmetic order as pw_kernel:
    //  1: y2 = [relu](o @ Wp[COUT,4] + bp [+ res2])                 forward conv1_0 after ConvA, conv1_2 after conv1_1
    //  2: y2 = [mask2 > 0](o[4:8] @ Wp[4,4]^T)                      backward of conv1_2 after ConvB^T
    //  3: o += x2 @ Wp[8,4]^T  (before the ReLU mask of o)          backward of conv1_0 inside conv0_0^T
    int pw_mode, pw_relu;
    int pw_w_off[MAXG], pw_b_off[MAXG];
    Tens y2, res2, x2, mask2;
};

constexpr int CONV_TPB = 128;
// rows per thread: the weights of an offset are read from shared memory once for all of them (the L1/shared
// wavefront pipe, not the FMA pipe, is what saturates first).  Measured (round 1): 4 rows help every 8-channel-input
// conv; the 4-channel-input and bit-input convs are gather-bound and lose occupancy with more rows.
template <int CIN, int COUT, int MODE, bool CW = false>
struct ConvCfg {
    // constant-bank weights cost issue slots of the uniform datapath instead of L1 wavefronts: always 4 rows
    static constexpr int RPT = (CW || (CIN == 8 && MODE != 1) || (CIN == 4 && COUT == 4)) ? 4 : 2;
    static constexpr int ROWS = CONV_TPB * RPT;  // rows per block
};

// Weight bank: the weights of the conv launches of one training call, staged per "fill" (<= 62 KB) into the constant
// bank by cudaMemcpyToSymbolAsync, so that the CW kernel variant reads them as uniform-register FFMA2 operands
// (LDCU) -- no shared-memory broadcasts, which share the L1 data pipe with the gathers (-15 % on the
// 8->8 forward).  One stream per process owns the bank (net.cu); every other caller runs the shared-memory variant,
// which computes the same bits.
constexpr int BANK_FLOATS = 15872;
__constant__ __align__(16) float c_bank[BANK_FLOATS];

// Packed fp32x2 FMA (Blackwell FFMA2): two independent IEEE fp32 FMAs per instruction, so results are bit-identical
// to scalar fmaf; `pack2(x, x)` compiles to the scalar-broadcast operand form (no extra moves).
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// in-place forms: "acc = fma(a, b, acc)" with the accumulator as a read-write operand, so that a predicated use
// compiles to one predicated FFMA2 (a separate destination costs two predicated moves per instruction)
__device__ __forceinline__ void ffma2_acc(u64 &acc, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ void fadd2_acc(u64 &acc, u64 b) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(b)); }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// Plan of the staged ranges of one tile: combine the 128-row sub-tile ranges and lay the three dx ranges out back to
// back (each starting on a 128-byte boundary of the staging area).  ok = false when the ranges do not fit.
struct StagePlan {
    int lo[3], len[3], base[3];
    bool ok;
    __device__ __forceinline__ int delta(int d) const { return base[d] - lo[d]; }  // staged row of neighbour row n = n + delta
    __device__ __forceinline__ uint32_t bytes(int ld) const { return ok ? (uint32_t)(len[0] + len[1] + len[2]) * ld * 4u : 0u; }
};
__device__ __forceinline__ StagePlan plan_ranges(const int32_t *tile_rng, int64_t n_rows, int64_t row0, int tile_rows, int ld, int cap_bytes) {
    StagePlan pl;
    pl.ok = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) pl.lo[d] = 0, pl.len[d] = 0, pl.base[d] = 0;
    if (!tile_rng) return pl;
    int lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {0, 0, 0};
    const int64_t t0 = row0 >> 7, t1 = (min(row0 + tile_rows, n_rows) + 127) >> 7;
    for (int64_t t = t0; t < t1; ++t) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int l = tile_rng[t * 6 + 2 * d], h = tile_rng[t * 6 + 2 * d + 1];
            if (h > l) lo[d] = min(lo[d], l), hi[d] = max(hi[d], h);
        }
    }
    int tot = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (hi[d] <= lo[d]) lo[d] = 0, hi[d] = 0;
        pl.lo[d] = lo[d], pl.len[d] = hi[d] - lo[d], pl.base[d] = tot;
        tot += (pl.len[d] + 3) & ~3;
    }
    pl.ok = (int64_t)tot * ld * 4 <= cap_bytes;
    return pl;
}
// Issue the bulk copies of a plan (the caller has armed `bar` with pl.bytes(ld) among its expected bytes).
__device__ __forceinline__ void issue_ranges(const StagePlan &pl, const float *xg, int ld, float *sx, uint64_t *bar) {
    if (!pl.ok) return;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        if (pl.len[d] > 0) bulk_g2s(sx + (int64_t)pl.base[d] * ld, xg + (int64_t)pl.lo[d] * ld, (uint32_t)pl.len[d] * ld * 4u, bar);
}

// Everything after the 27-offset accumulation of ONE output row `o` (bias, residual / accumulate, fused kernel_size-1
// convs, ReLU / ReLU mask, store; MODE 2: MLP_k + sigmoid + CDF + bits + backward seed).
template <int CIN, int COUT, int MODE>
__device__ __forceinline__ void conv_epilogue_row(const ConvArgs &a, int g, int64_t row, float (&o)[COUT], const float *s_b,
                                                  const float *s_pw, const float *s_head, float &bits) {
#pragma unroll
    for (int co = 0; co < COUT; ++co) o[co] += s_b[co];
    if constexpr (MODE != 2) {
        if (a.res.p) {
            float t[COUT];
            load_row<COUT>(tptr(a.res, g, row), t);
#pragma unroll
            for (int co = 0; co < COUT; ++co) o[co] += t[co];
        }
        if (a.accum) {
            float t[COUT];
            load_row<COUT>(tptr(a.y, g, row), t);
#pragma unroll
            for (int co = 0; co < COUT; ++co) o[co] += t[co];
        }
        if constexpr (COUT == 8) {
            if (a.pw_mode == 3) {
                float xv2[4], t[8];
                load_row<4>(tptr(a.x2, g, row), xv2);
#pragma unroll
                for (int i = 0; i < 8; ++i) t[i] = 0.f;
#pragma unroll
                for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                    for (int co = 0; co < 8; ++co) t[co] = fmaf(xv2[ci], s_pw[ci * 8 + co], t[co]);
#pragma unroll
                for (int co = 0; co < 8; ++co) o[co] = t[co] + o[co];
            }
        }
        if (a.relu) {
#pragma unroll
            for (int co = 0; co < COUT; ++co) o[co] = fmaxf(o[co], 0.f);
        }
        if (a.rmask.p) {
            float t[COUT];
            load_row<COUT>(tptr(a.rmask, g, row), t);
#pragma unroll
            for (int co = 0; co < COUT; ++co) o[co] = t[co] > 0.f ? o[co] : 0.f;
        }
        store_row<COUT>(tptr(a.y, g, row), o);
        if (a.pw_mode == 1) {
            float t[4];
#pragma unroll
            for (int co = 0; co < 4; ++co) t[co] = 0.f;
#pragma unroll
            for (int ci = 0; ci < COUT; ++ci)
#pragma unroll
                for (int co = 0; co < 4; ++co) t[co] = fmaf(o[ci], s_pw[ci * 4 + co], t[co]);
#pragma unroll
            for (int co = 0; co < 4; ++co) t[co] += s_pw[32 + co];
            if (a.res2.p) {
                float u[4];
                load_row<4>(tptr(a.res2, g, row), u);
#pragma unroll
                for (int co = 0; co < 4; ++co) t[co] += u[co];
            }
            if (a.pw_relu) {
#pragma unroll
                for (int co = 0; co < 4; ++co) t[co] = fmaxf(t[co], 0.f);
            }
            store_row<4>(tptr(a.y2, g, row), t);
        } else if (COUT == 8 && a.pw_mode == 2) {
            float t[4], u[4];
#pragma unroll
            for (int co = 0; co < 4; ++co) t[co] = 0.f;
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                for (int co = 0; co < 4; ++co) t[co] = fmaf(o[(COUT == 8 ? 4 : 0) + ci], s_pw[ci * 4 + co], t[co]);
            load_row<4>(tptr(a.mask2, g, row), u);
#pragma unroll
            for (int co = 0; co < 4; ++co) t[co] = u[co] > 0.f ? t[co] : 0.f;
            store_row<4>(tptr(a.y2, g, row), t);
        }
    } else {
        if (a.y.p) store_row<COUT>(tptr(a.y, g, row), o);
        // MLP_k 8 -> 24 -> 1 (models/upsample.py:49-55,156-160); hidden units in pairs, inputs in order
        float z = s_head[240];
#pragma unroll
        for (int j = 0; j < 24; j += 2) {
            u64 h2 = *reinterpret_cast<const u64 *>(s_head + 192 + j);
#pragma unroll
            for (int i = 0; i < 8; ++i) ffma2_acc(h2, pack2(o[i], o[i]), *reinterpret_cast<const u64 *>(s_head + i * 24 + j));
            float h0, h1;
            unpack2(h2, h0, h1);
            z = fmaf(s_head[216 + j], fmaxf(h0, 0.f), z);
            z = fmaf(s_head[217 + j], fmaxf(h1, 0.f), z);
        }
        const float p = 1.f / (1.f + expf(-z));
        const int stage = a.stage_base + g;
        const int64_t oi = (stage - a.stage_out_base) * a.out_ld + row;
        if (a.probs) a.probs[oi] = p;
        if (a.cdf) a.cdf[oi] = (uint16_t)(__float2int_rn(__fmul_rn(__fsub_rn(1.f, p), 65534.f)) + 1);
        if (a.bits_partial || a.dz) {
            const float y = (float)((a.occ[row] >> stage) & 1u);
            const float qv = __fsub_rn(1.f, p);
            // nn.BCELoss clamps log at -100 (models/model_core.py:14)
            const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(qv), -100.f);
            bits += -(y * lp + (1.f - y) * lq) * 1.4426950408889634f;
            if (a.dz) {
                // BCELoss backward (/max((1-p)p, 1e-12)) followed by sigmoid backward (*p(1-p))
                const float pq = qv * p;
                a.dz[oi] = (p - y) / fmaxf(pq, 1e-12f) * pq * a.dz_scale;
            }
        }
    }
}

// Per-group small arrays in shared memory: bias, fused pointwise weights, head MLP (W1 transposed [8][24], b1, w2, b2)
template <int CIN, int COUT, int MODE>
__device__ __forceinline__ void conv_stage_small(const ConvArgs &a, int g, float *s_b, float *s_pw, float *s_head) {
    const int TPB = blockDim.x;
    if constexpr (MODE != 2) {
        if (a.pw_mode && threadIdx.x < 36) {
            const float *wp = a.params + a.pw_w_off[g];
            const int i = threadIdx.x;
            float v = 0.f;
            if (a.pw_mode == 1) {          // s_pw[ci*4 + co] = W[ci][co], ci < COUT;  bias at 32..35
                if (i < COUT * 4) v = wp[i];
                else if (i >= 32) v = a.pw_b_off[g] >= 0 ? a.params[a.pw_b_off[g] + (i - 32)] : 0.f;
            } else if (a.pw_mode == 2) {   // s_pw[ci*4 + co] = W[co][ci]  (ci: channel of o[4:8], co: channel of y2)
                if (i < 16) v = wp[(i & 3) * 4 + (i >> 2)];
            } else {                       // s_pw[co*8 + ci] = W[ci][co]  (co: channel of x2, ci: channel of o)
                if (i < 32) v = wp[(i & 7) * 4 + (i >> 3)];
            }
            s_pw[i] = v;
        }
    }
    if (threadIdx.x < COUT)
        s_b[threadIdx.x] = a.bias_direct ? a.bias_direct[threadIdx.x] : (a.b_off[g] >= 0 ? a.params[a.b_off[g] + threadIdx.x] : 0.f);
    if (MODE == 2) {
        for (int i = threadIdx.x; i < 24 * 8; i += TPB) s_head[(i & 7) * 24 + (i >> 3)] = a.params[a.w1_off[g] + i];
        if (threadIdx.x < 24) {
            s_head[192 + threadIdx.x] = a.params[a.b1_off[g] + threadIdx.x];
            s_head[216 + threadIdx.x] = a.params[a.w2_off[g] + threadIdx.x];
        }
        if (threadIdx.x == 0) s_head[240] = a.params[a.b2_off[g]];
    }
}

template <int CIN, int COUT, int MODE, bool CW = false>
__global__ void __launch_bounds__(CONV_TPB) conv27_kernel(const ConvArgs a) {
    constexpr int WMAX = CW ? 4 : ((MODE == 1) ? 27 * 7 * COUT : 27 * CIN * COUT);
    constexpr int HQ = COUT / 2;  // accumulator pairs per row
    constexpr int CONV_RPT = ConvCfg<CIN, COUT, MODE, CW>::RPT, CONV_ROWS = ConvCfg<CIN, COUT, MODE, CW>::ROWS;
    __shared__ __align__(16) float s_w[WMAX];
    __shared__ float s_b[COUT];
    // head (MODE 2): W1 transposed [8][24] so that hidden-unit pairs are adjacent, then b1[24], w2[24], b2
    __shared__ __align__(16) float s_head[(MODE == 2) ? (24 * 8 + 24 + 24 + 4) : 4];
    __shared__ float s_red[CONV_TPB / 32];
    __shared__ float s_pw[(MODE == 2) ? 1 : 36];   // fused pointwise weights [in][out] + bias
    const int g = blockIdx.y;
    const int cin = (MODE == 1) ? (a.cin_base + g * a.cin_step) : CIN;
    conv_stage_small<CIN, COUT, MODE>(a, g, s_b, s_pw, s_head);
    {
        const float *w = a.params + a.w_off[g];
        const int n = CW ? 0 : 27 * cin * COUT;   // CW: the weights are already in the constant bank
        if (!a.flip) {
            for (int i = threadIdx.x; i < n; i += CONV_TPB) s_w[i] = w[i];
        } else {
            // s_w[k][ci][co] = W_f[26-k][co][ci],  W_f is [27][COUT][CIN]
            for (int i = threadIdx.x; i < n; i += CONV_TPB) {
                const int k = i / (CIN * COUT), r = i % (CIN * COUT), ci = r / COUT, co = r % COUT;
                s_w[i] = w[(26 - k) * CIN * COUT + co * CIN + ci];
            }
        }
    }
    __syncthreads();

    int64_t row[CONV_RPT];
    bool live[CONV_RPT];
    uint32_t m[CONV_RPT];
    u64 acc[CONV_RPT][HQ];
#pragma unroll
    for (int r = 0; r < CONV_RPT; ++r) {
        row[r] = blockIdx.x * (int64_t)CONV_ROWS + r * CONV_TPB + threadIdx.x;
        live[r] = row[r] < a.map.n_rows;
        m[r] = live[r] ? a.map.mask[row[r]] : 0u;
#pragma unroll
        for (int q = 0; q < HQ; ++q) acc[r][q] = 0ull;
    }
#pragma unroll 1
    for (int c = 0; c < 9; ++c) {
        // (measured, round 1: loading the anchors one or two columns ahead, or prefetching the next column's
        // neighbour rows into L2, does not pay -- the gathers themselves are the exposed latency)
        uint32_t m3[CONV_RPT];
        int nb[CONV_RPT];
        uint32_t any = 0;
#pragma unroll
        for (int r = 0; r < CONV_RPT; ++r) {
            m3[r] = (m[r] >> (3 * c)) & 7u;
            any |= m3[r];
            nb[r] = m3[r] ? a.map.anchor[c * a.map.ld + row[r]] : 0;
        }
        if (any == 0) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (!((any >> j) & 1u)) continue;
            const float *wk = s_w + (CW ? 0 : (c + 9 * j) * cin * COUT);
            const int bk = CW ? a.bank_off[g] + (c + 9 * j) * cin * COUT : 0;
            // A row without this neighbour multiplies by zero instead of branching: acc + 0*w == acc bit for bit
            // (acc is never -0: it starts at +0 and x*w + acc rounds an exact zero to +0), so the result does not
            // depend on which rows share a thread -- encoder and decoder batches stay bit-identical.
            float xv[CONV_RPT][CIN];
#pragma unroll
            for (int r = 0; r < CONV_RPT; ++r) {
                const bool on = (m3[r] >> j) & 1u;
#pragma unroll
                for (int i = 0; i < CIN; ++i) xv[r][i] = 0.f;
                if (on) {
                    if (MODE == 1) {
                        const unsigned o = a.occ[nb[r]];
#pragma unroll
                        for (int i = 0; i < 7; ++i) xv[r][i] = ((o >> i) & 1u) ? 1.f : 0.f;
                    } else {
                        gather_row<CIN>(tptr(a.x, g, nb[r]), xv[r]);
                    }
                    ++nb[r];
                }
            }
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                if (MODE == 1 && ci >= cin) break;
                u64 wq[HQ];
                if constexpr (CW) {
                    const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(&c_bank[bk + ci * COUT]);
#pragma unroll
                    for (int q = 0; q < HQ; q += 2) {
                        const ulonglong2 t = w2[q >> 1];
                        wq[q] = t.x, wq[q + 1] = t.y;
                    }
                } else {
                    const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(wk + ci * COUT);
#pragma unroll
                    for (int q = 0; q < HQ; q += 2) {
                        const ulonglong2 t = w2[q >> 1];
                        wq[q] = t.x, wq[q + 1] = t.y;
                    }
                }
#pragma unroll
                for (int r = 0; r < CONV_RPT; ++r) {
                    const u64 xx = pack2(xv[r][ci], xv[r][ci]);
#pragma unroll
                    for (int q = 0; q < HQ; ++q) ffma2_acc(acc[r][q], xx, wq[q]);
                }
            }
        }
    }

    float bits = 0.f;
#pragma unroll
    for (int r = 0; r < CONV_RPT; ++r) {
        if (!live[r]) continue;
        float o[COUT];
#pragma unroll
        for (int q = 0; q < HQ; ++q) unpack2(acc[r][q], o[2 * q], o[2 * q + 1]);
        conv_epilogue_row<CIN, COUT, MODE>(a, g, row[r], o, s_b, s_pw, s_head, bits);
    }
    if (MODE == 2 && a.bits_partial) {
#pragma unroll
        for (int o = 16; o; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = bits;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < CONV_TPB / 32; ++w) t += s_red[w];
            a.bits_partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
    }
}


// Stage the weights of a set of conv launches in the layout their kernels read (forward: as stored; grad-input:
// W'[k][ci][co] = W_f[26-k][co][ci]) into a global buffer, fill after fill; net.cu copies a fill into the bank
// right before the first launch that needs it.
struct BankItem {
    int w_off;   // parameter offset of the [27][..][..] kernel
    int flip;    // grad-input layout
    int cin, cout;  // dims of the CONSUMING launch (gathered channels, produced channels)
    int fill, dst;  // fill index, float offset inside the fill
};
constexpr int BANK_MAX_ITEMS = 64;
struct BankItems {
    int n;
    int fill_base[16];   // float offset of each fill inside the staging buffer
    BankItem it[BANK_MAX_ITEMS];
};
__global__ void __launch_bounds__(256) bank_stage_kernel(const float *__restrict__ params, const BankItems items, float *__restrict__ stage) {
    const BankItem it = items.it[blockIdx.x];
    const float *w = params + it.w_off;
    float *dst = stage + items.fill_base[it.fill] + it.dst;
    const int n = 27 * it.cin * it.cout;
    for (int i = threadIdx.x; i < n; i += 256) {
        if (!it.flip) {
            dst[i] = w[i];
        } else {
            const int k = i / (it.cin * it.cout), r = i % (it.cin * it.cout), ci = r / it.cout, co = r % it.cout;
            dst[i] = w[(26 - k) * it.cin * it.cout + co * it.cin + ci];
        }
    }
}


// Sum N (= 32 or 64) lane-private values over the 32 lanes of a warp with a transposing butterfly: 31 (63) shuffles
// instead of 5 per value.  Afterwards lane l holds the totals of elements l*N/32 .. l*N/32 + N/32 - 1 in v[0 .. N/32).
// The pairing order is fixed, so the result is bitwise reproducible.
template <int N>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[N], int lane) {
    static_assert(N == 32 || N == 64, "padded accumulator count");
#pragma unroll
    for (int o = 16, H = N / 2; o; o >>= 1, H >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float keep = up ? v[i + H] : v[i], send = up ? v[i] : v[i + H];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient of the 3x3x3 conv: dW[k][ci][co] = sum_o x[nbr(o,k)][ci] * dy[o][co], db = sum_o dy[o].
// Mapping: ONE WARP PER KERNEL OFFSET (slot), LANE = ROW.  A warp walks its row chunk 32 consecutive rows at a time:
// the kernel-map words, the dy rows and the gathered neighbour rows of 32 consecutive output rows are (nearly)
// consecutive in memory, so every load is coalesced and an L1 wavefront serves ~4 rows (the previous mapping, a warp
// across 16 offsets of ONE row, touched ~one 128-byte line per useful pair and was bound by the L1 wavefront pipe).
// Every lane keeps a private dW[k] (CI x COUT, packed FFMA2) over its rows; no shared memory, no barriers; loads of
// step s+1 are issued before the FMAs of step s.  At the end the 32 lane-private copies are summed by a transposing
// butterfly (fixed order), one partial vector per (row chunk, offset) -- summed over chunks by finalize_grad_kernel.
// Slot 27/NK (an extra warp) sums dy for the bias gradient.  No floating-point atomics anywhere.
// ------------------------------------------------------------------------------------------------
struct BwdWArgs {
    RowMap map;
    int w_off[MAXG], b_off[MAXG];
    Tens x, dy;
    const uint8_t *occ;
    int cin_base, cin_step;
    float *partial;  // [n_chunks][P]
    int64_t P;
    int64_t chunk;  // rows per chunk (multiple of 32)
    int no_xstage;  // v3: gather x through L1 instead of staging its row ranges
    int groups;     // v3: a block walks its chunk once per group (one long TMA pipeline instead of `groups` short blocks)
};
constexpr int BW_T = 256;  // chunk granularity (rows): whole pair-list tiles

template <int CIN, int COUT, int MODE>
struct BwdWCfg {
    static constexpr int CI = (MODE == 1) ? 7 : CIN;                    // accumulated input width (bit inputs: up to 7)
    static constexpr int NK = (CIN == 4 && COUT == 4) ? 3 : 1;           // offsets per thread: the dz planes of its column
    static constexpr int SLOTS = 27 / NK + 1;                            // warps of work per (chunk, group); last = bias
    // warps per block (measured: 7 for the 8->8 float conv -- dy sharing matters most; 4 for 8->4 and bit inputs)
    static constexpr int WPB = (NK == 3) ? 5 : ((CIN == 8 && COUT == 8 && MODE == 0) ? 7 : 4);
    static constexpr int GX = (SLOTS + WPB - 1) / WPB;                   // blocks per (chunk, group)
    static constexpr int TPB = 32 * WPB;
    // register cap: 128 -> 4 warps per SM sub-partition (16 K registers each), 80 -> 6
    static constexpr int MAXREG = (CIN == 8 && COUT == 4) ? 80 : 128;
    static constexpr int XW = (MODE == 1) ? 1 : CIN;                     // registers of one gathered neighbour
    static constexpr int DEPTH = (CIN == 8 && COUT == 4) ? 2 : 3;        // software pipeline depth (row steps)
    static constexpr int V = NK * CI * COUT;                             // accumulators per thread
    static constexpr int VP = V <= 32 ? 32 : 64;                         // padded for the butterfly
    static_assert(V <= 64, "accumulator tile");
};

template <int C>
__device__ __forceinline__ void load_row_nc2(const float *p, u64 (&v)[C / 2]) {
    if constexpr (C == 8) {
        asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
    } else {
        asm("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "l"(p));
    }
}

template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(BwdWCfg<CIN, COUT, MODE>::TPB) __maxnreg__((BwdWCfg<CIN, COUT, MODE>::MAXREG)) conv27_bwd_w_kernel(const BwdWArgs a) {
    using Cfg = BwdWCfg<CIN, COUT, MODE>;
    constexpr int CI = Cfg::CI, NK = Cfg::NK, XW = Cfg::XW, HQ = COUT / 2, V = Cfg::V, VP = Cfg::VP;
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * Cfg::WPB + (threadIdx.x >> 5);
    if (slot >= Cfg::SLOTS) return;
    const int g = blockIdx.z;
    const int cin = (MODE == 1) ? (a.cin_base + g * a.cin_step) : CIN;
    const int64_t r0 = blockIdx.y * a.chunk;
    const int64_t r1 = min(r0 + a.chunk, a.map.n_rows);
    float *out = a.partial + blockIdx.y * a.P;
    const float *dyg = a.dy.p + g * a.dy.gs + a.dy.off;
    const int dld = a.dy.ld;

    if (slot == Cfg::SLOTS - 1) {  // ---- bias gradient: column sums of dy over the chunk
        if (a.b_off[g] < 0) return;
        u64 bs[HQ];
#pragma unroll
        for (int q = 0; q < HQ; ++q) bs[q] = 0ull;
        for (int64_t r = r0 + lane; r < r1; r += 32) {
            u64 d[HQ];
            load_row_nc2<COUT>(dyg + r * dld, d);
#pragma unroll
            for (int q = 0; q < HQ; ++q) fadd2_acc(bs[q], d[q]);
        }
        float v[COUT];
#pragma unroll
        for (int q = 0; q < HQ; ++q) unpack2(bs[q], v[2 * q], v[2 * q + 1]);
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
#pragma unroll
            for (int o = 16; o; o >>= 1) v[co] += __shfl_xor_sync(0xffffffffu, v[co], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int co = 0; co < COUT; ++co) out[a.b_off[g] + co] = v[co];
        }
        return;
    }

    // ---- weight gradient of offsets k = c + 9*(j0 + nk), nk < NK
    const int c = (NK == 3) ? slot : slot % 9, j0 = (NK == 3) ? 0 : slot / 9;
    const int32_t *anch = a.map.anchor + c * a.map.ld;
    const float *xg = (MODE == 1) ? nullptr : (a.x.p + g * a.x.gs + a.x.off);
    const int xld = a.x.ld;
    u64 acc[NK][CI][HQ];
#pragma unroll
    for (int n = 0; n < NK; ++n)
#pragma unroll
        for (int i = 0; i < CI; ++i)
#pragma unroll
            for (int q = 0; q < HQ; ++q) acc[n][i][q] = 0ull;

    // stage A (kernel-map words) runs two steps ahead, stage B (neighbour rows, dy) one step ahead of the FMAs
    auto loadA = [&](int64_t r, uint32_t &m3, int &an) {
        m3 = 0u, an = 0;
        if (r < r1) {
            m3 = (a.map.mask[r] >> (3 * c)) & 7u;
            an = anch[r];
        }
    };
    auto loadB = [&](int64_t r, uint32_t m3, int an, float (&xb)[NK][XW], u64 (&db)[HQ]) {
#pragma unroll
        for (int q = 0; q < HQ; ++q) db[q] = 0ull;
        if (r < r1) load_row_nc2<COUT>(dyg + r * dld, db);
#pragma unroll
        for (int n = 0; n < NK; ++n) {
            const int j = j0 + n;
#pragma unroll
            for (int i = 0; i < XW; ++i) xb[n][i] = 0.f;
            if ((m3 >> j) & 1u) {
                const int nb = an + __popc(m3 & ((1u << j) - 1u));
                if constexpr (MODE == 1) xb[n][0] = __uint_as_float((unsigned)a.occ[nb]);
                else gather_row<CIN>(xg + (int64_t)nb * xld, xb[n]);
            }
        }
    };
    auto fma = [&](const float (&xb)[NK][XW], const u64 (&db)[HQ]) {
#pragma unroll
        for (int n = 0; n < NK; ++n) {
#pragma unroll
            for (int i = 0; i < CI; ++i) {
                float xs;
                if (MODE == 1) xs = ((__float_as_uint(xb[n][0]) >> i) & 1u) ? 1.f : 0.f;
                else xs = xb[n][i];
                const u64 xx = pack2(xs, xs);
#pragma unroll
                for (int q = 0; q < HQ; ++q) ffma2_acc(acc[n][i][q], xx, db[q]);
            }
        }
    };

    // Software pipeline over row steps (32 rows each): ring of D (neighbour row, dy row) buffers, loaded D-1 steps
    // before their FMAs; the kernel-map words those loads need (mask, anchor) sit in their own ring and are loaded D
    // steps before that, so no load is consumed less than ~2 steps after it was issued.
    constexpr int D = Cfg::DEPTH;
    float xb[D][NK][XW];
    u64 db[D][HQ];
    uint32_t mA[D];
    int aA[D];
    int64_t r = r0 + lane;
#pragma unroll
    for (int u = 0; u < D; ++u) loadA(r + 32 * u, mA[u], aA[u]);
#pragma unroll
    for (int u = 0; u < D - 1; ++u) {
        loadB(r + 32 * u, mA[u], aA[u], xb[u], db[u]);
        loadA(r + 32 * (u + D), mA[u], aA[u]);
    }
#pragma unroll 1
    for (; r - lane < r1; r += 32 * D) {  // warp-uniform trip count; rows past the chunk end contribute exact zeros
#pragma unroll
        for (int u = 0; u < D; ++u) {
            const int v = (u + D - 1) % D;   // ring slot of step (current + D - 1)
            loadB(r + 32 * (u + D - 1), mA[v], aA[v], xb[v], db[v]);
            loadA(r + 32 * (u + 2 * D - 1), mA[v], aA[v]);
            fma(xb[u], db[u]);
        }
    }

    // ---- sum the 32 lane-private copies: transposing butterfly, lane ends up with elements lane*VP/32 + i
    float v[VP];
#pragma unroll
    for (int e = 0; e < VP; ++e) v[e] = 0.f;
#pragma unroll
    for (int n = 0; n < NK; ++n)
#pragma unroll
        for (int i = 0; i < CI; ++i)
#pragma unroll
            for (int q = 0; q < HQ; ++q) unpack2(acc[n][i][q], v[(n * CI + i) * COUT + 2 * q], v[(n * CI + i) * COUT + 2 * q + 1]);
    warp_transpose_reduce<VP>(v, lane);
#pragma unroll
    for (int i = 0; i < VP / 32; ++i) {
        const int e = lane * (VP / 32) + i;
        if (e < V) {
            const int n = e / (CI * COUT), ci = (e / COUT) % CI, co = e % COUT;
            const int k = c + 9 * (j0 + n);
            if (ci < cin) out[a.w_off[g] + k * cin * COUT + ci * COUT + co] = v[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient, v3: staged + pair lists.  Same sums as conv27_bwd_w_kernel, different schedule:
//  * a block owns one row chunk of one group and half of the 28 slots (27 offsets + bias), ONE WARP PER SLOT, plus a
//    producer warp.  The chunk is walked in 256-row tiles through a ring of shared-memory stages that the producer
//    fills with bulk (TMA) copies: the dy rows of the tile and the three neighbour row ranges of x
//    (RowMap::tile_rng).  `full` mbarriers count the bytes of a stage, `empty` mbarriers the consumer warps that are
//    done with it.  No block-wide barrier inside the loop.
//  * a warp does not multiply by zero for absent neighbours: it walks the pair list of (tile, its offset)
//    (RowMap::pair_list) 32 pairs at a time, every lane one (dy row, neighbour row) pair, both read from shared
//    memory -- 14.3 of 27 slots are occupied on a surface, so this is ~1.9x fewer FMAs and loads than lane = row.
//    Offsets differ in density (centre 1.0, faces .65, edges .51, corners .41): the slot -> warp table balances the
//    sums per SM sub-partition (warp w runs on sub-partition w % 4).
//  * lane-private dW[k] (CI x COUT, packed FFMA2) over the warp's pairs, transposing butterfly at the end of the chunk,
//    one partial per (chunk, offset): fixed order, no floating-point atomics -> bitwise reproducible run to run.
// ------------------------------------------------------------------------------------------------
__constant__ int8_t c_bw3_slot[2][BW3_NW] = LINR_BW3_SLOTS;   // common.cuh

// DLD: row stride of dy in floats (COUT, or 8 when a 4-channel slice of an 8-wide tensor is the gradient)
template <int CIN, int COUT, int MODE, int DLD>
struct BwdW3Cfg {
    static constexpr int CI = (MODE == 1) ? 7 : CIN;
    static constexpr int XROWS = 1024;                                                   // staged neighbour rows per tile
    static constexpr int XS = (MODE == 1) ? 0 : (XROWS + 8) * CIN * 4;                   // + a zero row for idle lanes
    static constexpr int DYB = (BW3_T + 8) * DLD * 4;                                    // dy rows + a zero row
    static constexpr int LSB = BW3_NW * 1024;                                            // the pair lists of the block's slots
    static constexpr int STAGE = DYB + XS + LSB;
    static constexpr int NST = (MODE == 1) ? 6 : 4;
    static constexpr int SMEM = NST * STAGE + 256;
    static constexpr int V = CI * COUT, VP = V <= 32 ? 32 : 64;
    static_assert(STAGE % 128 == 0, "stages keep the 128-byte alignment of the bank swizzle");
};

// One staged row as raw 16-byte pieces.  8-channel rows are 32 bytes: eight lanes reading the first halves of eight
// consecutive rows would hit only four of the eight 16-byte bank groups.  So the ORDER of a lane's two loads depends on
// the lane: lanes with bit 2 set read the upper half first (b0 / b1 are the row base plus the half the lane reads
// first / second).  A quarter warp then covers all eight bank groups when its rows are consecutive, and because the
// order is a constant of the lane nothing is selected per row: the lane's sums are simply kept in its own channel
// order (channel c of the lane = channel c ^ 4 of the row for those lanes) and put right once, in flush().
template <int C>
struct RawRow {
    float4 u[C / 4];
};
template <int C>
__device__ __forceinline__ void raw_load(const float *b0, const float *b1, int p, int ld, RawRow<C> &r) {
    r.u[0] = *reinterpret_cast<const float4 *>(b0 + p * ld);
    if constexpr (C == 8) r.u[1] = *reinterpret_cast<const float4 *>(b1 + p * ld);
}
// ... as floats, in the lane's channel order
template <int C>
__device__ __forceinline__ void raw_resolve(const RawRow<C> &r, float (&v)[C]) {
    v[0] = r.u[0].x, v[1] = r.u[0].y, v[2] = r.u[0].z, v[3] = r.u[0].w;
    if constexpr (C == 8) v[4] = r.u[1].x, v[5] = r.u[1].y, v[6] = r.u[1].z, v[7] = r.u[1].w;
}

template <int CIN, int COUT, int MODE, int DLD>
__global__ void __launch_bounds__(32 * (BW3_NW + 1), 1) conv27_bwd_w3_kernel(const BwdWArgs a) {
    using Cfg = BwdW3Cfg<CIN, COUT, MODE, DLD>;
    constexpr int CI = Cfg::CI, HQ = COUT / 2, V = Cfg::V, VP = Cfg::VP, T = BW3_T, NST = Cfg::NST;
    constexpr int XW = (MODE == 1) ? 1 : CIN;
    extern __shared__ __align__(128) unsigned char bw3_smem[];
    __shared__ int s_plan[NST][4 + BW3_NW + 2];   // x-range shifts, staged flag, byte offset of every warp's list
    uint64_t *full = reinterpret_cast<uint64_t *>(bw3_smem + NST * Cfg::STAGE), *empty = full + NST;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = blockIdx.y * a.chunk;   // chunks are whole 256-row tiles
    const int64_t r1 = min(r0 + a.chunk, a.map.n_rows);
    const int n_tiles = r1 > r0 ? (int)((r1 - r0 + T - 1) / T) : 0;
    const int n_vt = n_tiles * a.groups;        // virtual tiles: the chunk once per group, group-major
    const int64_t tile0 = r0 / T;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NST; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], BW3_NW);
        mbar_init_fence();
    }
    // zero rows (dy row T, x row XROWS of every stage): where idle lanes of a partial batch point
    for (int i = threadIdx.x; i < NST * 8; i += blockDim.x) {
        unsigned char *sb = bw3_smem + (i >> 3) * Cfg::STAGE;
        if ((i & 7) < DLD) reinterpret_cast<float *>(sb)[T * DLD + (i & 7)] = 0.f;
        if (MODE != 1 && (i & 7) < CIN) reinterpret_cast<float *>(sb + Cfg::DYB)[Cfg::XROWS * CIN + (i & 7)] = 0.f;
    }
    __syncthreads();

    if (warp == BW3_NW) {   // ---- producer warp: NST - 1 tiles ahead of the slowest consumer
        // Lane 0 issues the (at most) five bulk copies of a tile: dy rows, the block's 14 pair lists (contiguous in global
        // memory, coords.cu) and the three x ranges -- a bulk copy costs its issuing warp ~170 cycles, and with one copy
        // per list (18 per tile) the producer was the pace of the whole kernel.  What the copies need from global memory
        // (the 14 list lengths, the 2 x 6 range words of the tile's two 128-row sub-tiles) is fetched one tile ahead.
        const int my_slot = lane < BW3_NW ? c_bw3_slot[blockIdx.x][lane] : 27;
        const int64_t nt128 = (a.map.n_rows + 127) >> 7;
        auto fetch = [&](int t, uint32_t &lbv, int &rv) {
            lbv = my_slot < 27 ? (uint32_t)((a.map.pair_cnt[(tile0 + t) * 32 + my_slot] + 3) & ~3) * 4u : 0u;
            rv = 0;
            if constexpr (MODE != 1) {
                if (lane < 12 && !a.no_xstage && a.map.tile_rng) {
                    const int64_t sub = ((r0 + (int64_t)t * T) >> 7) + lane / 6;
                    if (sub < nt128) rv = a.map.tile_rng[sub * 6 + lane % 6];
                }
            }
        };
        uint32_t lb = 0, lb_next = 0;
        int rv = 0, rv_next = 0;
        if (n_vt > 0) fetch(0, lb, rv);
        for (int vt = 0; vt < n_vt; ++vt) {
            const int g = vt / n_tiles, t = vt - g * n_tiles;
            const int st = vt % NST;
            const int64_t row0 = r0 + (int64_t)t * T;
            const int nrow = (int)min((int64_t)T, r1 - row0);
            unsigned char *sb = bw3_smem + st * Cfg::STAGE;
            if (vt + 1 < n_vt) fetch((t + 1 == n_tiles) ? 0 : t + 1, lb_next, rv_next);
            // the three neighbour ranges of the tile, laid out back to back (every lane computes the same plan)
            StagePlan pl;
            pl.ok = false;
            if constexpr (MODE != 1) {
                int tot = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int l0 = __shfl_sync(0xffffffffu, rv, 2 * d), h0 = __shfl_sync(0xffffffffu, rv, 2 * d + 1);
                    const int l1 = __shfl_sync(0xffffffffu, rv, 6 + 2 * d), h1 = __shfl_sync(0xffffffffu, rv, 7 + 2 * d);
                    int lo = INT32_MAX, hi = 0;
                    if (h0 > l0) lo = l0, hi = h0;
                    if (h1 > l1) lo = min(lo, l1), hi = max(hi, h1);
                    if (hi <= lo) lo = 0, hi = 0;
                    pl.lo[d] = lo, pl.len[d] = hi - lo, pl.base[d] = tot;
                    tot += (pl.len[d] + 3) & ~3;
                }
                pl.ok = !a.no_xstage && a.map.tile_rng && tot > 0 && tot <= Cfg::XROWS;
            }
            // the block's 14 lists lie back to back in global memory (coords.cu): one copy; lane w keeps where list w starts
            uint32_t lend = lb;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, lend, o);
                if (lane >= o) lend += v;
            }
            const uint32_t lsum = __shfl_sync(0xffffffffu, lend, BW3_NW - 1);
            if (vt >= NST) mbar_wait(&empty[st], ((vt / NST) - 1) & 1);   // every consumer warp is done with virtual tile vt - NST
            if (lane < BW3_NW) s_plan[st][4 + lane] = (int)(lend - lb);
            __syncwarp();   // lane 0's arrive below publishes these stores with its own
            if (lane == 0) {
                s_plan[st][0] = pl.ok ? pl.delta(0) : 0, s_plan[st][1] = pl.ok ? pl.delta(1) : 0, s_plan[st][2] = pl.ok ? pl.delta(2) : 0;
                s_plan[st][3] = pl.ok;
                const uint32_t dyb = (uint32_t)nrow * DLD * 4u;
                mbar_arrive_expect_tx(&full[st], dyb + lsum + (MODE != 1 ? pl.bytes(CIN) : 0u));
                bulk_g2s(sb, a.dy.p + g * a.dy.gs + a.dy.off + row0 * DLD, dyb, &full[st]);
                if (lsum) bulk_g2s(sb + Cfg::DYB + Cfg::XS, a.map.pair_list + (tile0 + t) * PAIR_TILE_ENTRIES + blockIdx.x * PAIR_HALF_ENTRIES, lsum, &full[st]);
                if constexpr (MODE != 1) issue_ranges(pl, a.x.p + g * a.x.gs + a.x.off, CIN, reinterpret_cast<float *>(sb + Cfg::DYB), &full[st]);
            }
            lb = lb_next, rv = rv_next;
        }
        return;
    }

    const int slot = c_bw3_slot[blockIdx.x][warp];
    float *out = a.partial + blockIdx.y * a.P;
    const bool is_bias = slot == 27;
    const int dxi = slot % 3;   // offset k = slot = c + 9 j, dx index = c % 3 = k % 3
    const int hs = (lane >> 2) & 1;                          // which half of a 32-byte row this lane reads first
    const int dh0 = (COUT == 8) ? 4 * hs : 0, dh1 = (COUT == 8) ? 4 * (hs ^ 1) : 0;
    const int xh0 = (XW == 8) ? 4 * hs : 0, xh1 = (XW == 8) ? 4 * (hs ^ 1) : 0;
    u64 acc[CI][HQ];
#pragma unroll
    for (int i = 0; i < CI; ++i)
#pragma unroll
        for (int q = 0; q < HQ; ++q) acc[i][q] = 0ull;

    // FMAs of one pair per lane
    auto fma_pair = [&](const float (&d)[COUT], const float (&xb)[XW]) {
        u64 db[HQ];
#pragma unroll
        for (int q = 0; q < HQ; ++q) db[q] = pack2(d[2 * q], d[2 * q + 1]);
#pragma unroll
        for (int i = 0; i < CI; ++i) {
            float xv;
            if (MODE == 1) xv = ((__float_as_uint(xb[0]) >> i) & 1u) ? 1.f : 0.f;
            else xv = xb[i];
            const u64 xx = pack2(xv, xv);
#pragma unroll
            for (int q = 0; q < HQ; ++q) ffma2_acc(acc[i][q], xx, db[q]);
        }
    };
    // the sum for (input channel i, output channel pair q) out of the lane's own channel order (see RawRow)
    auto lane_order = [&](int i, int q) -> u64 {
        constexpr int IX = (XW == 8) ? 4 : 0, QX = (COUT == 8) ? 2 : 0;
        if constexpr (IX == 0 && QX == 0) return acc[i][q];
        else return hs ? acc[i ^ IX][q ^ QX] : acc[i][q];
    };
    // end of a group's pass over the chunk: reduce the lane-private sums, store the partial, start over
    auto flush = [&](int g) {
        const int cin = (MODE == 1) ? (a.cin_base + g * a.cin_step) : CIN;
        if (is_bias) {
            if (a.b_off[g] >= 0) {
                float v[COUT];
#pragma unroll
                for (int q = 0; q < HQ; ++q) unpack2((COUT == 8 && hs) ? acc[0][q ^ (HQ / 2)] : acc[0][q], v[2 * q], v[2 * q + 1]);
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) v[co] += __shfl_xor_sync(0xffffffffu, v[co], o);
                }
                if (lane == 0) {
#pragma unroll
                    for (int co = 0; co < COUT; ++co) out[a.b_off[g] + co] = v[co];
                }
            }
        } else {
            float v[VP];
#pragma unroll
            for (int e = 0; e < VP; ++e) v[e] = 0.f;
#pragma unroll
            for (int i = 0; i < CI; ++i)
#pragma unroll
                for (int q = 0; q < HQ; ++q) unpack2(lane_order(i, q), v[i * COUT + 2 * q], v[i * COUT + 2 * q + 1]);
            warp_transpose_reduce<VP>(v, lane);
#pragma unroll
            for (int i = 0; i < VP / 32; ++i) {
                const int e = lane * (VP / 32) + i;
                if (e < V) {
                    const int ci = e / COUT, co = e % COUT;
                    if (ci < cin) out[a.w_off[g] + slot * cin * COUT + ci * COUT + co] = v[i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < CI; ++i)
#pragma unroll
            for (int q = 0; q < HQ; ++q) acc[i][q] = 0ull;
    };

    if (n_tiles == 0) {   // a chunk past the end of the rows: its partials are zeros
        for (int g = 0; g < a.groups; ++g) flush(g);
        return;
    }
    int cnt_next = !is_bias ? a.map.pair_cnt[tile0 * 32 + slot] : 0;
    int t = 0, g = 0;
#pragma unroll 1
    for (int vt = 0; vt < n_vt; ++vt) {
        const int st = vt % NST;
        const int cnt = cnt_next;
        const int tn = (t + 1 == n_tiles) ? 0 : t + 1;
        if (!is_bias) cnt_next = a.map.pair_cnt[(tile0 + tn) * 32 + slot];
        mbar_wait(&full[st], (vt / NST) & 1);
        const unsigned char *sb = bw3_smem + st * Cfg::STAGE;
        const float *sdy = reinterpret_cast<const float *>(sb);
        if (is_bias) {
            const int64_t row0 = r0 + (int64_t)t * T;
            const int nrow = (int)min((int64_t)T, r1 - row0);
            for (int rl = lane; rl < nrow; rl += 32) {
                RawRow<COUT> rr;
                float d[COUT];
                raw_load<COUT>(sdy + dh0, sdy + dh1, rl, DLD, rr);
                raw_resolve<COUT>(rr, d);
#pragma unroll
                for (int q = 0; q < HQ; ++q) fadd2_acc(acc[0][q], pack2(d[2 * q], d[2 * q + 1]));
            }
        } else if ((MODE == 1) || s_plan[st][3]) {
            // every operand in shared memory (MODE 1: the neighbour's occupancy byte comes from global memory).
            // Software pipeline: the rows of batch b+1 are loaded before the FMAs of batch b.
            // Addresses are 32-bit shared-window byte offsets; the per-tile terms (stage, the lane's halves, the shift of
            // the offset's neighbour range) are folded into four bases, so a pair costs a shift/mask and an add per row.
            constexpr int DSH = (DLD == 8) ? 5 : 4, XSH = (CIN == 8) ? 5 : 4;
            const int dlt = s_plan[st][dxi];
            const uint32_t a_l = smem_u32(sb + Cfg::DYB + Cfg::XS) + (uint32_t)s_plan[st][4 + warp] + (uint32_t)lane * 4u;
            const uint32_t a_d0 = smem_u32(sdy) + dh0 * 4u, a_d1 = smem_u32(sdy) + dh1 * 4u;
            const uint32_t a_x = smem_u32(sb + Cfg::DYB) + ((uint32_t)dlt << XSH);
            const uint32_t a_x0 = a_x + xh0 * 4u, a_x1 = a_x + xh1 * 4u;
            const uint32_t null_x = (uint32_t)(Cfg::XROWS - dlt) << XSH;   // the zero row, for idle lanes of the last batch
            RawRow<COUT> rd;
            RawRow<(MODE == 1) ? 4 : CIN> rx;
            unsigned ob = 0;
            auto load_batch = [&](int b) {
                const bool on = b + lane < cnt;
                const uint32_t e = lds_u32(a_l + (uint32_t)b * 4u);   // inside the warp's own 256-entry slot either way
                const uint32_t od = on ? ((e >> (24 - DSH)) & (0xffu << DSH)) : ((uint32_t)T << DSH);
                rd.u[0] = lds_f4(a_d0 + od);
                if constexpr (COUT == 8) rd.u[1] = lds_f4(a_d1 + od);
                if constexpr (MODE == 1) {
                    ob = on ? (unsigned)a.occ[e & 0xffffffu] : 0u;
                } else {
                    const uint32_t xo = on ? ((e & 0xffffffu) << XSH) : null_x;
                    rx.u[0] = lds_f4(a_x0 + xo);
                    if constexpr (CIN == 8) rx.u[1] = lds_f4(a_x1 + xo);
                }
            };
            if (cnt > 0) load_batch(0);
#pragma unroll 1
            for (int b = 0; b < cnt; b += 32) {
                float d[COUT], xb[XW];
                raw_resolve<COUT>(rd, d);
                if constexpr (MODE == 1) xb[0] = __uint_as_float(ob);
                else raw_resolve<CIN>(rx, xb);
                if (b + 32 < cnt) load_batch(b + 32);
                fma_pair(d, xb);
            }
        } else {
            // the neighbour ranges of this tile did not fit the staging area: gather x through L1
            const uint32_t *lst = reinterpret_cast<const uint32_t *>(sb + Cfg::DYB + Cfg::XS + s_plan[st][4 + warp]);
            const float *xg = (MODE == 1) ? nullptr : (a.x.p + g * a.x.gs + a.x.off);
#pragma unroll 1
            for (int b = 0; b < cnt; b += 32) {
                const bool on = b + lane < cnt;
                const uint32_t e = on ? lst[b + lane] : 0u;
                float d[COUT], xb[XW];
                RawRow<COUT> rd;
                raw_load<COUT>(sdy + dh0, sdy + dh1, on ? (int)(e >> 24) : T, DLD, rd);
                raw_resolve<COUT>(rd, d);
#pragma unroll
                for (int i = 0; i < XW; ++i) xb[i] = 0.f;
                if constexpr (MODE != 1) {
                    if (on) {   // the two halves in the lane's order, like the staged rows
                        const float *xr = xg + (int64_t)(e & 0xffffffu) * CIN;
                        RawRow<CIN> rx;
                        rx.u[0] = __ldg(reinterpret_cast<const float4 *>(xr + xh0));
                        if constexpr (CIN == 8) rx.u[1] = __ldg(reinterpret_cast<const float4 *>(xr + xh1));
                        raw_resolve<CIN>(rx, xb);
                    }
                }
                fma_pair(d, xb);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++t == n_tiles) {
            flush(g);
            t = 0, ++g;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Pointwise (kernel_size = 1) convs of the Inception block.  Forward and grad-input run in the epilogues of the
// 27-offset convs (pw_mode of conv27_kernel); only the weight gradients are separate launches.
// ------------------------------------------------------------------------------------------------
// dW[ci][co] = sum_r x[r][ci] dy[r][co], db[co] = sum_r dy[r][co]; one partial per row chunk.
struct PwBwdWArgs {
    int64_t n_rows;
    int w_off[MAXG], b_off[MAXG];
    Tens x, dy;
    float *partial;
    int64_t P, chunk;
};
constexpr int PWW_TPB = 256;

template <int N>
__device__ __forceinline__ void block_reduce_store(float (&acc)[N], float *s_red /* [TPB/32][N] */, int tpb) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[i] = v;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) s_red[warp * N + i] = acc[i];
    }
    __syncthreads();
    (void)tpb;
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(PWW_TPB) pw_bwd_w_kernel(const PwBwdWArgs a) {
    constexpr int N = CIN * COUT + COUT;
    __shared__ float s_red[(PWW_TPB / 32) * N];
    const int g = blockIdx.y;
    const int64_t r0 = blockIdx.x * a.chunk, r1 = min(r0 + a.chunk, a.n_rows);
    float acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.f;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += PWW_TPB) {
        float xv[CIN], dyv[COUT];
        load_row<CIN>(tptr(a.x, g, r), xv);
        load_row<COUT>(tptr(a.dy, g, r), dyv);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[ci * COUT + co] = fmaf(xv[ci], dyv[co], acc[ci * COUT + co]);
        }
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[CIN * COUT + co] += dyv[co];
    }
    block_reduce_store<N>(acc, s_red, PWW_TPB);
    float *out = a.partial + blockIdx.x * a.P;
    for (int i = threadIdx.x; i < N; i += PWW_TPB) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < PWW_TPB / 32; ++w) t += s_red[w * N + i];
        if (i < CIN * COUT) out[a.w_off[g] + i] = t;
        else out[a.b_off[g] + (i - CIN * COUT)] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// Scale-context extraction (models/model_core.py:46-53): f0 = W2_s relu(W1_s [emb_s | nbr7] + b1_s) + b2_s.
// The embedding part of layer 1 is folded into a per-scale effective bias.
// ------------------------------------------------------------------------------------------------
constexpr int MAXS = 12;
struct SceArgs {
    int64_t n_rows;
    int scale_num;
    const float *params;
    int emb_off;
    int w1_off[MAXS], b1_off[MAXS], w2_off[MAXS], b2_off[MAXS];
    const uint8_t *nbr7, *scale;
    int scale_fixed;  // >= 0: every row belongs to this scale (decoder path, scale array null)
    Tens f0;          // fwd out
    Tens df0;         // bwd in
    float *partial;
    int64_t P, chunk;
};
constexpr int SCE_TPB = 256;
constexpr int SCE_SM = 16 + 16 * 7 + 8 * 16 + 8;  // b1eff, W1n[16][7], W2[8][16], b2

__device__ __forceinline__ void sce_stage_scale(const SceArgs &a, int s, float *dst) {
    // dst: [b1eff 16][W1n 16*7][W2 8*16][b2 8]
    const float *w1 = a.params + a.w1_off[s];  // [16][15]
    for (int i = threadIdx.x; i < SCE_SM; i += blockDim.x) {
        float v;
        if (i < 16) {
            v = a.params[a.b1_off[s] + i];
            for (int e = 0; e < 8; ++e) v = fmaf(w1[i * 15 + e], a.params[a.emb_off + s * 8 + e], v);
        } else if (i < 16 + 112) {
            const int j = (i - 16) / 7, b = (i - 16) % 7;
            v = w1[j * 15 + 8 + b];
        } else if (i < 16 + 112 + 128) {
            v = a.params[a.w2_off[s] + (i - 128)];
        } else {
            v = a.params[a.b2_off[s] + (i - 256)];
        }
        dst[i] = v;
    }
}

__global__ void __launch_bounds__(SCE_TPB) sce_fwd_kernel(const SceArgs a) {
    __shared__ float s_p[MAXS * SCE_SM];
    // rows are grouped by scale in ascending order (frame.py), so a block only needs the scales of its own row range
    const int64_t row0 = blockIdx.x * (int64_t)SCE_TPB, rowl = min(row0 + SCE_TPB, a.n_rows) - 1;
    const int s_lo = a.scale_fixed >= 0 ? a.scale_fixed : a.scale[row0];
    const int s_hi = a.scale_fixed >= 0 ? a.scale_fixed : a.scale[rowl];
    for (int s = s_lo; s <= s_hi; ++s) sce_stage_scale(a, s, s_p + (s - s_lo) * SCE_SM);
    __syncthreads();
    const int64_t row = row0 + threadIdx.x;
    if (row >= a.n_rows) return;
    const float *p = a.scale_fixed >= 0 ? s_p : s_p + (a.scale[row] - s_lo) * SCE_SM;
    const unsigned bits = a.nbr7[row];
    float out[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) out[c] = p[256 + c];
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
        float h = p[j];
#pragma unroll
        for (int b = 0; b < 7; ++b)
            if ((bits >> b) & 1u) h += p[16 + j * 7 + b];
        h = fmaxf(h, 0.f);
#pragma unroll
        for (int c = 0; c < 8; ++c) out[c] = fmaf(p[128 + c * 16 + j], h, out[c]);
    }
    store_row<8>(tptr(a.f0, 0, row), out);
}

// Backward: warp w < 8 owns hidden units 2w, 2w+1, LANE = ROW (coalesced nbr7 / df0 loads, lane-private sums, one
// transposing butterfly per (chunk, scale)); warp 8 sums df0 for the second-layer bias.  Rows are grouped by scale,
// so a chunk touches few scales; the records of the others are zero-filled.
// Layout of one SCE partial record per scale (floats): [dW2 8*16][db2 8][dW1n 16*7][db1eff 16] = 264
constexpr int SCE_REC = 128 + 8 + 112 + 16;
constexpr int SCE_BWD_TPB = 288;
__global__ void __launch_bounds__(SCE_BWD_TPB) sce_bwd_kernel(const SceArgs a, float *rec /* [n_chunks][scale_num][SCE_REC] */) {
    __shared__ float s_p[MAXS * SCE_SM];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t r0 = blockIdx.x * a.chunk, r1 = min(r0 + a.chunk, a.n_rows);
    float *out = rec + (int64_t)blockIdx.x * a.scale_num * SCE_REC;
    int s_lo = a.scale_num, s_hi = -1;
    if (r0 < r1) {
        s_lo = a.scale_fixed >= 0 ? a.scale_fixed : a.scale[r0];
        s_hi = a.scale_fixed >= 0 ? a.scale_fixed : a.scale[r1 - 1];
    }
    for (int s = s_lo; s <= s_hi; ++s) sce_stage_scale(a, s, s_p + s * SCE_SM);
    __syncthreads();
    for (int s = 0; s < a.scale_num; ++s) {
        if (s < s_lo || s > s_hi) {
            for (int i = threadIdx.x; i < SCE_REC; i += blockDim.x) out[s * SCE_REC + i] = 0.f;
            continue;
        }
        const float *p = s_p + s * SCE_SM;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
        if (w < 8) {
            // per hidden unit jj: v[16jj + c] = dW2[c][j], v[16jj + 8 + b] = dW1n[j][b], v[16jj + 15] = db1eff[j]
            float b1e[2], w1n[2][7], w2c[2][8];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = 2 * w + jj;
                b1e[jj] = p[j];
#pragma unroll
                for (int b = 0; b < 7; ++b) w1n[jj][b] = p[16 + j * 7 + b];
#pragma unroll
                for (int c = 0; c < 8; ++c) w2c[jj][c] = p[128 + c * 16 + j];
            }
#pragma unroll 4
            for (int64_t r = r0 + lane; r < r1; r += 32) {
                const int rs = a.scale_fixed >= 0 ? a.scale_fixed : a.scale[r];
                if (rs != s) continue;
                const unsigned bits = a.nbr7[r];
                float d[8];
                load_row<8>(tptr(a.df0, 0, r), d);
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    float h = b1e[jj];
#pragma unroll
                    for (int b = 0; b < 7; ++b)
                        if ((bits >> b) & 1u) h += w1n[jj][b];   // same order as the forward kernel
                    h = fmaxf(h, 0.f);
                    float dh = 0.f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        v[16 * jj + c] = fmaf(d[c], h, v[16 * jj + c]);
                        dh = fmaf(w2c[jj][c], d[c], dh);
                    }
                    dh = h > 0.f ? dh : 0.f;
#pragma unroll
                    for (int b = 0; b < 7; ++b)
                        if ((bits >> b) & 1u) v[16 * jj + 8 + b] += dh;
                    v[16 * jj + 15] += dh;
                }
            }
            warp_transpose_reduce<32>(v, lane);
            const int j = 2 * w + (lane >> 4), e = lane & 15;
            float *o = out + s * SCE_REC;
            if (e < 8) o[e * 16 + j] = v[0];
            else if (e < 15) o[136 + j * 7 + (e - 8)] = v[0];
            else o[248 + j] = v[0];
        } else {
            for (int64_t r = r0 + lane; r < r1; r += 32) {
                const int rs = a.scale_fixed >= 0 ? a.scale_fixed : a.scale[r];
                if (rs != s) continue;
                float d[8];
                load_row<8>(tptr(a.df0, 0, r), d);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] += d[c];
            }
            warp_transpose_reduce<32>(v, lane);
            if (lane < 8) out[s * SCE_REC + 128 + lane] = v[0];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Head backward.  rows kernel: dc[k][row][8] = W1^T (dz * w2 * [hidden>0]);  weights kernel: (j, sub) mapping.
// ------------------------------------------------------------------------------------------------
struct HeadBwdArgs {
    int64_t n_rows;
    const float *params;
    int w1_off[MAXG], b1_off[MAXG], w2_off[MAXG], b2_off[MAXG];
    Tens c;           // saved conv outputs [8][rows][8]
    const float *dz;  // [8][rows]
    Tens dc;          // out
    float *partial;
    int64_t P, chunk;
};

constexpr int HEAD_RPT = 4;  // rows per thread: one pair of 128-bit weight reads per hidden unit serves all of them
__global__ void __launch_bounds__(128) head_bwd_rows_kernel(const HeadBwdArgs a) {
    __shared__ __align__(16) float s_head[241];
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < 192; i += 128) s_head[i] = a.params[a.w1_off[g] + i];
    if (threadIdx.x < 24) {
        s_head[192 + threadIdx.x] = a.params[a.b1_off[g] + threadIdx.x];
        s_head[216 + threadIdx.x] = a.params[a.w2_off[g] + threadIdx.x];
    }
    __syncthreads();
    int64_t row[HEAD_RPT];
    float c[HEAD_RPT][8], dc[HEAD_RPT][8], dz[HEAD_RPT];
#pragma unroll
    for (int r = 0; r < HEAD_RPT; ++r) {
        row[r] = blockIdx.x * (int64_t)(128 * HEAD_RPT) + r * 128 + threadIdx.x;
        const bool live = row[r] < a.n_rows;
#pragma unroll
        for (int i = 0; i < 8; ++i) c[r][i] = 0.f, dc[r][i] = 0.f;
        dz[r] = 0.f;
        if (live) {
            load_row<8>(tptr(a.c, g, row[r]), c[r]);
            dz[r] = a.dz[g * a.n_rows + row[r]];
        }
    }
#pragma unroll 2
    for (int j = 0; j < 24; ++j) {
        float w[8];
        load_row<8>(s_head + j * 8, w);
        const float b1 = s_head[192 + j], w2 = s_head[216 + j];
#pragma unroll
        for (int r = 0; r < HEAD_RPT; ++r) {
            float h = b1;
#pragma unroll
            for (int i = 0; i < 8; ++i) h = fmaf(w[i], c[r][i], h);
            const float dh = h > 0.f ? dz[r] * w2 : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) dc[r][i] = fmaf(w[i], dh, dc[r][i]);
        }
    }
#pragma unroll
    for (int r = 0; r < HEAD_RPT; ++r)
        if (row[r] < a.n_rows) store_row<8>(tptr(a.dc, g, row[r]), dc[r]);
}

// Weight gradients of MLP_k: warp w owns hidden units 3w .. 3w+2, LANE = ROW (coalesced loads of the saved conv
// outputs and dz, lane-private sums, transposing butterfly at the end of the chunk).  Element order of the 32
// per-warp sums: dW1[3][8], db1[3], dW2[3], db2 (warp 0 only), pad.
__global__ void __launch_bounds__(256) head_bwd_w_kernel(const HeadBwdArgs a) {
    const int g = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t r0 = blockIdx.x * a.chunk, r1 = min(r0 + a.chunk, a.n_rows);
    float w1[3][8], b1[3], w2[3];
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
        const int j = 3 * w + jj;
#pragma unroll
        for (int i = 0; i < 8; ++i) w1[jj][i] = a.params[a.w1_off[g] + j * 8 + i];
        b1[jj] = a.params[a.b1_off[g] + j], w2[jj] = a.params[a.w2_off[g] + j];
    }
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    const float *dzg = a.dz + g * a.n_rows;
#pragma unroll 4
    for (int64_t r = r0 + lane; r < r1; r += 32) {
        float c[8];
        gather_row<8>(tptr(a.c, g, r), c);
        const float dz = dzg[r];
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            float h = b1[jj];
#pragma unroll
            for (int i = 0; i < 8; ++i) h = fmaf(w1[jj][i], c[i], h);
            h = fmaxf(h, 0.f);
            const float dh = h > 0.f ? dz * w2[jj] : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * jj + i] = fmaf(dh, c[i], v[8 * jj + i]);
            v[24 + jj] += dh;
            v[27 + jj] = fmaf(dz, h, v[27 + jj]);
        }
        v[30] += dz;
    }
    warp_transpose_reduce<32>(v, lane);
    float *out = a.partial + blockIdx.x * a.P;
    if (lane < 24) out[a.w1_off[g] + (3 * w + lane / 8) * 8 + (lane & 7)] = v[0];
    else if (lane < 27) out[a.b1_off[g] + 3 * w + (lane - 24)] = v[0];
    else if (lane < 30) out[a.w2_off[g] + 3 * w + (lane - 27)] = v[0];
    else if (lane == 30 && w == 0) out[a.b2_off[g]] = v[0];
}

// dg[row][8] = sum over the 8 stages of dh_k[row][8] (h_k = g + LDFE_{k-1}: every stage feeds g), fixed order.
__global__ void sum_groups_kernel(const float *__restrict__ src, int64_t gs, int groups, int64_t n4, float *__restrict__ dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 s = reinterpret_cast<const float4 *>(src)[i];
    for (int g = 1; g < groups; ++g) {
        const float4 v = reinterpret_cast<const float4 *>(src + g * gs)[i];
        s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    reinterpret_cast<float4 *>(dst)[i] = s;
}

// ------------------------------------------------------------------------------------------------
// Final reductions, Adam, quantisation
// ------------------------------------------------------------------------------------------------
// grad[j] = sum over chunks (in chunk order) of partial[chunk][j] for the conv / MLP parameters.
__global__ void finalize_grad_kernel(const float *__restrict__ partial, int64_t P, int n_chunks, int64_t first, int64_t count,
                                     float *__restrict__ grad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t j = first + i;
    float t = 0.f;
    for (int c = 0; c < n_chunks; ++c) t += partial[c * P + j];
    grad[j] = t;
}

// SCE: reduce records over chunks, then expand the folded embedding terms (one block per scale).  A warp sums one
// record element at a time: lanes stride over the chunks, then a fixed xor tree.
__global__ void __launch_bounds__(1024) sce_finalize_kernel(const SceArgs a, const float *__restrict__ rec, int n_chunks,
                                                            float *__restrict__ grad) {
    __shared__ float s_r[SCE_REC];
    const int s = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = w; i < SCE_REC; i += nw) {
        float t = 0.f;
        for (int c0 = lane; c0 < n_chunks; c0 += 32 * 8) {   // eight independent loads in flight, summed in order
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + 32 * u;
                x[u] = c < n_chunks ? rec[((int64_t)c * a.scale_num + s) * SCE_REC + i] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) t += x[u];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) s_r[i] = t;
    }
    __syncthreads();
    const float *emb = a.params + a.emb_off + s * 8;
    const float *w1 = a.params + a.w1_off[s];
    for (int i = threadIdx.x; i < 16 * 15; i += blockDim.x) {
        const int j = i / 15, e = i % 15;
        grad[a.w1_off[s] + i] = e < 8 ? s_r[248 + j] * emb[e] : s_r[136 + j * 7 + (e - 8)];
    }
    for (int i = threadIdx.x; i < 16; i += blockDim.x) grad[a.b1_off[s] + i] = s_r[248 + i];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) grad[a.w2_off[s] + i] = s_r[i];
    for (int i = threadIdx.x; i < 8; i += blockDim.x) {
        grad[a.b2_off[s] + i] = s_r[128 + i];
        float t = 0.f;
        for (int j = 0; j < 16; ++j) t = fmaf(w1[j * 15 + i], s_r[248 + j], t);
        grad[a.emb_off + s * 8 + i] = t;
    }
}

__global__ void bits_finalize_kernel(const float *__restrict__ partial, int n, double *__restrict__ out) {
    // single thread-block, fixed order, double accumulation
    __shared__ double s[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) t += (double)partial[i];
    s[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s[0];
}

// torch.optim.Adam, weight decay added to the gradient (main.py:231-237). bc1/bc2: bias corrections.
__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                            int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i]);
    const float mi = fmaf(1.f - b1, gi, b1 * m[i]);
    const float vi = fmaf((1.f - b2) * gi, gi, b2 * v[i]);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}

// One block: min/max -> quantise -> mean -> mean abs dev -> dequantise (model_size_est.py:72-91,410-411).
template <typename QT>   // uint8_t symbols for bit depths <= 8, uint16_t for 9..16
__global__ void __launch_bounds__(1024) quant_kernel(const float *__restrict__ p, int64_t n, float smax, QT *__restrict__ q,
                                                     float *__restrict__ recon, float *__restrict__ stats) {
    __shared__ float s_a[1024], s_b[1024];
    __shared__ double s_d[1024];
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        mn = fminf(mn, p[i]);
        mx = fmaxf(mx, p[i]);
    }
    s_a[threadIdx.x] = mn, s_b[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if (threadIdx.x < o) {
            s_a[threadIdx.x] = fminf(s_a[threadIdx.x], s_a[threadIdx.x + o]);
            s_b[threadIdx.x] = fmaxf(s_b[threadIdx.x], s_b[threadIdx.x + o]);
        }
        __syncthreads();
    }
    mn = s_a[0], mx = s_b[0];
    const float rng = mx - mn;
    double sum = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        // round((w - min) / range * smax): same op order as torch (division, then multiply), half-to-even
        const float s = rintf(__fmul_rn(__fdiv_rn(__fsub_rn(p[i], mn), rng), smax));
        q[i] = (QT)s;
        recon[i] = __fadd_rn(__fmul_rn(__fdiv_rn(s, smax), rng), mn);
        sum += (double)s;
    }
    __syncthreads();
    s_d[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if (threadIdx.x < o) s_d[threadIdx.x] += s_d[threadIdx.x + o];
        __syncthreads();
    }
    const float mu = rintf((float)(s_d[0] / (double)n));
    __syncthreads();
    double dev = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) dev += fabs((double)q[i] - (double)mu);
    s_d[threadIdx.x] = dev;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if (threadIdx.x < o) s_d[threadIdx.x] += s_d[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[0] = mn, stats[1] = mx, stats[2] = mu;
        stats[3] = rintf((float)(s_d[0] / (double)n));
    }
}

__global__ void occ_set_stage_kernel(uint8_t *__restrict__ occ, const uint8_t *__restrict__ sym, int64_t n, int stage) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) occ[i] = (uint8_t)(occ[i] | ((sym[i] & 1u) << stage));
}

}  // namespace linr
