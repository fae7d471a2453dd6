// Shared helpers for the linr_b200 CUDA sources (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/linr_b200.h"

void linr_set_error(const char *fmt, ...);

#define LINR_CHECK_CUDA(expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            linr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return LINR_ECUDA;                                                                  \
        }                                                                                       \
    } while (0)

#define LINR_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            linr_set_error(__VA_ARGS__); \
            return LINR_EINVAL;          \
        }                                \
    } while (0)

#define LINR_LAUNCH_CHECK() LINR_CHECK_CUDA(cudaGetLastError())

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct WsCursor {
    char *base;
    size_t off, cap;
    bool ok;
    WsCursor(void *p, size_t bytes) : base((char *)p), off(0), cap(bytes), ok(true) {}
    template <typename T>
    T *take(size_t n) {
        off = align_up(off, 256);
        T *r = (T *)(base + off);
        off += n * sizeof(T);
        if (off > cap) ok = false;
        return r;
    }
};

int linr_sm_count();

// Weight-gradient v3 (net_kernels.cuh) and the pair lists it walks (coords.cu): the 27 offsets + the bias are dealt to
// two halves of 14 slots (27 = bias, no list); a block works on one half, warp i of the block on slot [half][i].  The
// table balances the offsets' densities (centre 1.0, faces .65, edges .51, corners .41) per SM sub-partition.  The
// pair lists of a tile are stored half by half in this order, so that a block fetches its 14 lists with ONE bulk copy.
#define LINR_BW3_SLOTS                                                                             \
    {                                                                                              \
        {1, 5, 12, 10, 3, 7, 14, 16, 0, 6, 9, 11, 2, 8}, { 15, 19, 4, 13, 17, 21, 22, 24, 18, 23, 25, 26, 20, 27 } \
    }
constexpr int BW3_NW = 14, BW3_T = 256;
constexpr int PAIR_TILE_ENTRIES = 27 * 256, PAIR_HALF_ENTRIES = BW3_NW * 256;   // list storage of a tile / where half 1 starts
