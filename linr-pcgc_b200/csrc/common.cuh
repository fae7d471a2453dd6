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
