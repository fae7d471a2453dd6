/* CPU oracle: range coder restating torchac 0.9.3's backend.  TEST INFRASTRUCTURE ONLY.
 *
 * torchac is a third-party dependency of the reference (enviroment.yaml:32) that is
 * absent from /root/reference; the reference calls it at models/module_utils.py:28,38
 * and model_compression/model_size_est.py:482,561.  This file restates its published
 * algorithm (32-bit low/high arithmetic coder, 16-bit CDF precision, E1/E2/E3
 * renormalisation with pending bits, MSB-first bit packing, zero-padded flush; the
 * decoder primes 32 bits and binary-searches the CDF row).  PARITY UNPINNED: the
 * reference holds no golden bitstreams; the only pinned property is the round trip
 * the reference itself asserts (models/upsample.py:235-237, model_core.py:217).
 *
 * CDF rows are uint16 [Lp] as produced by torchac's float->int16 conversion; the
 * upper bound of the last symbol (Lp-2) is hard-wired to 0x10000.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/librc_oracle.so oracle/rc_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint8_t *buf;
    size_t cap, len;
    uint8_t cache;
    int count;
} bitsink;

static void sink_bit(bitsink *s, int bit) {
    s->cache = (uint8_t)((s->cache << 1) | (bit & 1));
    if (++s->count == 8) {
        if (s->len < s->cap) s->buf[s->len] = s->cache;
        s->len++;
        s->count = 0;
        s->cache = 0;
    }
}
static void sink_bit_pending(bitsink *s, int bit, uint64_t *pending) {
    sink_bit(s, bit);
    while (*pending > 0) { sink_bit(s, !bit); (*pending)--; }
}

/* row_stride == 0 -> one shared CDF row for every symbol. Returns bytes needed (may exceed cap). */
long long rc_oracle_encode(const uint16_t *cdf, long long row_stride, int Lp, const int16_t *sym,
                           long long n, uint8_t *out, long long cap) {
    bitsink s = {out, (size_t)cap, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFu;
    uint64_t pending = 0;
    const int max_symbol = Lp - 2;
    for (long long i = 0; i < n; ++i) {
        const uint16_t *row = cdf + i * row_stride;
        const int si = sym[i];
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint32_t c_low = row[si];
        const uint32_t c_high = (si == max_symbol) ? 0x10000u : row[si + 1];
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (high < 0x80000000u) {
                sink_bit_pending(&s, 0, &pending);
                low <<= 1; high <<= 1; high |= 1;
            } else if (low >= 0x80000000u) {
                sink_bit_pending(&s, 1, &pending);
                low <<= 1; high <<= 1; high |= 1;
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                pending++;
                low <<= 1; low &= 0x7FFFFFFFu;
                high <<= 1; high |= 0x80000001u;
            } else break;
        }
    }
    pending += 1;
    if (low < 0x40000000u) sink_bit_pending(&s, 0, &pending);
    else sink_bit_pending(&s, 1, &pending);
    if (s.count > 0) { for (int i = s.count; i < 8; ++i) sink_bit(&s, 0); }
    return (long long)s.len;
}

typedef struct { const uint8_t *in; size_t n, pos; uint8_t cache; int cached; } bitsrc;
static void src_get(bitsrc *b, uint32_t *value) {
    if (b->cached == 0) {
        if (b->pos == b->n) { *value <<= 1; return; }
        b->cache = b->in[b->pos++];
        b->cached = 8;
    }
    *value <<= 1;
    *value |= (uint32_t)((b->cache >> (b->cached - 1)) & 1);
    b->cached--;
}

static int cdf_search(const uint16_t *row, uint16_t target, int max_sym) {
    int left = 0, right = max_sym + 1;
    while (left + 1 < right) {
        const int m = (left + right) / 2;
        const uint16_t v = row[m];
        if (v < target) left = m;
        else if (v > target) right = m;
        else return m;
    }
    return left;
}

int rc_oracle_decode(const uint16_t *cdf, long long row_stride, int Lp, const uint8_t *in, long long nbytes,
                     int16_t *sym, long long n) {
    bitsrc b = {in, (size_t)nbytes, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFu, value = 0;
    const int max_symbol = Lp - 2;
    for (int i = 0; i < 32; ++i) src_get(&b, &value);
    for (long long i = 0; i < n; ++i) {
        const uint16_t *row = cdf + i * row_stride;
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint16_t count = (uint16_t)((((uint64_t)value - (uint64_t)low + 1) * 0x10000ull - 1) / span);
        const int si = cdf_search(row, count, max_symbol);
        sym[i] = (int16_t)si;
        if (i == n - 1) break;
        const uint32_t c_low = row[si];
        const uint32_t c_high = (si == max_symbol) ? 0x10000u : row[si + 1];
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (low >= 0x80000000u || high < 0x80000000u) {
                low <<= 1; high <<= 1; high |= 1;
                src_get(&b, &value);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low <<= 1; low &= 0x7FFFFFFFu;
                high <<= 1; high |= 0x80000001u;
                value -= 0x40000000u;
                src_get(&b, &value);
            } else break;
        }
    }
    return 0;
}
