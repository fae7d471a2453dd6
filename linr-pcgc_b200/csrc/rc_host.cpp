// Host range coder of the product path (the serial coder stays on the CPU by design, BASELINE.json north_star).
// Bitstream-compatible with torchac 0.9.3 (32-bit low/high, 16-bit CDFs, pending-bit renormalisation,
// MSB-first packing, zero-padded flush) as called by the reference at models/module_utils.py:28,38 and
// model_compression/model_size_est.py:482,561.  Binary streams take the 16-bit boundary P(sym=0) per
// symbol (what linr_net_forward emits on the GPU), so no float CDF is ever built on the host; output bits
// are accumulated in a 64-bit register and independent streams are coded on a pool of host threads.
#include <atomic>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/linr_b200.h"

namespace {

struct BitWriter {
    uint8_t *out;
    int64_t cap, len = 0;
    uint64_t acc = 0;
    int nacc = 0;
    BitWriter(uint8_t *o, int64_t c) : out(o), cap(c) {}
    inline void put(unsigned bit) {
        acc = (acc << 1) | (bit & 1u);
        if (++nacc == 64) spill();
    }
    inline void spill() {
        for (int i = 56; i >= 0; i -= 8) {
            if (len < cap) out[len] = (uint8_t)(acc >> i);
            ++len;
        }
        acc = 0, nacc = 0;
    }
    inline void put_with_pending(unsigned bit, uint64_t &pending) {
        put(bit);
        while (pending) {
            put(bit ^ 1u);
            --pending;
        }
    }
    int64_t finish() {
        // whole bytes first, then the zero-padded tail
        int full = nacc / 8, rem = nacc % 8;
        for (int i = 0; i < full; ++i) {
            const int sh = nacc - 8 * (i + 1);
            if (len < cap) out[len] = (uint8_t)(acc >> sh);
            ++len;
        }
        if (rem) {
            if (len < cap) out[len] = (uint8_t)((acc & ((1u << rem) - 1u)) << (8 - rem));
            ++len;
        }
        return len;
    }
};

struct Coder {
    uint32_t low = 0, high = 0xFFFFFFFFu;
    uint64_t pending = 0;
    inline void narrow(uint32_t c_low, uint32_t c_high, BitWriter &w) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (high < 0x80000000u) {
                w.put_with_pending(0, pending);
            } else if (low >= 0x80000000u) {
                w.put_with_pending(1, pending);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                ++pending;
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                continue;
            } else {
                break;
            }
            low <<= 1;
            high = (high << 1) | 1u;
        }
    }
    // Binary alphabet {[0,m), [m,2^16)}: one multiply; the renormalisation loop is entered only when a bit can leave
    // (same interval arithmetic as narrow(): sym 1 -> low += t, sym 0 -> high = low + t - 1, t = span*m >> 16).
    inline void narrow_binary(uint32_t m, unsigned sym, BitWriter &w) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint32_t t = (uint32_t)((span * (uint64_t)m) >> 16);
        const uint32_t bound = low + t;
        const uint32_t sel = 0u - (uint32_t)(sym & 1u);      // branch-free select
        low = (bound & sel) | (low & ~sel);
        high = (high & sel) | ((bound - 1u) & ~sel);
        if ((((low ^ high) & 0x80000000u) != 0u) && !(low >= 0x40000000u && high < 0xC0000000u)) return;
        for (;;) {
            if (high < 0x80000000u) {
                w.put_with_pending(0, pending);
            } else if (low >= 0x80000000u) {
                w.put_with_pending(1, pending);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                ++pending;
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                continue;
            } else {
                break;
            }
            low <<= 1;
            high = (high << 1) | 1u;
        }
    }
    inline void flush(BitWriter &w) {
        pending += 1;
        w.put_with_pending(low < 0x40000000u ? 0u : 1u, pending);
    }
};

struct BitReader {
    const uint8_t *in;
    int64_t n, pos = 0;
    uint32_t cache = 0;
    int cached = 0;
    BitReader(const uint8_t *p, int64_t nb) : in(p), n(nb) {}
    inline void pull(uint32_t &value) {
        if (cached == 0) {
            if (pos == n) {
                value <<= 1;
                return;
            }
            cache = in[pos++];
            cached = 8;
        }
        value = (value << 1) | ((cache >> (cached - 1)) & 1u);
        --cached;
    }
};

// MSB-first reader with a 64-bit window: take(n) returns the next n (1..32) bits, zeros past the end (torchac pads
// the tail with zero bits).
struct BitWindow {
    const uint8_t *in;
    int64_t n, pos = 0;
    uint64_t buf = 0;
    int have = 0;
    BitWindow(const uint8_t *p, int64_t nb) : in(p), n(nb) {}
    inline void refill() {
        while (have <= 56) {
            const uint64_t byte = pos < n ? in[pos] : 0u;
            ++pos;
            buf |= byte << (56 - have);
            have += 8;
        }
    }
    inline uint32_t take(int k) {
        if (have < k) refill();
        const uint32_t v = (uint32_t)(buf >> (64 - k));
        buf <<= k;
        have -= k;
        return v;
    }
};

struct Decoder {
    uint32_t low = 0, high = 0xFFFFFFFFu, value = 0;
    void prime(BitReader &r) {
        for (int i = 0; i < 32; ++i) r.pull(value);
    }
    inline void narrow(uint32_t c_low, uint32_t c_high, uint64_t span, BitReader &r) {
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (low >= 0x80000000u || high < 0x80000000u) {
                low <<= 1;
                high = (high << 1) | 1u;
                r.pull(value);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                value -= 0x40000000u;
                r.pull(value);
            } else {
                break;
            }
        }
    }
};

int64_t encode_binary(const uint16_t *mid, const uint8_t *sym, int shift, int64_t n, uint8_t *out, int64_t cap) {
    BitWriter w(out, cap);
    Coder c;
    for (int64_t i = 0; i < n; ++i) {
        c.narrow_binary(mid[i], (sym[i] >> shift) & 1u, w);
    }
    c.flush(w);
    return w.finish();
}

}  // namespace

extern "C" {

int64_t linr_rc_encode_binary(const uint16_t *h_cdf_mid, const uint8_t *h_sym, int64_t n, uint8_t *h_out, int64_t cap) {
    const int64_t need = encode_binary(h_cdf_mid, h_sym, 0, n, h_out, cap);
    return need <= cap ? need : -need;
}

int linr_rc_decode_binary(const uint16_t *h_cdf_mid, const uint8_t *h_in, int64_t nbytes, uint8_t *h_sym, int64_t n) {
    BitWindow r(h_in, nbytes);
    uint32_t low = 0, high = 0xFFFFFFFFu, value = r.take(32);
    for (int64_t i = 0; i < n; ++i) {
        // torchac: count = ((value-low+1)*2^16 - 1) / span; sym = (count >= mid)  <=>  (value-low+1)*2^16 > mid*span
        // <=> value - low >= t with t = mid*span >> 16, which is also the interval split -> one multiply per symbol
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint32_t t = (uint32_t)((span * (uint64_t)h_cdf_mid[i]) >> 16);
        const uint32_t bound = low + t;
        const uint32_t s = value >= bound ? 1u : 0u;      // value, bound in [low, high]: no wrap
        h_sym[i] = (uint8_t)s;
        if (i == n - 1) break;
        const uint32_t sel = 0u - s;                       // branch-free select: the symbol is not predictable
        low = (bound & sel) | (low & ~sel);
        high = (high & sel) | ((bound - 1u) & ~sel);
        // renormalise: all leading bits low and high share leave at once (the bit-by-bit loop would take them one after
        // the other before it ever looks at the underflow case), then the underflow (E3) steps, and again
        for (;;) {
            const uint32_t diff = low ^ high;
            if ((diff & 0x80000000u) == 0u) {
                const int k = diff ? __builtin_clz(diff) : 32;
                if (k == 32) {
                    low = 0u, high = 0xFFFFFFFFu, value = r.take(32);
                } else {
                    low <<= k;
                    high = (high << k) | ((1u << k) - 1u);
                    value = (value << k) | r.take(k);
                }
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                value = ((value - 0x40000000u) << 1) | r.take(1);
            } else {
                break;
            }
        }
    }
    return LINR_OK;
}

int linr_rc_encode_binary_batch(int n_streams, const uint16_t *const *h_cdf_mid, const uint8_t *const *h_sym,
                                const int *h_shift, const int64_t *n, uint8_t *const *h_out, const int64_t *cap,
                                int64_t *h_written, int threads) {
    if (n_streams <= 0) return LINR_OK;
    std::vector<int> order(n_streams);
    for (int i = 0; i < n_streams; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return n[a] > n[b]; });  // longest first
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int j = next.fetch_add(1);
            if (j >= n_streams) return;
            const int i = order[j];
            const int64_t need = encode_binary(h_cdf_mid[i], h_sym[i], h_shift ? (h_shift[i] & 7) : 0, n[i], h_out[i], cap[i]);
            h_written[i] = need <= cap[i] ? need : -need;
        }
    };
    int nt = threads < 1 ? 1 : threads;
    if (nt > n_streams) nt = n_streams;
    if (nt == 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        pool.reserve(nt - 1);
        for (int t = 0; t < nt - 1; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    }
    for (int i = 0; i < n_streams; ++i)
        if (h_written[i] < 0) return LINR_ENOMEM;
    return LINR_OK;
}

int linr_rc_decode_binary_batch(int n_streams, const uint16_t *const *h_cdf_mid, const uint8_t *const *h_in, const int64_t *nbytes,
                                uint8_t *const *h_sym, const int64_t *n, int threads) {
    if (n_streams <= 0) return LINR_OK;
    std::vector<int> order(n_streams);
    for (int i = 0; i < n_streams; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return n[a] > n[b]; });  // longest first
    std::atomic<int> next(0);
    std::atomic<int> bad(0);
    auto work = [&]() {
        for (;;) {
            const int j = next.fetch_add(1);
            if (j >= n_streams) return;
            const int i = order[j];
            if (n[i] > 0 && linr_rc_decode_binary(h_cdf_mid[i], h_in[i], nbytes[i], h_sym[i], n[i]) != LINR_OK) bad.store(1);
        }
    };
    int nt = threads < 1 ? 1 : threads;
    if (nt > n_streams) nt = n_streams;
    if (nt == 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        pool.reserve(nt - 1);
        for (int t = 0; t < nt - 1; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    }
    return bad.load() ? LINR_EINVAL : LINR_OK;
}

int64_t linr_rc_encode_shared(const uint16_t *h_cdf_row, int Lp, const int16_t *h_sym, int64_t n, uint8_t *h_out, int64_t cap) {
    if (Lp < 3) return 0;
    BitWriter w(h_out, cap);
    Coder c;
    const int max_symbol = Lp - 2;
    for (int64_t i = 0; i < n; ++i) {
        const int s = h_sym[i];
        if (s < 0 || s > max_symbol) return 0;
        c.narrow(h_cdf_row[s], s == max_symbol ? 0x10000u : (uint32_t)h_cdf_row[s + 1], w);
    }
    c.flush(w);
    const int64_t need = w.finish();
    return need <= cap ? need : -need;
}

int linr_rc_decode_shared(const uint16_t *h_cdf_row, int Lp, const uint8_t *h_in, int64_t nbytes, int16_t *h_sym, int64_t n) {
    if (Lp < 3) return LINR_EINVAL;
    BitReader r(h_in, nbytes);
    Decoder d;
    d.prime(r);
    const int max_symbol = Lp - 2;
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t span = (uint64_t)d.high - (uint64_t)d.low + 1;
        const uint16_t count = (uint16_t)(((((uint64_t)d.value - (uint64_t)d.low + 1) << 16) - 1) / span);
        int left = 0, right = max_symbol + 1;
        while (left + 1 < right) {
            const int m = (left + right) / 2;
            const uint16_t v = h_cdf_row[m];
            if (v < count) left = m;
            else if (v > count) right = m;
            else {
                left = m;
                break;
            }
        }
        h_sym[i] = (int16_t)left;
        if (i == n - 1) break;
        d.narrow(h_cdf_row[left], left == max_symbol ? 0x10000u : (uint32_t)h_cdf_row[left + 1], span, r);
    }
    return LINR_OK;
}

}  // extern "C"
