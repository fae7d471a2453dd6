// Host-side orchestration of the occupancy network: parameter layout, workspace carving and the
// stream-ordered launch sequences for forward / backward / sequential decode.  All launches go to the
// caller's stream; nothing here synchronises.  Mirrors (does not copy) the structure of
// models/model_core.py:38-81, models/upsample.py:137-295 and models/resnet.py:55-60 of the reference.
#include <math.h>
#include <stdlib.h>
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "net_kernels.cuh"
#include "prof.cuh"

using namespace linr;

// ---------------------------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
void linr_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
// SM count of the CURRENT device (cached per device: ranks of one process may drive different GPUs)
int linr_sm_count() {
    static std::atomic<int> sm[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = sm[dev].load(std::memory_order_relaxed);
    if (!v) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sm[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------- profiler
namespace linr {
namespace {
struct ProfState {
    std::atomic<uint32_t> mask{0};
    std::atomic<int64_t> launches[K_NCLASS];
    std::atomic<int64_t> units[K_NCLASS];
    std::vector<cudaEvent_t> ev[K_NCLASS];  // begin/end pairs recorded so far
    std::vector<cudaEvent_t> pool;          // recycled events
    double ms_done[K_NCLASS] = {0};
    std::mutex mu;  // launches come from several host threads (GOP coder, concurrent decoders)
} g_prof;
const char *const kNames[K_NCLASS] = {"conv27<8,8>", "conv27<8,4>", "conv27<4,8>", "conv27<4,4>", "conv27_bits<8>", "conv27_head",
                                      "bwd_w<8,8>", "bwd_w<8,4>", "bwd_w<4,4>", "bwd_w_bits<8>", "pointwise", "pointwise_bwd_w",
                                      "head_bwd", "sce", "reduce", "adam_quant", "coord"};
cudaEvent_t prof_event() {
    if (!g_prof.pool.empty()) {
        cudaEvent_t e = g_prof.pool.back();
        g_prof.pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
// Fold the recorded pairs of one class into ms_done (synchronises on them) and recycle the events.
void prof_drain(int c) {
    auto &v = g_prof.ev[c];
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
        float ms = 0.f;
        if (cudaEventSynchronize(v[i + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, v[i], v[i + 1]) == cudaSuccess) g_prof.ms_done[c] += ms;
        g_prof.pool.push_back(v[i]);
        g_prof.pool.push_back(v[i + 1]);
    }
    v.clear();
}
}  // namespace
// Launch counters are relaxed atomics: with timing off (mask 0, the normal state) a launch takes no lock.
void prof_begin(int cls, int64_t units, cudaStream_t s) {
    g_prof.launches[cls].fetch_add(1, std::memory_order_relaxed);
    g_prof.units[cls].fetch_add(units, std::memory_order_relaxed);
    if (g_prof.mask.load(std::memory_order_relaxed) >> cls & 1u) {
        std::lock_guard<std::mutex> lock(g_prof.mu);
        cudaEvent_t e = prof_event();
        cudaEventRecord(e, s);
        g_prof.ev[cls].push_back(e);
    }
}
void prof_end(int cls, cudaStream_t s) {
    if (g_prof.mask.load(std::memory_order_relaxed) >> cls & 1u) {
        std::lock_guard<std::mutex> lock(g_prof.mu);
        if (g_prof.ev[cls].size() % 2 == 0) return;   // timing was switched on between begin and end: no open pair
        cudaEvent_t e = prof_event();
        cudaEventRecord(e, s);
        g_prof.ev[cls].push_back(e);
        if (g_prof.ev[cls].size() >= 16384) prof_drain(cls);  // bound the number of live events
    }
}
}  // namespace linr

// ---------------------------------------------------------------------------------------------- layout
namespace {

struct BlockL {
    int A_w, A_b, c00_w, c00_b, c01_w, c01_b, c10_w, c10_b, c11_w, c11_b, c12_w, c12_b, B_w, B_b;
};
struct Layout {
    int S = 0;
    int emb = 0;
    int sce_w1[MAXS], sce_b1[MAXS], sce_w2[MAXS], sce_b2[MAXS];
    BlockL bin;
    int mlp_w1[8], mlp_b1[8], mlp_w2[8], mlp_b2[8];
    int pr_w[8], pr_b[8];
    BlockL ob[7];
    BlockL blk8[8];  // {ob[0..6], bin}: the five inner layers of all eight blocks run as one 8-group launch
    int conv_first = 0;  // first parameter that is not SCE (embedding + scale MLPs)
    int total = 0;
    std::vector<int64_t> offsets;  // one per tensor, parameters() order
    // weight bank plan (net_kernels.cuh): fills 0..3 serve the forward launches, 4..7 the grad-input launches
    std::vector<BankItem> bank;
    int fill_len[8] = {0};
    int fill_base[8] = {0};
    int stage_floats = 0;
};

// parameters() order of LINR_PCGC_Model (checkpoint contract)
Layout make_layout(int S) {
    Layout L;
    L.S = S;
    int o = 0;
    auto take = [&](int n) {
        L.offsets.push_back(o);
        int r = o;
        o += n;
        return r;
    };
    L.emb = take(S * 8);
    for (int s = 0; s < S; ++s) {
        L.sce_w1[s] = take(16 * 15);
        L.sce_b1[s] = take(16);
        L.sce_w2[s] = take(8 * 16);
        L.sce_b2[s] = take(8);
    }
    L.conv_first = o;
    auto block = [&](int cin) {
        BlockL b;
        b.A_w = take(27 * cin * 8), b.A_b = take(8);
        b.c00_w = take(27 * 8 * 4), b.c00_b = take(4);
        b.c01_w = take(27 * 4 * 4), b.c01_b = take(4);
        b.c10_w = take(8 * 4), b.c10_b = take(4);
        b.c11_w = take(27 * 4 * 4), b.c11_b = take(4);
        b.c12_w = take(4 * 4), b.c12_b = take(4);
        b.B_w = take(27 * 8 * 8), b.B_b = take(8);
        return b;
    };
    L.bin = block(8);
    for (int k = 0; k < 8; ++k) {
        L.mlp_w1[k] = take(24 * 8), L.mlp_b1[k] = take(24), L.mlp_w2[k] = take(24), L.mlp_b2[k] = take(1);
    }
    for (int k = 0; k < 8; ++k) L.pr_w[k] = take(27 * 8 * 8), L.pr_b[k] = take(8);
    for (int k = 0; k < 7; ++k) L.ob[k] = block(k + 1);
    for (int k = 0; k < 7; ++k) L.blk8[k] = L.ob[k];
    L.blk8[7] = L.bin;
    L.total = o;
    // ---- weight bank plan: every fill <= BANK_FLOATS, items in launch order
    auto add = [&](int fill, int w_off, int flip, int cin, int cout) {
        L.bank.push_back(BankItem{w_off, flip, cin, cout, fill, L.fill_len[fill]});
        L.fill_len[fill] += 27 * cin * cout;
    };
    add(0, L.bin.A_w, 0, 8, 8);   // forward: ConvA of GDFE (the bit-input ConvA of the LDFE blocks is faster from shared memory)
    for (int g = 0; g < 8; ++g) add(1, L.blk8[g].c00_w, 0, 8, 4);        // the three 27-offset inner layers
    for (int g = 0; g < 8; ++g) add(1, L.blk8[g].c01_w, 0, 4, 4);
    for (int g = 0; g < 8; ++g) add(1, L.blk8[g].c11_w, 0, 4, 4);
    add(2, L.bin.B_w, 0, 8, 8);                                        // ConvB
    for (int g = 0; g < 7; ++g) add(2, L.ob[g].B_w, 0, 8, 8);
    for (int k = 0; k < 8; ++k) add(3, L.pr_w[k], 0, 8, 8);            // SConv_k of the heads
    for (int k = 0; k < 8; ++k) add(4, L.pr_w[k], 1, 8, 8);            // backward: grad-input layouts
    for (int g = 0; g < 8; ++g) add(5, L.blk8[g].B_w, 1, 8, 8);
    for (int g = 0; g < 8; ++g) add(6, L.blk8[g].c01_w, 1, 4, 4);
    for (int g = 0; g < 8; ++g) add(6, L.blk8[g].c00_w, 1, 4, 8);        // launch<4,8>: dt0 (4) -> dy (8)
    for (int g = 0; g < 8; ++g) add(6, L.blk8[g].c11_w, 1, 4, 4);
    add(7, L.bin.A_w, 1, 8, 8);
    for (int f = 0; f < 8; ++f) {
        L.fill_base[f] = L.stage_floats;
        L.stage_floats += (L.fill_len[f] + 63) / 64 * 64;
    }
    return L;
}

const Layout &layout_for(int S) {
    static thread_local Layout cache[MAXS + 1];
    if (cache[S].S != S) cache[S] = make_layout(S);
    return cache[S];
}

// ---------------------------------------------------------------------------------------------- workspace
struct NetWs {
    int64_t R = 0;
    int n_chunks = 0;
    int64_t chunk = 0;
    // forward
    float *f0, *bi_y, *bi_t1, *bi_t0, *bi_t2, *bi_z;
    float *hh;  // [8][R][8]: hh[0] = g, hh[k] = g + LDFE_{k-1}
    float *ob_y, *ob_t1, *ob_t0, *ob_t2, *ob_z;  // [7][R][C]
    float *hc;                                   // [8][R][8]
    float *dzs;                                  // [8][R]
    float *bits_partial;
    // backward
    float *dc, *dhh, *dg, *g_dz, *g_dt0, *g_dy, *g_dt2, *g_dt1, *df0;
    float *partial, *sce_rec;
    float *stage = nullptr;  // weight-bank staging (train only)
    bool ok = false;
    size_t used = 0;
};

int chunks_for(int64_t R) {
    int64_t nb = ceil_div64(R, 64);
    int64_t cap = (int64_t)linr_sm_count();  // weight-gradient partial sums: one per row chunk
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return (int)nb;
}

// Row chunks of the weight-gradient partial sums: at most one per SM, whole 256-row tiles each
void set_chunks(NetWs &w, int64_t R) {
    const int cap = chunks_for(R);
    w.chunk = (ceil_div64(R > 0 ? R : 1, cap) + BW_T - 1) / BW_T * BW_T;
    w.n_chunks = (int)ceil_div64(R > 0 ? R : 1, w.chunk);
}

NetWs carve_net(void *ws, size_t bytes, int64_t R, int train, int P, int S) {
    NetWs w;
    w.R = R;
    set_chunks(w, R);
    WsCursor c(ws, bytes);
    const size_t r = (size_t)(R > 0 ? R : 1);
    w.f0 = c.take<float>(r * 8);
    w.hh = c.take<float>(r * 64);
    // block activations: groups 0..6 = LDFE blocks, group 7 = GDFE block (block_in)
    w.ob_y = c.take<float>(r * 64), w.ob_t1 = c.take<float>(r * 32), w.ob_t0 = c.take<float>(r * 32);
    w.ob_t2 = c.take<float>(r * 32), w.ob_z = c.take<float>(r * 64);
    w.bi_y = w.ob_y + 7 * R * 8, w.bi_t1 = w.ob_t1 + 7 * R * 4, w.bi_t0 = w.ob_t0 + 7 * R * 4;
    w.bi_t2 = w.ob_t2 + 7 * R * 4, w.bi_z = w.ob_z + 7 * R * 8;
    w.hc = c.take<float>(r * 64);
    w.dzs = c.take<float>(r * 8);
    w.bits_partial = c.take<float>((size_t)ceil_div64(r, 256) * 8);   // one per head block of any kernel variant (>= 256 rows each)
    if (train) {
        w.dc = c.take<float>(r * 64), w.dhh = c.take<float>(r * 72);
        w.dg = w.dhh + 8 * R * 8;  // group 8 of dhh: gradient of the GDFE output, next to the LDFE output gradients
        w.g_dz = c.take<float>(r * 64), w.g_dt0 = c.take<float>(r * 32), w.g_dy = c.take<float>(r * 64);
        w.g_dt2 = c.take<float>(r * 32), w.g_dt1 = c.take<float>(r * 32);
        w.df0 = c.take<float>(r * 8);
        w.partial = c.take<float>((size_t)w.n_chunks * P);
        w.sce_rec = c.take<float>((size_t)w.n_chunks * S * SCE_REC);
        w.stage = c.take<float>(8 * (size_t)BANK_FLOATS);
    }
    w.ok = c.ok;
    w.used = c.off;
    return w;
}

// ---------------------------------------------------------------------------------------------- launch helpers
inline Tens T(float *p, int64_t gs, int ld, int off = 0) { return Tens{p, gs, ld, off}; }
inline Tens TN() { return Tens{nullptr, 0, 0, 0}; }

// LINR_NO_STAGING=1: every kernel gathers through L1 as if no tile ranges were given (A/B runs, parity tests)
bool staging_enabled() {
    static const bool off = getenv("LINR_NO_STAGING") != nullptr;
    return !off;
}

RowMap map_of(const linr_rows *r) {
    const bool st = staging_enabled();
    return RowMap{r->d_anchor, r->ld, r->d_mask, r->n_rows, st ? r->d_tile_rng : nullptr, st ? r->d_pair_cnt : nullptr,
                  st ? r->d_pair_list : nullptr};
}

// ---- contexts and the weight bank.  The constant bank exists once per device, so ONE context per device holds it at a
// time, for the duration of a training call (linr_net_forward(train) / linr_net_backward): the call claims the bank,
// stages its fills, launches, and at its end records an event on its stream and lets go.  The next claimant -- the same
// or another context, on the same or another stream -- first makes its stream wait for that event, so a fill never
// overwrites weights that launches in flight still read.  A call that finds the bank busy (another host thread is in
// the middle of a training call) runs the shared-memory kernel variant instead, as every coding forward (train = 0)
// does: it executes the same FMAs in the same order, so the outputs are bit-identical.  Calls made without a current
// context use the device's default context.
}  // namespace
struct linr_ctx {
    int device = 0;
    bool is_default = false;
    cudaEvent_t done = nullptr;               // end of this context's last training call (created on first use)
    std::atomic<int64_t> bank_launches{0};    // constant-bank (CW) conv launches made through this context
    std::atomic<int64_t> bank_calls{0};       // training calls that held the bank
    int same_params = 0;                      // one-shot hint: the next training call sees the parameters of the previous one
    bool staged[2] = {false, false};          // forward / backward weight layouts of those parameters are in the staging area
    // second stream of a backward call: the weight-gradient launches (leaves of the dependency graph) run beside the
    // grad-input chain.  Made on first use; fork events are reused round robin (SideLane below).
    cudaStream_t side = nullptr;
    cudaEvent_t fork_ev[32] = {nullptr};
    cudaEvent_t join_ev = nullptr;
    int fork_next = 0;
};
namespace {
struct BankState {
    std::mutex mu;
    linr_ctx *holder = nullptr;   // context inside a training call right now
    linr_ctx *last = nullptr;     // context whose `done` event guards the bank's current contents
} g_bank[64];
linr_ctx g_default_ctx[64];
thread_local linr_ctx *t_ctx = nullptr;

linr_ctx *current_ctx() {
    if (t_ctx) return t_ctx;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    g_default_ctx[dev].device = dev, g_default_ctx[dev].is_default = true;
    return &g_default_ctx[dev];
}

bool bank_claim(cudaStream_t s) {
    static const bool disabled = getenv("LINR_NO_WEIGHT_BANK") != nullptr;
    if (disabled) return false;
    linr_ctx *me = current_ctx();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || dev != me->device) return false;
    BankState &b = g_bank[dev];
    std::lock_guard<std::mutex> lock(b.mu);
    if (b.holder != nullptr) return false;                                   // busy: another host thread is mid-call
    if (!me->done && cudaEventCreateWithFlags(&me->done, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        me->done = nullptr;
        return false;
    }
    if (b.last && b.last->done) cudaStreamWaitEvent(s, b.last->done, 0);     // the previous user's launches retire first
    b.holder = me;
    me->bank_calls.fetch_add(1, std::memory_order_relaxed);
    return true;
}
void bank_release(cudaStream_t s) {
    linr_ctx *me = current_ctx();
    BankState &b = g_bank[me->device];
    std::lock_guard<std::mutex> lock(b.mu);
    if (b.holder != me) return;
    cudaEventRecord(me->done, s);
    b.last = me;
    b.holder = nullptr;
}

struct BankCtx {   // lives for one linr_net_forward(train) / linr_net_backward call
    const Layout *L;
    float *stage;      // the call's conv weights in the layouts its launches read (bank_stage_kernel), in global memory
    int cur_fill;
    cudaStream_t stream;
    bool have_bank;    // this call holds the constant bank (else its lane = row launches read weights from shared memory)
};
thread_local BankCtx *t_bank = nullptr;

// Take the bank for this call and stage the items of fills [f0, f1) of its parameters in the layouts the launches read;
// launch_conv copies a fill into the bank right before the first launch that needs it.  Returns false (the call's
// launches then read their weights from shared memory) when the bank is busy.
bool bank_begin(BankCtx &ctx, const Layout &L, const float *params, float *stage, int f0, int f1, cudaStream_t s) {
    linr_ctx *me = current_ctx();
    const int dir = f0 >= 4 ? 1 : 0;
    // the caller vouches: same parameters, same workspace as its previous call -- and that call did stage them
    const bool staged_already = me->same_params != 0 && me->staged[dir];
    me->same_params = 0;
    if (!staged_already) me->staged[dir] = false;
    if (!stage || !bank_claim(s)) return false;
    ctx.L = &L, ctx.stage = stage, ctx.cur_fill = -1, ctx.stream = s, ctx.have_bank = true;
    t_bank = &ctx;
    if (staged_already) return true;
    me->staged[dir] = true;
    BankItems items;
    items.n = 0;
    for (int f = 0; f < 8; ++f) items.fill_base[f] = L.fill_base[f];
    for (int f = 8; f < 16; ++f) items.fill_base[f] = 0;
    for (const BankItem &it : L.bank)
        if (it.fill >= f0 && it.fill < f1 && items.n < BANK_MAX_ITEMS) items.it[items.n++] = it;
    {
        ProfScope prof(K_REDUCE, items.n, s);
        bank_stage_kernel<<<items.n, 256, 0, s>>>(params, items, stage);
    }
    return true;
}
struct BankScope {
    ~BankScope() {
        if (t_bank && t_bank->have_bank) bank_release(t_bank->stream);
        t_bank = nullptr;
    }
};

// The weight-gradient launches of a backward call produce chunk partials that nothing reads before the final reduction:
// they run on the context's second stream, each behind an event recorded on the caller's stream after the kernel that
// made its inputs, and the caller's stream waits for them once, before the call returns (or before the reduction).  They
// read no constant-bank weights, so the bank's fills stay ordered on the caller's stream.  Only explicit contexts
// (linr_ctx_create) have a second stream; LINR_NO_SIDE_STREAM=1 keeps everything on the caller's stream.
std::atomic<int> g_side_on{1};
struct SideLane {
    cudaStream_t main = nullptr, side = nullptr;
    linr_ctx *ctx = nullptr;
    bool used = false;
    explicit SideLane(cudaStream_t s) : main(s) {
        static const bool off = getenv("LINR_NO_SIDE_STREAM") != nullptr;
        linr_ctx *c = current_ctx();
        if (off || c->is_default || !g_side_on.load(std::memory_order_relaxed)) return;
        int dev = -1;
        if (cudaGetDevice(&dev) != cudaSuccess || dev != c->device) return;   // the context's streams live on its device
        if (!c->side) {
            if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess) {
                cudaGetLastError();
                c->side = nullptr;
                return;
            }
            bool ok = cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming) == cudaSuccess;
            for (auto &e : c->fork_ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
            if (!ok) {
                cudaGetLastError();
                return;   // half-made: side stays unused (destroy frees what exists)
            }
        }
        if (!c->join_ev || !c->fork_ev[31]) return;
        ctx = c, side = c->side;
    }
    // stream for a leaf launch whose inputs are complete on the caller's stream at this point
    cudaStream_t leaf() {
        if (!side) return main;
        cudaEvent_t e = ctx->fork_ev[ctx->fork_next];
        ctx->fork_next = (ctx->fork_next + 1) & 31;
        cudaEventRecord(e, main);
        cudaStreamWaitEvent(side, e, 0);
        used = true;
        return side;
    }
    void join() {
        if (!used) return;
        cudaEventRecord(ctx->join_ev, side);
        cudaStreamWaitEvent(main, ctx->join_ev, 0);
        used = false;
    }
    ~SideLane() { join(); }
};

// dynamic shared memory above 48 KB has to be opted into once per function and device (`done`: the caller's flags
// for THIS kernel instantiation)
bool opt_in_smem(const void *kernel, int bytes, bool (&done)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    if (!done[dev]) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        done[dev] = true;
    }
    return true;
}

// returns gridDim.x of the launch (the head kernel writes one bit-count partial per block)
template <int CIN, int COUT, int MODE>
int launch_conv(const ConvArgs &a, int G, cudaStream_t s) {
    if (a.map.n_rows <= 0) return 0;
    constexpr int cls = MODE == 2 ? K_CONVHEAD : MODE == 1 ? K_CONVBITS : (CIN == 8 ? (COUT == 8 ? K_CONV88 : K_CONV84) : (COUT == 8 ? K_CONV48 : K_CONV44));
    if (MODE != 1 && t_bank && t_bank->have_bank && !a.bias_direct) {
        // every group's weights must sit in ONE fill of the plan; otherwise this launch stays on shared memory
        ConvArgs b = a;
        int fill = -1;
        bool ok = true;
        for (int g = 0; g < G && ok; ++g) {
            const BankItem *hit = nullptr;
            for (const BankItem &it : t_bank->L->bank)
                if (it.w_off == a.w_off[g] && it.flip == (a.flip ? 1 : 0)) {
                    hit = &it;
                    break;
                }
            if (!hit || (fill >= 0 && hit->fill != fill)) ok = false;
            else fill = hit->fill, b.bank_off[g] = hit->dst;
        }
        if (ok) {
            if (t_bank->cur_fill != fill) {
                cudaMemcpyToSymbolAsync(c_bank, t_bank->stage + t_bank->L->fill_base[fill], sizeof(float) * t_bank->L->fill_len[fill],
                                        0, cudaMemcpyDeviceToDevice, s);
                t_bank->cur_fill = fill;
            }
            dim3 grid((unsigned)ceil_div64(a.map.n_rows, (ConvCfg<CIN, COUT, MODE, true>::ROWS)), (unsigned)G);
            ProfScope prof(cls, a.map.n_rows * G, s);
            current_ctx()->bank_launches.fetch_add(1, std::memory_order_relaxed);
            conv27_kernel<CIN, COUT, MODE, true><<<grid, CONV_TPB, 0, s>>>(b);
            return (int)grid.x;
        }
    }
    dim3 grid((unsigned)ceil_div64(a.map.n_rows, (ConvCfg<CIN, COUT, MODE>::ROWS)), (unsigned)G);
    ProfScope prof(cls, a.map.n_rows * G, s);
    conv27_kernel<CIN, COUT, MODE><<<grid, CONV_TPB, 0, s>>>(a);
    return (int)grid.x;
}
ConvArgs conv_args(const RowMap &m, const float *params) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.map = m;
    a.params = params;
    for (int g = 0; g < MAXG; ++g) a.b_off[g] = -1;
    a.out_ld = m.n_rows;
    return a;
}
struct BlockBufs {  // activations of G blocks, group stride = R * C
    float *y, *t1, *t0, *t2, *z;
};

// Block(x) = ConvB(IRN(ReLU(ConvA(x))))  (models/upsample.py:88-97, models/resnet.py:55-60), in three pieces so that
// the five inner layers of the GDFE block and of the seven LDFE blocks can share one 8-group launch.
// in_bits: input are the low (cin_base + g*cin_step) occupancy bits; else float x [R,8].
void block_A_forward(const float *params, const BlockL *L, int G, const RowMap &m, bool in_bits, const uint8_t *occ, int cin_base,
                     int cin_step, Tens x, float *y, float *t1, cudaStream_t s) {
    const int64_t R = m.n_rows;
    ConvArgs a = conv_args(m, params);  // ConvA + ReLU -> y, and in its epilogue conv1_0 (k=1) + ReLU -> t1
    for (int g = 0; g < G; ++g) a.w_off[g] = L[g].A_w, a.b_off[g] = L[g].A_b;
    a.y = T(y, R * 8, 8), a.relu = 1;
    a.pw_mode = 1, a.pw_relu = 1, a.y2 = T(t1, R * 4, 4);
    for (int g = 0; g < G; ++g) a.pw_w_off[g] = L[g].c10_w, a.pw_b_off[g] = L[g].c10_b;
    if (in_bits) {
        a.occ = occ, a.cin_base = cin_base, a.cin_step = cin_step;
        launch_conv<8, 8, 1>(a, G, s);
    } else {
        a.x = x;
        launch_conv<8, 8, 0>(a, G, s);
    }
}

void block_mid_forward(const float *params, const BlockL *L, int G, const RowMap &m, const BlockBufs &b, cudaStream_t s) {
    const int64_t R = m.n_rows;
    {  // conv0_0 + ReLU -> t0
        ConvArgs a = conv_args(m, params);
        for (int g = 0; g < G; ++g) a.w_off[g] = L[g].c00_w, a.b_off[g] = L[g].c00_b;
        a.x = T(b.y, R * 8, 8), a.y = T(b.t0, R * 4, 4), a.relu = 1;
        launch_conv<8, 4, 0>(a, G, s);
    }
    {  // conv0_1 -> z[:, 0:4] = u + y[:, 0:4]
        ConvArgs a = conv_args(m, params);
        for (int g = 0; g < G; ++g) a.w_off[g] = L[g].c01_w, a.b_off[g] = L[g].c01_b;
        a.x = T(b.t0, R * 4, 4), a.y = T(b.z, R * 8, 8, 0), a.res = T(b.y, R * 8, 8, 0);
        launch_conv<4, 4, 0>(a, G, s);
    }
    {  // conv1_1 + ReLU -> t2, and in its epilogue conv1_2 (k=1) -> z[:, 4:8] = v + y[:, 4:8]
        ConvArgs a = conv_args(m, params);
        for (int g = 0; g < G; ++g) a.w_off[g] = L[g].c11_w, a.b_off[g] = L[g].c11_b;
        a.x = T(b.t1, R * 4, 4), a.y = T(b.t2, R * 4, 4), a.relu = 1;
        a.pw_mode = 1, a.pw_relu = 0, a.y2 = T(b.z, R * 8, 8, 4), a.res2 = T(b.y, R * 8, 8, 4);
        for (int g = 0; g < G; ++g) a.pw_w_off[g] = L[g].c12_w, a.pw_b_off[g] = L[g].c12_b;
        launch_conv<4, 4, 0>(a, G, s);
    }
}

// ConvB (+ residual g for the LDFE blocks: h_k = g + block, models/upsample.py:213)
void block_B_forward(const float *params, const BlockL *L, int G, const RowMap &m, float *z, Tens out, Tens res_out, cudaStream_t s) {
    const int64_t R = m.n_rows;
    ConvArgs a = conv_args(m, params);
    for (int g = 0; g < G; ++g) a.w_off[g] = L[g].B_w, a.b_off[g] = L[g].B_b;
    a.x = T(z, R * 8, 8), a.y = out, a.res = res_out;
    launch_conv<8, 8, 0>(a, G, s);
}

void block_forward(const float *params, const BlockL *L, int G, const RowMap &m, bool in_bits, const uint8_t *occ, int cin_base,
                   int cin_step, Tens x, const BlockBufs &b, Tens out, Tens res_out, cudaStream_t s) {
    block_A_forward(params, L, G, m, in_bits, occ, cin_base, cin_step, x, b.y, b.t1, s);
    block_mid_forward(params, L, G, m, b, s);
    block_B_forward(params, L, G, m, b.z, out, res_out, s);
}

struct BlockGrads {  // scratch of G blocks
    float *dz, *dt0, *dy, *dt2, *dt1;
};

template <int CIN, int COUT, int MODE>
void launch_bwd_w(const RowMap &m, const NetWs &w, int P, const int *w_off, const int *b_off, int G, Tens x, Tens dy,
                  const uint8_t *occ, int cin_base, int cin_step, cudaStream_t s) {
    BwdWArgs a;
    memset(&a, 0, sizeof(a));
    a.map = m;
    for (int g = 0; g < G; ++g) a.w_off[g] = w_off[g], a.b_off[g] = b_off[g];
    a.x = x, a.dy = dy, a.occ = occ, a.cin_base = cin_base, a.cin_step = cin_step;
    a.partial = w.partial, a.P = P, a.chunk = w.chunk;
    static const int no_x = getenv("LINR_BW3_NOX") ? atoi(getenv("LINR_BW3_NOX")) : 0;
    a.no_xstage = no_x;
    constexpr int cls = MODE == 1 ? K_BWDWBITS : (COUT == 8 ? K_BWDW88 : (CIN == 8 ? K_BWDW84 : K_BWDW44));
    ProfScope prof(cls, m.n_rows * G, s);
    // v3 (staged + pair lists) needs the tile ranges, the pair lists and plain, 16-byte aligned row arrays
    const bool plain = m.tile_rng && m.pair_cnt && m.pair_list && (dy.off & 3) == 0 && (dy.gs & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(dy.p) & 15) == 0 && (w.chunk % BW3_T) == 0 &&
                       (MODE == 1 || (x.ld == CIN && (x.off & 3) == 0 && (x.gs & 3) == 0 && (reinterpret_cast<uintptr_t>(x.p) & 15) == 0));
    auto launch3 = [&, G](auto kernel, int smem, bool (&ok)[64]) {
        if (!opt_in_smem(reinterpret_cast<const void *>(kernel), smem, ok)) return false;
        dim3 grid(2u, (unsigned)w.n_chunks, 1u);
        a.groups = G;
        kernel<<<grid, 32 * (BW3_NW + 1), smem, s>>>(a);
        return true;
    };
    // which classes run v3 (bit 0: 8->8, 1: 8->4, 2: 4->4, 3: bit inputs); measured choice in profiles/README.md
    static const int v3_mask = getenv("LINR_BW3_MASK") ? atoi(getenv("LINR_BW3_MASK")) : 11;   // 4->4 stays on the lane = row kernel
    constexpr int my_bit = MODE == 1 ? 8 : (CIN == 8 ? (COUT == 8 ? 1 : 2) : 4);
    if (plain && (v3_mask & my_bit)) {
        if (dy.ld == COUT) {
            static bool ok[64] = {false};
            if (launch3(conv27_bwd_w3_kernel<CIN, COUT, MODE, COUT>, BwdW3Cfg<CIN, COUT, MODE, COUT>::SMEM, ok)) return;
        }
        if constexpr (COUT == 4 && MODE == 0) {
            if (dy.ld == 8) {
                static bool ok[64] = {false};
                if (launch3(conv27_bwd_w3_kernel<CIN, COUT, MODE, 8>, BwdW3Cfg<CIN, COUT, MODE, 8>::SMEM, ok)) return;
            }
        }
    }
    using Cfg = BwdWCfg<CIN, COUT, MODE>;
    dim3 grid((unsigned)Cfg::GX, (unsigned)w.n_chunks, (unsigned)G);  // offset slots fastest: blocks sharing a row chunk run together
    conv27_bwd_w_kernel<CIN, COUT, MODE><<<grid, Cfg::TPB, 0, s>>>(a);
}
template <int CIN, int COUT>
void launch_pw_bwd_w(int64_t R, const NetWs &w, int P, const int *w_off, const int *b_off, int G, Tens x, Tens dy, cudaStream_t s) {
    PwBwdWArgs a;
    memset(&a, 0, sizeof(a));
    a.n_rows = R;
    for (int g = 0; g < G; ++g) a.w_off[g] = w_off[g], a.b_off[g] = b_off[g];
    a.x = x, a.dy = dy, a.partial = w.partial, a.P = P, a.chunk = w.chunk;
    dim3 grid((unsigned)w.n_chunks, (unsigned)G);
    ProfScope prof(K_PWBWDW, R * G, s);
    pw_bwd_w_kernel<CIN, COUT><<<grid, PWW_TPB, 0, s>>>(a);
}

// Backward of ConvB and of the five inner layers for G blocks.  dout: gradient wrt the block output; leaves the
// gradient wrt ConvA's (post-ReLU) output in gr.dy for block_A_backward.
void block_Bmid_backward(const float *params, const BlockL *L, int G, const RowMap &m, const NetWs &w, int P, const BlockBufs &b,
                         const BlockGrads &gr, Tens dout, cudaStream_t s, SideLane &sl) {
    const int64_t R = m.n_rows;
    int wo[MAXG], bo[MAXG];
    auto offs = [&](int BlockL::*pw, int BlockL::*pb) {
        for (int g = 0; g < G; ++g) wo[g] = L[g].*pw, bo[g] = L[g].*pb;
    };
    const Tens y = T(b.y, R * 8, 8), t0 = T(b.t0, R * 4, 4), t1 = T(b.t1, R * 4, 4), t2 = T(b.t2, R * 4, 4), z = T(b.z, R * 8, 8);
    const Tens dz = T(gr.dz, R * 8, 8), dzl = T(gr.dz, R * 8, 8, 0), dzh = T(gr.dz, R * 8, 8, 4);
    const Tens dt0 = T(gr.dt0, R * 4, 4), dt1 = T(gr.dt1, R * 4, 4), dt2 = T(gr.dt2, R * 4, 4), dy = T(gr.dy, R * 8, 8);

    // ConvB: dW, then dz = B^T dout; its epilogue also applies conv1_2^T (k=1): dt2 = [t2 > 0](dz[:, 4:8] @ W12^T)
    offs(&BlockL::B_w, &BlockL::B_b);
    launch_bwd_w<8, 8, 0>(m, w, P, wo, bo, G, z, dout, nullptr, 0, 0, sl.leaf());
    {
        ConvArgs a = conv_args(m, params);
        for (int g = 0; g < G; ++g) a.w_off[g] = L[g].B_w, a.pw_w_off[g] = L[g].c12_w, a.pw_b_off[g] = -1;
        a.flip = 1, a.x = dout, a.y = dz;
        a.pw_mode = 2, a.y2 = dt2, a.mask2 = t2;
        launch_conv<8, 8, 0>(a, G, s);
    }
    // path 1 first (its dt1 is needed by the fused epilogue of conv0_0^T): conv1_2 (k=1) weights, conv1_1, conv1_0
    offs(&BlockL::c12_w, &BlockL::c12_b);
    {
        cudaStream_t l = sl.leaf();   // one fork for the two: both inputs are complete here
        launch_pw_bwd_w<4, 4>(R, w, P, wo, bo, G, t2, dzh, l);
        offs(&BlockL::c11_w, &BlockL::c11_b);
        launch_bwd_w<4, 4, 0>(m, w, P, wo, bo, G, t1, dt2, nullptr, 0, 0, l);
    }
    {
        ConvArgs a = conv_args(m, params);
        for (int g = 0; g < G; ++g) a.w_off[g] = L[g].c11_w;
        a.flip = 1, a.x = dt2, a.y = dt1, a.rmask = t1;
        launch_conv<4, 4, 0>(a, G, s);
    }
    offs(&BlockL::c10_w, &BlockL::c10_b);
    {
        cudaStream_t l = sl.leaf();
        launch_pw_bwd_w<8, 4>(R, w, P, wo, bo, G, y, dt1, l);
        // path 0: conv0_1 then conv0_0
        offs(&BlockL::c01_w, &BlockL::c01_b);
        launch_bwd_w<4, 4, 0>(m, w, P, wo, bo, G, t0, dzl, nullptr, 0, 0, l);
    }
    {
        ConvArgs a = conv_args(m, params);
        for (int g = 0; g < G; ++g) a.w_off[g] = L[g].c01_w;
        a.flip = 1, a.x = dzl, a.y = dt0, a.rmask = t0;
        launch_conv<4, 4, 0>(a, G, s);
    }
    offs(&BlockL::c00_w, &BlockL::c00_b);
    launch_bwd_w<8, 4, 0>(m, w, P, wo, bo, G, y, dt0, nullptr, 0, 0, sl.leaf());
    {  // dy = [y > 0](dz (residual) + c00^T dt0 + dt1 @ W10^T): conv1_0^T (k=1) is applied in the epilogue
        ConvArgs a = conv_args(m, params);
        for (int g = 0; g < G; ++g) a.w_off[g] = L[g].c00_w, a.pw_w_off[g] = L[g].c10_w, a.pw_b_off[g] = -1;
        a.flip = 1, a.x = dt0, a.y = dy, a.res = dz, a.rmask = y;
        a.pw_mode = 3, a.x2 = dt1;
        launch_conv<4, 8, 0>(a, G, s);
    }
}

// Backward of ConvA for G blocks: weight gradient from (x | occupancy bits, dy); input gradient only if dx.p != null.
void block_A_backward(const float *params, const BlockL *L, int G, const RowMap &m, const NetWs &w, int P, bool in_bits,
                      const uint8_t *occ, int cin_base, int cin_step, Tens x, Tens dy, Tens dx, cudaStream_t s, SideLane &sl) {
    int wo[MAXG], bo[MAXG];
    for (int g = 0; g < G; ++g) wo[g] = L[g].A_w, bo[g] = L[g].A_b;
    if (in_bits) launch_bwd_w<8, 8, 1>(m, w, P, wo, bo, G, TN(), dy, occ, cin_base, cin_step, sl.leaf());
    else launch_bwd_w<8, 8, 0>(m, w, P, wo, bo, G, x, dy, nullptr, 0, 0, sl.leaf());
    if (dx.p) {
        ConvArgs a = conv_args(m, params);
        for (int g = 0; g < G; ++g) a.w_off[g] = L[g].A_w;
        a.flip = 1, a.x = dy, a.y = dx;
        launch_conv<8, 8, 0>(a, G, s);
    }
}

SceArgs sce_args(const float *params, const Layout &L, const linr_rows *rows) {
    SceArgs a;
    memset(&a, 0, sizeof(a));
    a.n_rows = rows->n_rows, a.scale_num = L.S, a.params = params, a.emb_off = L.emb;
    for (int s = 0; s < L.S; ++s) a.w1_off[s] = L.sce_w1[s], a.b1_off[s] = L.sce_b1[s], a.w2_off[s] = L.sce_w2[s], a.b2_off[s] = L.sce_b2[s];
    a.nbr7 = rows->d_nbr7, a.scale = rows->d_scale, a.scale_fixed = -1;
    return a;
}

int head_forward(const float *params, const Layout &L, const RowMap &m, int first_stage, int G, Tens x, float *hc,
                  const uint8_t *occ, float *probs, uint16_t *cdf, float *dz, float dz_scale, float *bits_partial,
                  int stage_out_base, cudaStream_t s) {
    ConvArgs a = conv_args(m, params);
    for (int g = 0; g < G; ++g) {
        const int k = first_stage + g;
        a.w_off[g] = L.pr_w[k], a.b_off[g] = L.pr_b[k];
        a.w1_off[g] = L.mlp_w1[k], a.b1_off[g] = L.mlp_b1[k], a.w2_off[g] = L.mlp_w2[k], a.b2_off[g] = L.mlp_b2[k];
    }
    a.x = x;
    a.y = hc ? T(hc, m.n_rows * 8, 8) : TN();
    a.occ = occ, a.stage_base = first_stage, a.stage_out_base = stage_out_base;
    a.probs = probs, a.cdf = cdf, a.dz = dz, a.dz_scale = dz_scale, a.bits_partial = bits_partial;
    return launch_conv<8, 8, 2>(a, G, s);
}

int check_rows(const linr_rows *rows, int S, bool need_occ) {
    LINR_REQUIRE(rows != nullptr, "rows is null");
    LINR_REQUIRE(S >= 1 && S <= MAXS, "scale_num %d out of range [1,%d]", S, MAXS);
    LINR_REQUIRE(rows->n_rows >= 0 && rows->n_rows < (1ll << 31) / 64, "n_rows out of range");
    LINR_REQUIRE(rows->ld >= rows->n_rows, "anchor ld < n_rows");
    LINR_REQUIRE(rows->n_rows == 0 || (rows->d_anchor && rows->d_mask && rows->d_nbr7 && rows->d_scale), "null row tables");
    LINR_REQUIRE(!need_occ || rows->n_rows == 0 || rows->d_occ, "occupancy required");
    return LINR_OK;
}

}  // namespace

namespace {

// Group ranges of a stage range [lo, hi): LDFE block j = stage - 1 serves stage j + 1, so the LDFE groups are
// [jl, jh) = [max(lo,1) - 1, hi - 1); the GDFE block is group 7 of the eight-group activation arrays and always runs.
// fn(first, count) is called for [jl, jh) + {7}, as one range when they are adjacent.
template <typename F>
void for_block_groups(int jl, int jh, F fn) {
    if (jh > jl && jh == 7) {
        fn(jl, 8 - jl);
        return;
    }
    if (jh > jl) fn(jl, jh - jl);
    fn(7, 1);
}
BlockBufs bufs_at(const NetWs &w, int first) {
    const int64_t R = w.R;
    return BlockBufs{w.ob_y + first * R * 8, w.ob_t1 + first * R * 4, w.ob_t0 + first * R * 4, w.ob_t2 + first * R * 4, w.ob_z + first * R * 8};
}

// Phases of a training iteration (stage split with ONE rank owning block_in, include/linr_b200.h):
//   forward : FWD_GDFE  SCE + block_in -> g (= hh[0])
//             FWD_PRE   ConvA + inner layers of the range's LDFE blocks (they read occupancy bits, not g)
//             FWD_POST  ConvB of those blocks (+ g), the heads of the range, the bit count
//   backward: BWD_HEADS heads, dh_k, dg = sum over the range of dh_k
//             BWD_LDFE  the range's LDFE blocks (their inputs are bits: nothing flows further)
//             BWD_GDFE  block_in + SCE from the dg found in the workspace (the caller has summed it over the ranks)
//             BWD_FINAL chunk partials -> d_grad over the parameter ranges this rank wrote (own_gdfe: block_in + SCE too)
enum { FWD_GDFE = 1, FWD_PRE = 2, FWD_POST = 4, FWD_ALL = 7, BWD_HEADS = 1, BWD_LDFE = 2, BWD_GDFE = 4, BWD_FINAL = 8, BWD_ALL = 15 };

static int net_forward_impl(const float *d_params, int scale_num, const linr_rows *rows, int lo, int hi, int phases, int train,
                            float loss_scale, float *d_probs, uint16_t *d_cdf, double *d_bits, void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_rows(rows, scale_num, true);
    if (rc) return rc;
    LINR_REQUIRE(lo >= 0 && lo < hi && hi <= 8, "stage range [%d,%d) is not inside [0,8)", lo, hi);
    LINR_REQUIRE(phases > 0 && (phases & ~FWD_ALL) == 0, "forward phase mask %d out of range", phases);
    const Layout &L = layout_for(scale_num);
    const int64_t R = rows->n_rows;
    NetWs w = carve_net(d_ws, ws_bytes, R, train, L.total, L.S);
    if (!w.ok) {
        linr_set_error("linr_net_forward: workspace too small (%zu < %zu)", ws_bytes, w.used);
        return LINR_ENOMEM;
    }
    if (R == 0) {
        if (d_bits && (phases & FWD_POST)) LINR_CHECK_CUDA(cudaMemsetAsync(d_bits, 0, sizeof(double), s));
        return LINR_OK;
    }
    const RowMap m = map_of(rows);
    const int jl = (lo > 1 ? lo : 1) - 1, jh = hi - 1;   // LDFE groups of this stage range
    BankCtx bank_ctx;
    BankScope bank_scope;
    if (train) bank_begin(bank_ctx, L, d_params, w.stage, 0, 4, s);   // training forward: weights via the constant bank
    // ConvA of the LDFE blocks reads occupancy bits only (teacher forcing) and its weights from shared memory: in a
    // training call it runs on the context's second stream beside SCE + ConvA of block_in
    SideLane sl(s);
    const bool a_bits = (phases & FWD_PRE) && jh > jl;
    const bool a_bits_side = a_bits && train && (phases & FWD_GDFE);
    if (a_bits_side)
        block_A_forward(d_params, L.ob + jl, jh - jl, m, true, rows->d_occ, jl + 1, 1, TN(), w.ob_y + jl * R * 8, w.ob_t1 + jl * R * 4, sl.leaf());
    if (phases & FWD_GDFE) {
        {  // SCE
            SceArgs a = sce_args(d_params, L, rows);
            a.f0 = T(w.f0, 0, 8);
            ProfScope prof(K_SCE, R, s);
            sce_fwd_kernel<<<(unsigned)ceil_div64(R, SCE_TPB), SCE_TPB, 0, s>>>(a);
        }
        // g = block_in(f0) -> hh[0] (group 7 of the block activation arrays)
        block_A_forward(d_params, &L.bin, 1, m, false, nullptr, 0, 0, T(w.f0, 0, 8), w.bi_y, w.bi_t1, s);
    }
    if (a_bits && !a_bits_side)   // ConvA of the LDFE blocks (occupancy bits, teacher forcing)
        block_A_forward(d_params, L.ob + jl, jh - jl, m, true, rows->d_occ, jl + 1, 1, TN(), w.ob_y + jl * R * 8, w.ob_t1 + jl * R * 4, s);
    sl.join();
    // the five inner layers of the blocks in multi-group launches
    {
        const bool gd = phases & FWD_GDFE, pre = (phases & FWD_PRE) && jh > jl;
        if (gd && pre) for_block_groups(jl, jh, [&](int first, int count) { block_mid_forward(d_params, L.blk8 + first, count, m, bufs_at(w, first), s); });
        else if (gd) block_mid_forward(d_params, L.blk8 + 7, 1, m, bufs_at(w, 7), s);
        else if (pre) block_mid_forward(d_params, L.blk8 + jl, jh - jl, m, bufs_at(w, jl), s);
    }
    if (phases & FWD_GDFE) block_B_forward(d_params, &L.bin, 1, m, w.bi_z, T(w.hh, 0, 8), TN(), s);
    if (phases & FWD_POST) {
        // ConvB: hh[k+1] = g + LDFE_k(occ[:, :k+1])
        if (jh > jl)
            block_B_forward(d_params, L.ob + jl, jh - jl, m, w.ob_z + jl * R * 8, T(w.hh + (jl + 1) * R * 8, R * 8, 8), T(w.hh, 0, 8), s);
        // heads of the stages [lo, hi)
        const bool want_bits = d_bits != nullptr || train;
        const int head_gx = head_forward(d_params, L, m, lo, hi - lo, T(w.hh + lo * R * 8, R * 8, 8), train ? w.hc + lo * R * 8 : nullptr,
                                         rows->d_occ, d_probs, d_cdf, train ? w.dzs : nullptr, loss_scale * 1.4426950408889634f,
                                         want_bits ? w.bits_partial : nullptr, 0, s);
        if (d_bits) {
            ProfScope prof(K_REDUCE, 1, s);
            bits_finalize_kernel<<<1, 256, 0, s>>>(w.bits_partial, head_gx * (hi - lo), d_bits);
        }
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

static int net_backward_impl(const float *d_params, int scale_num, const linr_rows *rows, int lo, int hi, int phases, int own_gdfe,
                             float *d_grad, void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_rows(rows, scale_num, true);
    if (rc) return rc;
    LINR_REQUIRE(lo >= 0 && lo < hi && hi <= 8, "stage range [%d,%d) is not inside [0,8)", lo, hi);
    LINR_REQUIRE(phases > 0 && (phases & ~BWD_ALL) == 0, "backward phase mask %d out of range", phases);
    const Layout &L = layout_for(scale_num);
    const int64_t R = rows->n_rows;
    const int P = L.total;
    NetWs w = carve_net(d_ws, ws_bytes, R, 1, P, L.S);
    if (!w.ok) {
        linr_set_error("linr_net_backward: workspace too small (%zu < %zu)", ws_bytes, w.used);
        return LINR_ENOMEM;
    }
    const bool partial_range = lo != 0 || hi != 8 || !own_gdfe;
    if ((phases & BWD_FINAL) && (R == 0 || partial_range))
        LINR_CHECK_CUDA(cudaMemsetAsync(d_grad, 0, sizeof(float) * P, s));   // parameters this rank did not touch: zero
    if (R == 0) return LINR_OK;
    const RowMap m = map_of(rows);
    const int jl = (lo > 1 ? lo : 1) - 1, jh = hi - 1, G = hi - lo;
    BankCtx bank_ctx;
    BankScope bank_scope;
    if (phases & (BWD_HEADS | BWD_LDFE | BWD_GDFE)) bank_begin(bank_ctx, L, d_params, w.stage, 4, 8, s);
    BlockGrads gr{w.g_dz, w.g_dt0, w.g_dy, w.g_dt2, w.g_dt1};
    SideLane sl(s);
    auto blocks_backward = [&](int first, int count) {
        BlockGrads g2{gr.dz + first * R * 8, gr.dt0 + first * R * 4, gr.dy + first * R * 8, gr.dt2 + first * R * 4, gr.dt1 + first * R * 4};
        // output gradients: dhh[j+1] for LDFE block j, dg = dhh[8] for block_in = group 7
        block_Bmid_backward(d_params, L.blk8 + first, count, m, w, P, bufs_at(w, first), g2, T(w.dhh + (first + 1) * R * 8, R * 8, 8), s, sl);
    };
    if (phases & BWD_HEADS) {
        // heads: dc, MLP weight partials, SConv weight partials, dh_k
        HeadBwdArgs a;
        memset(&a, 0, sizeof(a));
        a.n_rows = R, a.params = d_params;
        for (int g = 0; g < G; ++g) {
            const int k = lo + g;
            a.w1_off[g] = L.mlp_w1[k], a.b1_off[g] = L.mlp_b1[k], a.w2_off[g] = L.mlp_w2[k], a.b2_off[g] = L.mlp_b2[k];
        }
        a.c = T(w.hc + lo * R * 8, R * 8, 8), a.dz = w.dzs + lo * R, a.dc = T(w.dc + lo * R * 8, R * 8, 8);
        a.partial = w.partial, a.P = P, a.chunk = w.chunk;
        {
            ProfScope prof(K_HEADBWD, R * G, s);
            head_bwd_rows_kernel<<<dim3((unsigned)ceil_div64(R, 128 * HEAD_RPT), (unsigned)G), 128, 0, s>>>(a);
        }
        {
            cudaStream_t l = sl.leaf();
            {
                ProfScope prof(K_HEADBWD, R * G, l);
                head_bwd_w_kernel<<<dim3((unsigned)w.n_chunks, (unsigned)G), 256, 0, l>>>(a);
            }
            launch_bwd_w<8, 8, 0>(m, w, P, L.pr_w + lo, L.pr_b + lo, G, T(w.hh + lo * R * 8, R * 8, 8), T(w.dc + lo * R * 8, R * 8, 8), nullptr, 0, 0, l);
        }
        {
            ConvArgs ca = conv_args(m, d_params);
            for (int g = 0; g < G; ++g) ca.w_off[g] = L.pr_w[lo + g];
            ca.flip = 1, ca.x = T(w.dc + lo * R * 8, R * 8, 8), ca.y = T(w.dhh + lo * R * 8, R * 8, 8);
            launch_conv<8, 8, 0>(ca, G, s);
        }
        // every h_k contains g: dg = sum over this range's stages of dh_k.  The backward pass is linear in dg, so the ranks
        // of a stage split add their pieces (one reduce of [R,8] floats) before the owner of block_in runs BWD_GDFE.
        ProfScope prof(K_REDUCE, R, s);
        sum_groups_kernel<<<(unsigned)ceil_div64(R * 2, 256), 256, 0, s>>>(w.dhh + lo * R * 8, R * 8, G, R * 2, w.dg);
    }
    {
        const bool ld = (phases & BWD_LDFE) && jh > jl, gd = phases & BWD_GDFE;
        // ConvB + inner layers of the blocks
        if (ld && gd) for_block_groups(jl, jh, blocks_backward);
        else if (ld) blocks_backward(jl, jh - jl);
        else if (gd) blocks_backward(7, 1);
        // ConvA: LDFE blocks read occupancy bits (no input gradient); block_in propagates to f0
        if (ld)
            block_A_backward(d_params, L.ob + jl, jh - jl, m, w, P, true, rows->d_occ, jl + 1, 1, TN(), T(w.g_dy + jl * R * 8, R * 8, 8), TN(), s, sl);
        if (gd) {
            block_A_backward(d_params, &L.bin, 1, m, w, P, false, nullptr, 0, 0, T(w.f0, 0, 8), T(w.g_dy + 7 * R * 8, 0, 8), T(w.df0, 0, 8), s, sl);
            SceArgs sa = sce_args(d_params, L, rows);
            sa.df0 = T(w.df0, 0, 8), sa.chunk = w.chunk;
            ProfScope prof(K_SCE, R, s);
            sce_bwd_kernel<<<(unsigned)w.n_chunks, SCE_BWD_TPB, 0, s>>>(sa, w.sce_rec);
        }
    }
    sl.join();   // the chunk partials of the weight gradients are complete on the caller's stream from here
    if (phases & BWD_FINAL) {
        if (own_gdfe) {
            SceArgs sa = sce_args(d_params, L, rows);
            sa.df0 = T(w.df0, 0, 8), sa.chunk = w.chunk;
            ProfScope prof(K_SCE, L.S, s);
            sce_finalize_kernel<<<(unsigned)L.S, 1024, 0, s>>>(sa, w.sce_rec, w.n_chunks, d_grad);
        }
        // chunk partials -> gradient, only over the parameter ranges this rank wrote
        auto finalize = [&](int first, int end) {
            const int64_t cnt = end - first;
            if (cnt <= 0) return;
            ProfScope prof(K_REDUCE, cnt, s);
            finalize_grad_kernel<<<(unsigned)ceil_div64(cnt, 256), 256, 0, s>>>(w.partial, P, w.n_chunks, first, cnt, d_grad);
        };
        if (!partial_range) {
            finalize(L.conv_first, P);
        } else {
            if (own_gdfe) finalize(L.conv_first, L.mlp_w1[0]);                       // block_in
            finalize(L.mlp_w1[lo], hi < 8 ? L.mlp_w1[hi] : L.pr_w[0]);                   // MLP_k of the stages
            finalize(L.pr_w[lo], hi < 8 ? L.pr_w[hi] : L.ob[0].A_w);                     // SConv_k of the stages
            if (jh > jl) finalize(L.ob[jl].A_w, jh < 7 ? L.ob[jh].A_w : P);              // LDFE blocks
        }
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

}  // namespace

// ============================================================================================== C ABI
extern "C" {

int linr_version(void) { return 100; }
const char *linr_last_error(void) { return g_err; }

int linr_device_info(int device, int *sm_count, int64_t *l2_bytes) {
    int sm = 0, l2 = 0;
    LINR_CHECK_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device));
    LINR_CHECK_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, device));
    if (sm_count) *sm_count = sm;
    if (l2_bytes) *l2_bytes = l2;
    return LINR_OK;
}

int linr_ctx_create(int device, linr_ctx **out) {
    LINR_REQUIRE(out != nullptr, "linr_ctx_create: null output");
    int n = 0;
    LINR_CHECK_CUDA(cudaGetDeviceCount(&n));
    LINR_REQUIRE(device >= 0 && device < n && device < 64, "linr_ctx_create: device %d out of range", device);
    linr_ctx *c = new linr_ctx();
    c->device = device;
    *out = c;
    return LINR_OK;
}
int linr_ctx_destroy(linr_ctx *ctx) {
    if (!ctx) return LINR_OK;
    LINR_REQUIRE(!ctx->is_default, "linr_ctx_destroy: not a context made by linr_ctx_create");
    if (t_ctx == ctx) t_ctx = nullptr;
    {
        BankState &b = g_bank[ctx->device];
        std::lock_guard<std::mutex> lock(b.mu);
        LINR_REQUIRE(b.holder != ctx, "linr_ctx_destroy: the context is inside a training call");
        if (b.last == ctx) {
            // its launches may still read the bank: wait for them, then nobody guards the bank's contents
            if (ctx->done) cudaEventSynchronize(ctx->done);
            b.last = nullptr;
        }
    }
    if (ctx->done) cudaEventDestroy(ctx->done);
    if (ctx->side) {
        cudaStreamSynchronize(ctx->side);
        cudaStreamDestroy(ctx->side);
    }
    if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
    for (auto &e : ctx->fork_ev)
        if (e) cudaEventDestroy(e);
    delete ctx;
    return LINR_OK;
}
int linr_ctx_set_current(linr_ctx *ctx) {
    t_ctx = ctx;
    return LINR_OK;
}
int linr_ctx_hint_same_params(linr_ctx *ctx) {
    (ctx ? ctx : current_ctx())->same_params = 1;
    return LINR_OK;
}
int linr_side_stream_enable(int on) { return g_side_on.exchange(on ? 1 : 0); }
int64_t linr_ctx_bank_calls(const linr_ctx *ctx) {
    const linr_ctx *c = ctx ? ctx : current_ctx();
    return c->bank_calls.load(std::memory_order_relaxed);
}
int64_t linr_ctx_bank_launches(const linr_ctx *ctx) {
    const linr_ctx *c = ctx ? ctx : current_ctx();
    return c->bank_launches.load(std::memory_order_relaxed);
}

int linr_prof_enable(uint32_t class_mask) {
    std::lock_guard<std::mutex> lock(g_prof.mu);
    for (int c = 0; c < K_NCLASS; ++c) {
        prof_drain(c);
        g_prof.launches[c].store(0), g_prof.units[c].store(0), g_prof.ms_done[c] = 0.0;
    }
    g_prof.mask.store(class_mask);
    return LINR_OK;
}
int linr_prof_read(int cls, double *ms_total, int64_t *launches, int64_t *units) {
    LINR_REQUIRE(cls >= 0 && cls < K_NCLASS, "profiler class out of range");
    std::lock_guard<std::mutex> lock(g_prof.mu);
    prof_drain(cls);
    if (ms_total) *ms_total = g_prof.ms_done[cls];
    if (launches) *launches = g_prof.launches[cls].load();
    if (units) *units = g_prof.units[cls].load();
    return LINR_OK;
}
int linr_prof_classes(void) { return K_NCLASS; }
const char *linr_prof_name(int cls) { return cls >= 0 && cls < K_NCLASS ? kNames[cls] : ""; }

int64_t linr_param_count(int scale_num) {
    if (scale_num < 1 || scale_num > MAXS) return -1;
    return layout_for(scale_num).total;
}
int linr_param_offsets(int scale_num, int64_t *h_offsets, int cap) {
    if (scale_num < 1 || scale_num > MAXS) return LINR_EINVAL;
    const Layout &L = layout_for(scale_num);
    for (int i = 0; i < (int)L.offsets.size() && i < cap; ++i) h_offsets[i] = L.offsets[i];
    return (int)L.offsets.size();
}

size_t linr_net_ws_bytes(int64_t n_rows, int train) {
    NetWs w = carve_net(nullptr, ~(size_t)0, n_rows, train, (int)layout_for(MAXS).total, MAXS);
    return w.used + 4096;
}

int linr_net_forward(const float *d_params, int scale_num, const linr_rows *rows, int train, float loss_scale,
                     float *d_probs, uint16_t *d_cdf, double *d_bits, void *d_ws, size_t ws_bytes, void *stream) {
    return net_forward_impl(d_params, scale_num, rows, 0, 8, FWD_ALL, train, loss_scale, d_probs, d_cdf, d_bits, d_ws, ws_bytes, stream);
}

int linr_net_backward(const float *d_params, int scale_num, const linr_rows *rows, float *d_grad, void *d_ws,
                      size_t ws_bytes, void *stream) {
    return net_backward_impl(d_params, scale_num, rows, 0, 8, BWD_ALL, 1, d_grad, d_ws, ws_bytes, stream);
}

int linr_net_forward_stages(const float *d_params, int scale_num, const linr_rows *rows, int stage_lo, int stage_hi, int phases,
                            int train, float loss_scale, float *d_probs, uint16_t *d_cdf, double *d_bits, void *d_ws, size_t ws_bytes,
                            void *stream) {
    return net_forward_impl(d_params, scale_num, rows, stage_lo, stage_hi, phases, train, loss_scale, d_probs, d_cdf, d_bits, d_ws,
                            ws_bytes, stream);
}

int linr_net_backward_stages(const float *d_params, int scale_num, const linr_rows *rows, int stage_lo, int stage_hi, int phases,
                             int own_gdfe, float *d_grad, void *d_ws, size_t ws_bytes, void *stream) {
    return net_backward_impl(d_params, scale_num, rows, stage_lo, stage_hi, phases, own_gdfe, d_grad, d_ws, ws_bytes, stream);
}

int linr_net_ws_offsets(int64_t n_rows, int train, int scale_num, int64_t *h_g_bytes, int64_t *h_dg_bytes) {
    LINR_REQUIRE(scale_num >= 1 && scale_num <= MAXS, "scale_num out of range");
    const Layout &L = layout_for(scale_num);
    char *base = reinterpret_cast<char *>(4096);
    NetWs w = carve_net(base, ~(size_t)0 >> 1, n_rows, train, L.total, L.S);
    if (h_g_bytes) *h_g_bytes = reinterpret_cast<char *>(w.hh) - base;
    if (h_dg_bytes) *h_dg_bytes = train ? reinterpret_cast<char *>(w.dg) - base : -1;
    return LINR_OK;
}

int linr_net_decode_begin(const float *d_params, int scale_num, const linr_rows *rows, void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_rows(rows, scale_num, false);
    if (rc) return rc;
    const Layout &L = layout_for(scale_num);
    const int64_t R = rows->n_rows;
    NetWs w = carve_net(d_ws, ws_bytes, R, 0, L.total, L.S);
    if (!w.ok) {
        linr_set_error("linr_net_decode_begin: workspace too small (%zu < %zu)", ws_bytes, w.used);
        return LINR_ENOMEM;
    }
    if (R == 0) return LINR_OK;
    const RowMap m = map_of(rows);
    SceArgs a = sce_args(d_params, L, rows);
    a.f0 = T(w.f0, 0, 8);
    {
        ProfScope prof(K_SCE, R, s);
        sce_fwd_kernel<<<(unsigned)ceil_div64(R, SCE_TPB), SCE_TPB, 0, s>>>(a);
    }
    BlockBufs bi{w.bi_y, w.bi_t1, w.bi_t0, w.bi_t2, w.bi_z};
    block_forward(d_params, &L.bin, 1, m, false, nullptr, 0, 0, T(w.f0, 0, 8), bi, T(w.hh, 0, 8), TN(), s);
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_net_decode_stage(const float *d_params, int scale_num, const linr_rows *rows, int stage, float *d_probs_stage,
                          uint16_t *d_cdf_stage, void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_rows(rows, scale_num, true);
    if (rc) return rc;
    LINR_REQUIRE(stage >= 0 && stage < 8, "stage out of range");
    const Layout &L = layout_for(scale_num);
    const int64_t R = rows->n_rows;
    NetWs w = carve_net(d_ws, ws_bytes, R, 0, L.total, L.S);
    if (!w.ok) {
        linr_set_error("linr_net_decode_stage: workspace too small");
        return LINR_ENOMEM;
    }
    if (R == 0) return LINR_OK;
    const RowMap m = map_of(rows);
    Tens h = T(w.hh, 0, 8);
    if (stage > 0) {
        // same kernels, same per-row arithmetic as the batched encoder path: one group, cin = stage
        BlockBufs ob{w.ob_y, w.ob_t1, w.ob_t0, w.ob_t2, w.ob_z};
        block_forward(d_params, &L.ob[stage - 1], 1, m, true, rows->d_occ, stage, 0, TN(), ob, T(w.hh + R * 8, 0, 8), T(w.hh, 0, 8), s);
        h = T(w.hh + R * 8, 0, 8);
    }
    head_forward(d_params, L, m, stage, 1, h, nullptr, rows->d_occ, d_probs_stage, d_cdf_stage, nullptr, 0.f, nullptr, stage, s);
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_occ_set_stage(uint8_t *d_occ, const uint8_t *d_sym, int64_t n_rows, int stage, void *stream) {
    LINR_REQUIRE(stage >= 0 && stage < 8, "stage out of range");
    if (n_rows <= 0) return LINR_OK;
    ProfScope prof(K_ADAM, n_rows, (cudaStream_t)stream);
    occ_set_stage_kernel<<<(unsigned)ceil_div64(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(d_occ, d_sym, n_rows, stage);
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_net_decode_scale(const float *d_params, int scale_num, const linr_rows *rows, const uint8_t *const *h_streams,
                          const int64_t *h_nbytes, uint16_t *d_cdf, uint8_t *d_sym, uint16_t *h_cdf, uint8_t *h_sym,
                          void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_rows(rows, scale_num, true);
    if (rc) return rc;
    const int64_t n = rows->n_rows;
    if (n == 0) return LINR_OK;
    LINR_REQUIRE(h_streams && h_nbytes && d_cdf && d_sym && h_cdf && h_sym, "linr_net_decode_scale: null buffer");
    rc = linr_net_decode_begin(d_params, scale_num, rows, d_ws, ws_bytes, stream);
    if (rc) return rc;
    for (int k = 0; k < 8; ++k) {
        rc = linr_net_decode_stage(d_params, scale_num, rows, k, nullptr, d_cdf, d_ws, ws_bytes, stream);
        if (rc) return rc;
        LINR_CHECK_CUDA(cudaMemcpyAsync(h_cdf, d_cdf, sizeof(uint16_t) * n, cudaMemcpyDeviceToHost, s));
        LINR_CHECK_CUDA(cudaStreamSynchronize(s));
        rc = linr_rc_decode_binary(h_cdf, h_streams[k], h_nbytes[k], h_sym, n);
        if (rc) return rc;
        // h_sym is rewritten only after the next stage's synchronise, which follows this copy in stream order
        LINR_CHECK_CUDA(cudaMemcpyAsync(d_sym, h_sym, (size_t)n, cudaMemcpyHostToDevice, s));
        rc = linr_occ_set_stage(const_cast<uint8_t *>(rows->d_occ), d_sym, n, k, stream);
        if (rc) return rc;
    }
    return LINR_OK;
}

int linr_net_decode_scale_batch(const float *d_params, int scale_num, const linr_rows *rows, int n_seg, const int64_t *h_seg_off,
                                const uint8_t *const *h_streams, const int64_t *h_nbytes, uint16_t *d_cdf, uint8_t *d_sym,
                                uint16_t *h_cdf, uint8_t *h_sym, int threads, void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_rows(rows, scale_num, true);
    if (rc) return rc;
    const int64_t n = rows->n_rows;
    if (n == 0 || n_seg <= 0) return LINR_OK;
    LINR_REQUIRE(h_seg_off && h_streams && h_nbytes && d_cdf && d_sym && h_cdf && h_sym, "linr_net_decode_scale_batch: null buffer");
    LINR_REQUIRE(h_seg_off[0] == 0 && h_seg_off[n_seg] == n, "linr_net_decode_scale_batch: segments must cover the rows");
    rc = linr_net_decode_begin(d_params, scale_num, rows, d_ws, ws_bytes, stream);
    if (rc) return rc;
    std::vector<const uint16_t *> cdfs(n_seg);
    std::vector<const uint8_t *> ins(n_seg);
    std::vector<uint8_t *> syms(n_seg);
    std::vector<int64_t> nb(n_seg), ns(n_seg);
    for (int f = 0; f < n_seg; ++f) cdfs[f] = h_cdf + h_seg_off[f], syms[f] = h_sym + h_seg_off[f], ns[f] = h_seg_off[f + 1] - h_seg_off[f];
    for (int k = 0; k < 8; ++k) {
        rc = linr_net_decode_stage(d_params, scale_num, rows, k, nullptr, d_cdf, d_ws, ws_bytes, stream);
        if (rc) return rc;
        LINR_CHECK_CUDA(cudaMemcpyAsync(h_cdf, d_cdf, sizeof(uint16_t) * n, cudaMemcpyDeviceToHost, s));
        LINR_CHECK_CUDA(cudaStreamSynchronize(s));
        for (int f = 0; f < n_seg; ++f) ins[f] = h_streams[f * 8 + k], nb[f] = h_nbytes[f * 8 + k];
        rc = linr_rc_decode_binary_batch(n_seg, cdfs.data(), ins.data(), nb.data(), syms.data(), ns.data(), threads);
        if (rc) return rc;
        LINR_CHECK_CUDA(cudaMemcpyAsync(d_sym, h_sym, (size_t)n, cudaMemcpyHostToDevice, s));
        rc = linr_occ_set_stage(const_cast<uint8_t *>(rows->d_occ), d_sym, n, k, stream);
        if (rc) return rc;
    }
    return LINR_OK;
}

// ---- single-layer entry points ---------------------------------------------------------------------------
static int conv_dims_ok(int cin, int cout) { return (cin == 4 || cin == 8) && (cout == 4 || cout == 8); }

static void conv_dispatch(const ConvArgs &a, int cin, int cout, cudaStream_t s) {
    if (cin == 8 && cout == 8) launch_conv<8, 8, 0>(a, 1, s);
    else if (cin == 8 && cout == 4) launch_conv<8, 4, 0>(a, 1, s);
    else if (cin == 4 && cout == 8) launch_conv<4, 8, 0>(a, 1, s);
    else launch_conv<4, 4, 0>(a, 1, s);
}

int linr_spconv27_fwd(const float *d_x, int cin, const float *d_w, const float *d_bias, float *d_y, int cout,
                      const linr_rows *rows, int relu, void *stream) {
    LINR_REQUIRE(conv_dims_ok(cin, cout), "linr_spconv27_fwd: channels must be 4 or 8 (got %d -> %d)", cin, cout);
    LINR_REQUIRE(rows && rows->ld >= rows->n_rows, "bad rows");
    ConvArgs a = conv_args(map_of(rows), d_w);
    a.w_off[0] = 0;
    a.bias_direct = d_bias;
    a.x = T(const_cast<float *>(d_x), 0, cin), a.y = T(d_y, 0, cout), a.relu = relu;
    conv_dispatch(a, cin, cout, (cudaStream_t)stream);
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_spconv27_bwd_in(const float *d_dy, int cin, const float *d_w, float *d_dx, int cout, const linr_rows *rows, void *stream) {
    LINR_REQUIRE(conv_dims_ok(cin, cout), "linr_spconv27_bwd_in: channels must be 4 or 8");
    LINR_REQUIRE(rows && rows->ld >= rows->n_rows, "bad rows");
    ConvArgs a = conv_args(map_of(rows), d_w);
    a.w_off[0] = 0, a.flip = 1;
    a.x = T(const_cast<float *>(d_dy), 0, cout), a.y = T(d_dx, 0, cin);
    conv_dispatch(a, cout, cin, (cudaStream_t)stream);  // this launch maps cout channels -> cin channels
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

size_t linr_spconv27_bwd_w_ws_bytes(int64_t n_rows, int cin, int cout) {
    return (size_t)chunks_for(n_rows) * (27 * cin * cout + cout) * sizeof(float) + 256;
}

int linr_spconv27_bwd_w(const float *d_x, int cin, const float *d_dy, int cout, const linr_rows *rows, float *d_dw,
                        float *d_dbias, void *d_ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    LINR_REQUIRE(conv_dims_ok(cin, cout) && !(cin == 4 && cout == 8), "linr_spconv27_bwd_w: unsupported channels %d -> %d", cin, cout);
    LINR_REQUIRE(rows && rows->ld >= rows->n_rows, "bad rows");
    const int64_t R = rows->n_rows;
    const int P = 27 * cin * cout + cout;
    NetWs w;
    set_chunks(w, R);
    LINR_REQUIRE(ws_bytes >= (size_t)w.n_chunks * P * sizeof(float), "linr_spconv27_bwd_w: workspace too small");
    w.partial = (float *)d_ws;
    const int wo[1] = {0}, bo[1] = {27 * cin * cout};
    const RowMap m = map_of(rows);
    const Tens x = T(const_cast<float *>(d_x), 0, cin), dy = T(const_cast<float *>(d_dy), 0, cout);
    if (cin == 8 && cout == 8) launch_bwd_w<8, 8, 0>(m, w, P, wo, bo, 1, x, dy, nullptr, 0, 0, s);
    else if (cin == 8 && cout == 4) launch_bwd_w<8, 4, 0>(m, w, P, wo, bo, 1, x, dy, nullptr, 0, 0, s);
    else launch_bwd_w<4, 4, 0>(m, w, P, wo, bo, 1, x, dy, nullptr, 0, 0, s);
    {
        ProfScope prof(K_REDUCE, P, s);
        finalize_grad_kernel<<<(unsigned)ceil_div64(27 * cin * cout, 256), 256, 0, s>>>(w.partial, P, w.n_chunks, 0, 27 * cin * cout, d_dw);
    }
    if (d_dbias) {
        ProfScope prof(K_REDUCE, cout, s);
        finalize_grad_kernel<<<1, 256, 0, s>>>(w.partial, P, w.n_chunks, 27 * cin * cout, cout, d_dbias - 27 * cin * cout);
    }
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_adam_fused(float *d_params, const float *d_grad, float *d_m, float *d_v, int64_t n, int64_t step, float lr,
                    float beta1, float beta2, float eps, float weight_decay, void *stream) {
    LINR_REQUIRE(step >= 1, "Adam step is 1-based");
    if (n <= 0) return LINR_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    ProfScope prof(K_ADAM, n, (cudaStream_t)stream);
    adam_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(d_params, d_grad, d_m, d_v, n, lr, beta1, beta2, eps,
                                                                                 weight_decay, (float)bc1, (float)sqrt(bc2));
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_param_quant(const float *d_params, int64_t n, int bitdepth, uint8_t *d_q, float *d_recon, float *d_stats, void *stream) {
    LINR_REQUIRE(bitdepth >= 1 && bitdepth <= 8, "bitdepth must be in [1,8]");
    LINR_REQUIRE(n > 0, "empty parameter vector");
    ProfScope prof(K_ADAM, n, (cudaStream_t)stream);
    quant_kernel<uint8_t><<<1, 1024, 0, (cudaStream_t)stream>>>(d_params, n, (float)((1 << bitdepth) - 1), d_q, d_recon, d_stats);
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

int linr_param_quant16(const float *d_params, int64_t n, int bitdepth, uint16_t *d_q, float *d_recon, float *d_stats, void *stream) {
    LINR_REQUIRE(bitdepth >= 1 && bitdepth <= 16, "bitdepth must be in [1,16]");
    LINR_REQUIRE(n > 0, "empty parameter vector");
    ProfScope prof(K_ADAM, n, (cudaStream_t)stream);
    quant_kernel<uint16_t><<<1, 1024, 0, (cudaStream_t)stream>>>(d_params, n, (float)((1 << bitdepth) - 1), d_q, d_recon, d_stats);
    LINR_LAUNCH_CHECK();
    return LINR_OK;
}

}  // extern "C"
