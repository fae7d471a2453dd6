/* linr_b200.h — C ABI of the B200-native LINR-PCGC overfit + coding hot path.
 *
 * Drop-in boundary: these entry points are what a maintainer of the reference
 * binds (ctypes) in place of the work the reference hands to its un-vendored
 * natives — MinkowskiEngine 0.5.4 (coordinate hash, kernel map, sparse conv fwd/bwd), torch
 * (unique/sort/searchsorted/Linear/BCELoss/Adam) and torchac 0.9.3 (range coder).  Every function
 * cites the reference interface it replaces as file:line under /root/reference.
 *
 * Conventions
 *  - plain pointers and sizes only; device pointers are marked `d_`, host pointers `h_`;
 *  - the CALLER owns all memory (outputs and workspaces are pre-allocated; *_ws_bytes() sizes them);
 *  - `stream` is a cudaStream_t passed as void*; every device function is stream-ordered and does
 *    NOT synchronise unless its comment says so;
 *  - return 0 on success, negative LINR_E* on failure; linr_last_error() gives the thread-local text;
 *  - coordinates are non-negative int32 < 2^20, rows of one coordinate set are x-major
 *    lexicographically sorted and unique (the reference's invariant, datautils/custom_dataset.py:308);
 *  - all floating point is fp32 with a fixed per-row accumulation order (offset-column major, then
 *    input channel), independent of grid shape / batching, so encoder-side and decoder-side
 *    probabilities are bit-identical (SURVEY.md section 7 "hard parts").
 */
#ifndef LINR_B200_H
#define LINR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LINR_OK 0
#define LINR_EINVAL (-1)
#define LINR_ECUDA (-2)
#define LINR_ENOMEM (-3) /* caller workspace too small */

int linr_version(void);
const char *linr_last_error(void);
/* device properties the host side needs for grid sizing (SM count, L2 bytes). Synchronous. */
int linr_device_info(int device, int *sm_count, int64_t *l2_bytes);

/* Contexts ("no global state except an opaque, explicitly created/destroyed linr_ctx* per device").
 * A context stands for one user of a device -- one trainer, one coder -- driven by one host thread at a time;
 * linr_ctx_set_current binds it to the calling host thread, and entry points called with no current context use a
 * per-device default context (never destroyed).  What a context carries: its turn at the device's constant weight
 * bank and its launch statistics.  The training kernels read their conv weights from the bank; a training call
 * (linr_net_forward with train != 0, linr_net_backward) holds the bank from its first fill to its end, then records an
 * event on its stream; the next holder -- any context, any stream -- makes its stream wait for that event before it
 * fills the bank, so trainers that run one after the other all use the fast kernels, and a trainer that finds the bank
 * held by a call of ANOTHER host thread runs the shared-memory kernel variant for that call (same bits).
 * linr_ctx_destroy waits for the context's last training call if its weights are still the bank's contents.
 * Everything else a call needs travels in its arguments; the SM count is cached per device; the profiler below is
 * process-wide by design (bench.py reads one table).
 *  linr_ctx_bank_calls / linr_ctx_bank_launches (NULL = the calling thread's current context): training calls that
 *  held the bank, and constant-bank conv launches made, through the context. */
typedef struct linr_ctx linr_ctx;
int linr_ctx_create(int device, linr_ctx **out);
int linr_ctx_destroy(linr_ctx *ctx);
int linr_ctx_set_current(linr_ctx *ctx);
/* One-shot hint for the NEXT training call through the context: it gets the same parameter values and the same workspace
 * as the previous call of the same direction (forward / backward), so the conv weights staged by that call are still
 * valid and are not staged again.  Used between the phases of one iteration (linr_net_*_stages); cleared by the call. */
int linr_ctx_hint_same_params(linr_ctx *ctx);
int64_t linr_ctx_bank_calls(const linr_ctx *ctx);
int64_t linr_ctx_bank_launches(const linr_ctx *ctx);
/* Second stream of a context (no reference counterpart: autograd runs the reference's backward on one stream,
 * main.py:316).  A training call made through an explicit context runs the kernels nothing in the call depends on -- the
 * weight-gradient launches of linr_net_backward(_stages), the bit-input ConvA of the LDFE blocks in the training forward --
 * on a stream owned by the context, forked from and joined to the caller's stream by events inside the call: when the
 * call returns, everything it launched is ordered before later work on the caller's stream, as without it.  The results
 * are bit-identical either way (same kernels, same partial sums).  linr_side_stream_enable(0) keeps every launch on
 * the caller's stream (process-wide switch, returns the previous setting; also LINR_NO_SIDE_STREAM=1): bench.py uses
 * it to time kernel classes one at a time. */
int linr_side_stream_enable(int on);

/* Launch accounting / live kernel timing (no reference counterpart: the reference has no profiler hooks,
 * SURVEY.md section 5).  Kernel classes are the K_* values of csrc/prof.cuh; linr_prof_name() names them.
 *  linr_prof_enable(mask): bit c set -> launches of class c are bracketed by CUDA events on their stream
 *                          (mask 0 disables).  Resets all counters.
 *  linr_prof_read(c, ...): synchronises the recorded events of class c; total milliseconds, launches and the
 *                          "units" (rows x groups) those launches processed.  Launch counts are kept for
 *                          every class even when its timing is disabled. */
int linr_prof_enable(uint32_t class_mask);
int linr_prof_read(int cls, double *ms_total, int64_t *launches, int64_t *units);
int linr_prof_classes(void);
const char *linr_prof_name(int cls);

/* ------------------------------------------------------------------------------------------------
 * Coordinate stage (integer, bit-exact).  Replaces torch.unique/sort/searchsorted chains.
 * ---------------------------------------------------------------------------------------------- */

/* Workspace needed by sort-based functions below for up to n items. */
size_t linr_coord_ws_bytes(int64_t n);

/* torch.unique(xyz, dim=0) + x-major lexicographic order.
 * Replaces datautils/custom_dataset.py:280, models/module_utils.py:248 (QuickSearchCoord.__init__),
 * models/sort_functions.py:17-60.  `bits` = bit width of the largest coordinate (<=20).
 * d_n_out receives the number of unique rows (device int64). */
int linr_coord_sort_unique(const int32_t *d_xyz_in, int64_t n, int bits, int32_t *d_xyz_out, int64_t *d_n_out,
                           void *d_ws, size_t ws_bytes, void *stream);

/* Stable lexicographic sort without dedup: models/sort_functions.py:17-30 (sort_by_coord_sum_c). */
int linr_coord_sort(const int32_t *d_xyz_in, int64_t n, int bits, int32_t *d_xyz_out, void *d_ws, size_t ws_bytes,
                    void *stream);

/* Per-axis minimum over points and subtraction (datautils/custom_dataset.py:273-276). d_min: int32[3]. */
int linr_coord_min_sub(const int32_t *d_xyz_in, int64_t n, int32_t *d_xyz_out, int32_t *d_min, void *stream);

/* octree_level.forward (models/module_utils.py:97-115) + quantize (models/quantize_functions.py:19-30):
 * parents = unique(floor(child/2)); occ[n] bit i = child (2p + (i>>2&1, i>>1&1, i&1)) present.
 * d_parent_xyz / d_occ sized for nc rows (upper bound; d_occ rounded up to a multiple of 4 bytes). */
int linr_octree_down(const int32_t *d_child_xyz, int64_t nc, int bits, int32_t *d_parent_xyz, uint8_t *d_occ,
                     int64_t *d_n_parent, void *d_ws, size_t ws_bytes, void *stream);

/* octree_level.upper_layer (models/module_utils.py:117-127), two calls around one host read of the
 * child count: _count fills d_child_off[n+1] (exclusive scan of popcount(occ)); _expand writes the
 * n_child sorted children. */
int linr_octree_up_count(const uint8_t *d_occ, int64_t n, int64_t *d_child_off, void *d_ws, size_t ws_bytes, void *stream);
int linr_octree_up_expand(const int32_t *d_parent_xyz, const uint8_t *d_occ, const int64_t *d_child_off, int64_t n,
                          int64_t n_child, int bits, int32_t *d_child_xyz, void *d_ws, size_t ws_bytes, void *stream);

/* Open-addressing hash over the concatenated rows of all scales of a frame
 * (replaces ME's CoordinateManager.insert_and_map, reached from models/function_utils.py:16).
 * cap = power of two >= 2*n.  d_scale[row] tags the coordinate set a row belongs to. */
size_t linr_hash_bytes(int64_t cap);
int linr_hash_build(const int32_t *d_xyz, const uint8_t *d_scale, int64_t n, void *d_table, int64_t cap, void *stream);

/* Kernel map of the 3x3x3 stride-1 convolution + the 7 face-neighbour occupancy bits.
 * Replaces ME's kernel-map construction (implicit in models/upsample.py:21,90,95, models/resnet.py:15-51)
 * and qscTensor.set_offset_tensor (models/module_utils.py:210-213).
 *  d_nbr27   [n,27] int32 row or -1, offset k=(dx+1)+3(dy+1)+9(dz+1)   (optional, may be NULL)
 *  d_anchor  [9,ld] int32, d_mask [n] uint32: compact form — column c=(dx+1)+3(dy+1); bit 3c+j set if
 *            offset k=c+9j present; present rows of a column are consecutive starting at anchor[c]
 *  d_nbr7    [n] uint8: bit j = offsets_ini[j] present (main.py:24)                                  */
int linr_nbr_build(const int32_t *d_xyz, const uint8_t *d_scale, int64_t n, const void *d_table, int64_t cap,
                   int32_t *d_nbr27, int32_t *d_anchor, int64_t ld, uint32_t *d_mask, uint8_t *d_nbr7, void *stream);

/* QuickSearchCoord.search_coord_idx (models/module_utils.py:276-283): row of each query or -1. */
int linr_hash_lookup(const int32_t *d_query_xyz, const uint8_t *d_query_scale, int64_t nq, const void *d_table,
                     int64_t cap, int32_t *d_rows, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Network (fp32).  Parameters live in ONE flat buffer in `model.parameters()` order
 * (checkpoint contract, model_compression/model_size_est.py:391).
 * ---------------------------------------------------------------------------------------------- */

/* Number of parameters / offset of a named tensor in the flat buffer for `scale_num` scales. */
int64_t linr_param_count(int scale_num);
/* Writes up to `cap` offsets (in floats) in parameters() order, returns the tensor count (189 for 7). */
int linr_param_offsets(int scale_num, int64_t *h_offsets, int cap);

/* A frame (or one scale of it) as the kernels see it: R rows = concatenated parent voxels. */
typedef struct {
    int64_t n_rows;
    int64_t ld;              /* row stride of d_anchor (>= n_rows, multiple of 32) */
    const int32_t *d_anchor; /* [9,ld]   */
    const uint32_t *d_mask;  /* [n_rows] */
    const uint8_t *d_nbr7;   /* [n_rows] */
    const uint8_t *d_scale;  /* [n_rows] scale index of each row, non-decreasing (scales are concatenated in order) */
    const uint8_t *d_occ;    /* [n_rows] 8-bit child occupancy (teacher forcing / decoded so far) */
    const int32_t *d_tile_rng; /* [ceil(n_rows/128),6] from linr_tile_ranges, or NULL: per 128-row tile and dx = -1,0,+1 the
                                * row range [lo,hi) of its neighbours.  With it the conv / weight-gradient kernels stage their
                                * inputs in shared memory with bulk (TMA) copies; without it they gather through L1.  Same
                                * results bit for bit either way.  When set, d_mask must be readable up to `ld` entries. */
    const int32_t *d_pair_cnt;   /* [ceil(n_rows/256),32] and */
    const uint32_t *d_pair_list; /* [ceil(n_rows/256),27*256] from linr_pair_lists, or NULL: per 256-row tile and kernel offset
                                  * the existing (row, neighbour) pairs.  With them the weight-gradient kernel only touches
                                  * occupied offsets. */
} linr_rows;

/* Neighbour row ranges per 128-row tile (see linr_rows.d_tile_rng); d_rng int32 [ceil(n_rows/128), 6] =
 * (lo,hi) for dx = -1, 0, +1.  No reference counterpart (ME keeps per-offset in/out index lists instead). */
int linr_tile_ranges(const linr_rows *rows, int32_t *d_rng, void *stream);

/* Pair lists of the kernel map (see linr_rows.d_pair_cnt / d_pair_list): list (t,k) holds the rows of 256-row tile t that
 * have a neighbour at offset k as (row - 256 t) << 24 | neighbour_row; d_cnt[t*32 + k] is its length.  Same information
 * as ME's per-offset in/out kernel maps, tiled.  n_rows < 2^24.
 * Layout (made for the weight-gradient kernel, which fetches half of a tile's lists with one bulk copy and reads pairs
 * 32 at a time from shared memory): the 27 offsets are dealt to two halves, order28 = 2 x 14 offsets from
 * linr_pair_list_order (27 = no list); the lists of half h lie back to back from entry h*14*256 of the tile's 27*256
 * entries, in that order, each starting on a multiple of 4 entries.  Inside a list, aligned groups of four entries have
 * rows AND neighbours that differ modulo 4 wherever the kernel map allows it (bank-conflict-free 16-byte loads), the
 * remaining entries follow in row order. */
int linr_pair_lists(const linr_rows *rows, int32_t *d_cnt, uint32_t *d_list, void *stream);
void linr_pair_list_order(int32_t *order28);

/* Workspace sizes (bytes) for n_rows rows. `train`!=0 includes saved activations and gradients. */
size_t linr_net_ws_bytes(int64_t n_rows, int train);

/* Teacher-forced forward over all 8 stages (LINR_PCGC_Model.forward, models/model_core.py:72-81;
 * CNP.forward, models/upsample.py:163-217; InceptionResNet.forward, models/resnet.py:55-60).
 *  d_probs   [8,n_rows] f32 sigmoid outputs (stage-major), may be NULL
 *  d_cdf     [8,n_rows] u16 = round((1-p)*65534)+1, torchac's 16-bit CDF midpoint
 *            (models/module_utils.py:11-16 + torchac float->int conversion), may be NULL
 *  d_bits    double[1]: sum over rows and stages of BCE/ln2 with torch's log clamp at -100, may be NULL
 *  loss_scale: d(loss)/d(bits) used for the saved dz (1/point_num, main.py:315); ignored unless train. */
int linr_net_forward(const float *d_params, int scale_num, const linr_rows *rows, int train, float loss_scale,
                     float *d_probs, uint16_t *d_cdf, double *d_bits, void *d_ws, size_t ws_bytes, void *stream);

/* Backward of the above (autograd of main.py:316), deterministic: no floating-point atomics; block
 * partial sums are reduced in a fixed order.  Adds nothing: d_grad[param_count] is overwritten. */
int linr_net_backward(const float *d_params, int scale_num, const linr_rows *rows, float *d_grad, void *d_ws,
                      size_t ws_bytes, void *stream);

/* The same two passes restricted to the stages [stage_lo, stage_hi) of the 8, and cut into phases: a "stage split" of ONE
 * frame over several GPUs.  The reference steps the optimiser once per frame (main.py:305-321), so
 * frames cannot be dealt to ranks without changing the result; the 8 stages of a frame can: stage k needs
 * g = block_in(SCE) and its own LDFE block outter_blocks[k-1] + head k (models/upsample.py:203-216).  One rank of the
 * group owns SCE + block_in; the phases let the caller put the two exchanges of the split between the kernels:
 *   forward  phases (bit mask): 1 GDFE  SCE + block_in -> g           (owner; then g is broadcast, [n_rows,8] floats)
 *                               2 PRE   ConvA + inner layers of the range's LDFE blocks (occupancy bits in: no g needed,
 *                                       overlaps the broadcast)
 *                               4 POST  ConvB of those blocks + g, the heads of the range, d_probs / d_cdf rows of the
 *                                       range (others untouched), d_bits = bit count of the range
 *   backward phases (bit mask): 1 HEADS heads of the range, dh_k, dg = sum over the range of dh_k  (then dg is reduced
 *                                       to the owner, [n_rows,8] floats)
 *                               2 LDFE  the range's LDFE blocks (overlaps the reduce)
 *                               4 GDFE  block_in + SCE from the dg in the workspace (owner)
 *                               8 FINAL chunk partials -> d_grad (overwritten): zeros for parameters this rank did not
 *                                       touch; own_gdfe != 0 adds block_in + SCE
 * The SUM of d_grad over the ranks of a partition of [0,8) equals linr_net_backward's gradient up to fp32 summation order:
 * one all-reduce(sum) of the flat 219 kB vector, then the same fused Adam step on every rank.  phases 7 / 15 with
 * own_gdfe 1 on the range [0,8) are linr_net_forward / linr_net_backward.  linr_net_ws_offsets gives the byte offsets of
 * g and dg inside a workspace carved for n_rows (both [n_rows,8] fp32; dg is -1 without train). */
int linr_net_forward_stages(const float *d_params, int scale_num, const linr_rows *rows, int stage_lo, int stage_hi, int phases,
                            int train, float loss_scale, float *d_probs, uint16_t *d_cdf, double *d_bits, void *d_ws,
                            size_t ws_bytes, void *stream);
int linr_net_backward_stages(const float *d_params, int scale_num, const linr_rows *rows, int stage_lo, int stage_hi, int phases,
                             int own_gdfe, float *d_grad, void *d_ws, size_t ws_bytes, void *stream);
int linr_net_ws_offsets(int64_t n_rows, int train, int scale_num, int64_t *h_g_bytes, int64_t *h_dg_bytes);

/* Sequential decoding (CNP.decode, models/upsample.py:249-295), one coordinate set at a time:
 *  _begin: SCE + block_in (GDFE);  _stage k: [LDFE_{k-1} on the k bits decoded so far] + SConv_k + MLP_k
 *  -> d_cdf_stage[n_rows] (+ d_probs_stage optional).  rows->d_occ must hold stages < k. */
int linr_net_decode_begin(const float *d_params, int scale_num, const linr_rows *rows, void *d_ws, size_t ws_bytes,
                          void *stream);
int linr_net_decode_stage(const float *d_params, int scale_num, const linr_rows *rows, int stage, float *d_probs_stage,
                          uint16_t *d_cdf_stage, void *d_ws, size_t ws_bytes, void *stream);

/* Scatter one decoded stage back into the occupancy bytes: occ[row] |= sym[row] << stage. */
int linr_occ_set_stage(uint8_t *d_occ, const uint8_t *d_sym, int64_t n_rows, int stage, void *stream);

/* One whole scale of the sequential decoder in ONE call (CNP.decode, models/upsample.py:249-295, driven per scale by
 * decode_one_frame, decoder.py:153-176): _begin, then for k = 0..7: _stage k -> CDF midpoints to the host -> range
 * decoder on stream k (torchac.decode_float_cdf, models/module_utils.py:38) -> symbols back to the device ->
 * occ |= sym << k.  The eight device<->host round trips happen inside the call (it synchronises `stream` eight
 * times), so a host thread per frame keeps a GPU stream busy without holding an interpreter lock.
 *  h_streams[8] / h_nbytes[8]: the stage bitstreams (unpack_bitstream, models/function_utils.py:119-132)
 *  d_cdf u16[n_rows], d_sym u8[n_rows]: device scratch;  h_cdf u16[n_rows], h_sym u8[n_rows]: PINNED host scratch
 *  rows->d_occ: zero-filled by the caller, holds the decoded 8-bit occupancy on return. */
int linr_net_decode_scale(const float *d_params, int scale_num, const linr_rows *rows, const uint8_t *const *h_streams,
                          const int64_t *h_nbytes, uint16_t *d_cdf, uint8_t *d_sym, uint16_t *h_cdf, uint8_t *h_sym,
                          void *d_ws, size_t ws_bytes, void *stream);

/* The same for SEVERAL frames at once: the rows of `rows` are the concatenation of n_seg frames' parents of one scale
 * (segment f = rows [h_seg_off[f], h_seg_off[f+1]); the caller keeps the frames apart in space, e.g. by an x offset per
 * frame, so that no neighbourhood crosses a segment boundary).  Every stage is ONE set of launches over all frames, one
 * CDF download, n_seg range decoders on `threads` host threads, one symbol upload: 56 launch sets and round trips per
 * batch instead of per frame.  h_streams / h_nbytes are [n_seg][8]. */
int linr_net_decode_scale_batch(const float *d_params, int scale_num, const linr_rows *rows, int n_seg, const int64_t *h_seg_off,
                                const uint8_t *const *h_streams, const int64_t *h_nbytes, uint16_t *d_cdf, uint8_t *d_sym,
                                uint16_t *h_cdf, uint8_t *h_sym, int threads, void *d_ws, size_t ws_bytes, void *stream);

/* Single-layer entry points (used by the MinkowskiEngine-shaped shim and by unit tests).
 * ME.MinkowskiConvolution(kernel_size=3, stride=1) forward on one coordinate set (models/upsample.py:17,90,95):
 *   y[n,cout] = sum_k x[row(C+delta_k)] @ W[k] + bias;  W [27,cin,cout], bias [cout] or NULL; cin,cout in {4,8}. */
int linr_spconv27_fwd(const float *d_x, int cin, const float *d_w, const float *d_bias, float *d_y, int cout,
                      const linr_rows *rows, int relu, void *stream);
/* grad wrt input: dx[n,cin] = sum_k dy[row(C-delta_k)] @ W[k]^T */
int linr_spconv27_bwd_in(const float *d_dy, int cin, const float *d_w, float *d_dx, int cout, const linr_rows *rows,
                         void *stream);
/* grad wrt kernel and bias, deterministic two-pass reduction. d_ws from linr_spconv27_bwd_w_ws_bytes. */
size_t linr_spconv27_bwd_w_ws_bytes(int64_t n_rows, int cin, int cout);
int linr_spconv27_bwd_w(const float *d_x, int cin, const float *d_dy, int cout, const linr_rows *rows, float *d_dw,
                        float *d_dbias, void *d_ws, size_t ws_bytes, void *stream);

/* torch.optim.Adam (L2 weight decay added to the gradient) + the StepLR value, one launch over the flat
 * buffers (main.py:231-237,252,319-321).  step is 1-based. */
int linr_adam_fused(float *d_params, const float *d_grad, float *d_m, float *d_v, int64_t n, int64_t step, float lr,
                    float beta1, float beta2, float eps, float weight_decay, void *stream);

/* Model_Estimate.quant_uniform2 + Laplace stats (model_compression/model_size_est.py:72-91,410-411):
 *  d_q u8[n] symbols, d_recon f32[n] dequantised, d_stats float[4] = {min, max, mu, b}. bitdepth <= 8. */
int linr_param_quant(const float *d_params, int64_t n, int bitdepth, uint8_t *d_q, float *d_recon, float *d_stats,
                     void *stream);
/* The same with 16-bit symbols, bitdepth <= 16 (`--model_bitdepth` 9..16: the reference quantises these but cannot decode
 * them, model_compression/model_size_est.py:546-548 reads the symbols back as uint8). */
int linr_param_quant16(const float *d_params, int64_t n, int bitdepth, uint16_t *d_q, float *d_recon, float *d_stats,
                       void *stream);

/* ------------------------------------------------------------------------------------------------
 * Host range coder (stays on the CPU by design; replaces torchac.encode_float_cdf / decode_float_cdf,
 * models/module_utils.py:28,38 and model_compression/model_size_est.py:482,561).
 * ---------------------------------------------------------------------------------------------- */

/* Binary streams: cdf_mid[i] = 16-bit P(sym=0) boundary from linr_net_forward. Returns bytes written,
 * or -(needed) if cap is too small. */
int64_t linr_rc_encode_binary(const uint16_t *h_cdf_mid, const uint8_t *h_sym, int64_t n, uint8_t *h_out, int64_t cap);
int linr_rc_decode_binary(const uint16_t *h_cdf_mid, const uint8_t *h_in, int64_t nbytes, uint8_t *h_sym, int64_t n);
/* Many independent binary streams decoded at once on `threads` host threads (one stage of several frames). */
int linr_rc_decode_binary_batch(int n_streams, const uint16_t *const *h_cdf_mid, const uint8_t *const *h_in, const int64_t *nbytes,
                                uint8_t *const *h_sym, const int64_t *n, int threads);
/* Many independent binary streams at once on `threads` host threads (the 8 stages x S scales of a frame,
 * models/upsample.py:219-246).  Stream i codes bit h_shift[i] of each byte of h_sym[i] (h_shift NULL: bit 0), so
 * the 8 stage streams of a scale read the packed occupancy bytes in place.  h_written[i] = bytes, or -(needed). */
int linr_rc_encode_binary_batch(int n_streams, const uint16_t *const *h_cdf_mid, const uint8_t *const *h_sym,
                                const int *h_shift, const int64_t *n, uint8_t *const *h_out, const int64_t *cap,
                                int64_t *h_written, int threads);
/* General alphabet with one shared CDF row of Lp uint16 entries (model.bin Laplace coder). */
int64_t linr_rc_encode_shared(const uint16_t *h_cdf_row, int Lp, const int16_t *h_sym, int64_t n, uint8_t *h_out,
                              int64_t cap);
int linr_rc_decode_shared(const uint16_t *h_cdf_row, int Lp, const uint8_t *h_in, int64_t nbytes, int16_t *h_sym,
                          int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* LINR_B200_H */
