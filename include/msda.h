/*
 * msda.h -- C ABI of the B200 (sm_100a) multi-scale deformable attention library.
 *
 * This is the drop-in boundary for the native backend of the reference's models/ops package.
 * Every entry point takes plain pointers and sizes (no torch / ATen types), enqueues work on the
 * CUDA stream it is given and returns without synchronising.  The caller owns every buffer; the
 * library allocates no device memory of its own except inside an msda_host_session.
 *
 * What each function replaces in the reference (paths relative to /root/reference):
 *
 *   msda_forward    -> ms_deform_attn_forward / ms_deform_attn_cuda_forward
 *                      models/ops/src/ms_deform_attn.h:20-39, models/ops/src/cuda/ms_deform_attn_cuda.cu:20-80,
 *                      kernel models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299
 *   msda_backward   -> ms_deform_attn_backward / ms_deform_attn_cuda_backward
 *                      models/ops/src/ms_deform_attn.h:41-61, models/ops/src/cuda/ms_deform_attn_cuda.cu:83-153,
 *                      kernels models/ops/src/cuda/ms_deform_im2col_cuda.cuh:301-920, launcher :956-1327
 *   (pybind surface models/ops/src/vision.cpp:13-16 is rebuilt in Python over these two calls.)
 *
 * Tensor layouts (contiguous, row-major; identical to the reference):
 *   value         (N, S, M, D)            dtype T
 *   spatial_shapes(L, 2) int64 [H_l, W_l] DEVICE memory   level_start_index (L,) int64 DEVICE memory
 *   sampling_loc  (N, Lq, M, L, P, 2)     (x, y) normalised to [0,1]; dtype LOC(T)
 *   attn_weight   (N, Lq, M, L, P)        dtype LOC(T)
 *   output / grad_output (N, Lq, M, D)    dtype T   (viewed as (N, Lq, M*D) by the caller)
 * with T in {f32, f64, bf16} and LOC(f32)=f32, LOC(f64)=f64, LOC(bf16)=f32 (bf16 carries too few
 * mantissa bits for pixel coordinates; arithmetic and accumulation are fp32 for f32/bf16, fp64 for f64).
 *
 * Differences from the reference, all deliberate:
 *   - im2col_step (batch chunking, ms_deform_attn_cuda.cu:50-72) does not exist at this level: one launch
 *     covers the whole batch, so there is no "batch % im2col_step" restriction.
 *   - launch failures are returned (non-zero + msda_last_error()), not printf'd (.cuh:948-952, 1321-1325).
 *   - grad_sampling_loc and grad_attn_weight are fully overwritten (no zero-fill needed);
 *     grad_value is accumulated into and must be zero on entry unless MSDA_FLAG_ZERO_GRAD_VALUE is set.
 *   - bf16 is new; the reference dispatches float/double only (ms_deform_attn_cuda.cu:64,134).
 */
#ifndef MSDA_H_
#define MSDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_ABI_VERSION 1

/* element type of value / output / grad_output / grad_value */
enum { MSDA_F32 = 0, MSDA_F64 = 1, MSDA_BF16 = 2 };

/* status codes */
enum {
    MSDA_OK = 0,
    MSDA_ERR_INVALID_ARGUMENT = 1, /* null pointer, non-positive dimension, unknown dtype, overflow */
    MSDA_ERR_WORKSPACE = 2,        /* workspace missing or too small                               */
    MSDA_ERR_CUDA = 3,             /* a CUDA runtime call or kernel launch failed                  */
    MSDA_ERR_UNSUPPORTED = 4       /* shape outside what this build instantiates                   */
};

/* flags */
enum {
    MSDA_FLAG_ZERO_GRAD_VALUE = 1u << 0, /* backward: memset grad_value on the stream before accumulating   */
    MSDA_FLAG_DETERMINISTIC = 1u << 1,   /* backward (f32/bf16): bit-reproducible grad_value -- int64 fixed-point   */
                                         /* accumulation with integer atomics; needs the workspace                  */
    MSDA_FLAG_FORCE_GENERIC = 1u << 2,   /* testing: skip the specialised kernels, use the generic ones     */
    MSDA_FLAG_ALIGNED16 = 1u << 3        /* msda_backward_workspace_bytes only: the caller guarantees that every    */
                                         /* tensor it will pass is 16-byte aligned (lets bf16 report 0 bytes when   */
                                         /* the owned backward, which needs no fp32 image, will run)                */
};

/* Problem geometry.  All counts are element counts, not bytes. */
typedef struct msda_dims {
    int64_t batch;        /* N  */
    int64_t spatial_size; /* S = sum_l H_l*W_l */
    int64_t num_heads;    /* M  */
    int64_t channels;     /* D  (per head) */
    int64_t num_levels;   /* L  */
    int64_t num_query;    /* Lq */
    int64_t num_point;    /* P  */
} msda_dims;

int msda_abi_version(void);

/* Thread-local description of the last failure on the calling thread ("" if none). */
const char *msda_last_error(void);

/* Name of the kernel variant the last successful forward/backward on this thread selected
 * (e.g. "fwd_vec<f32,D32,L4,P4>" or "bwd_generic<f64>"); for tests and bench logs. */
const char *msda_last_kernel(void);

/* Kernel launches enqueued by this thread since the counter was last reset (memsets excluded). */
int64_t msda_launch_count(int reset);

/* Tuning / A-B testing knob (process-wide; for benchmarks and tests).  Keys:
 *   "variant"        forward: 0 auto (default: the row kernel; with "staged_auto" = 1 the staged forward for D=32
 *                    problems whose coarse levels -- ~S/4 pixels of a 4:1 pyramid -- fit in shared memory and that have
 *                    >= "staged_min_rows" (head, query) rows per image per SM) | 5 lean row kernel |
 *                    3 shared-memory-staged forward
 *   "staged_auto"    0 (default) | 1: see "variant" (the row kernel measured faster on every shape since round 2)
 *   "hoist"          0 | 1   (row forward: issue all tap loads of a row before the first FMA; D=32 L=P=4 only)
 *   "bf16_x4"        1 (default) | 0   (bf16 forward kernels at D=32 L=P=4: 8-byte lane chunks = the fp32 kernels' shape,
 *                    2.58 vs 2.68 ms at 800x1333 N=32; 0 = 16-byte chunks)
 *   "warps"          4 | 8   (row kernels: warps per CTA; D=32 L=P=4 only)
 *   "v3_threads"     512 | 768 | 1024   (staged forward CTA size)
 *   "bwd_mode"       backward strategy: 0 auto (default) | 1 row kernel only (every tap is a global vector red) |
 *                    2 row kernel + on-SM aggregation of the coarse levels (msda_bwd_binned) |
 *                    3 owned: every grad_value line written once by its owner, no zero-fill / workspace (sparse problems) |
 *                    4 planes: one kernel; the coarse levels' grad_value accumulated in shared memory as int32 fixed
 *                    point (native ATOMS.ADD, per-item power-of-two scale from a rigorous bound), fine levels by reds
 *   "planes_rows"    planes backward: query rows per work item / CTA (default 0 = 256 for 256-thread CTAs, else 1024)
 *   "planes_threads" planes backward CTA size: 768 (default) | 512 | 1024: one CTA per SM holding every level that fits |
 *                    256: four CTAs per SM, each with the smallest levels' planes (faster when timed alone, slower
 *                    inside a long step at the power cap)
 *   "planes_budget"  planes backward: cap in bytes on the shared-memory planes (default 2^30 = all; 0 = every level by reds)
 *   "planes_auto"    auto rule: 1 (default) = mode 4 for dense D=32 problems (more than "owned_max_taps" taps per value
 *                    pixel, >= "staged_min_rows" (head, query) rows per image per SM); 0 = never
 *   "staged_min_rows" see "variant"   (default 200)
 *   "staged_rows"    staged forward: query rows per work item / CTA (default 1024)
 *   "staged_persistent" staged forward A/B: 1 = one CTA per SM walking the items round-robin instead of one CTA per item
 *   "bin_min_rows"   auto rule: mode 2 when num_query >= this; 0 = never (default: measured slower than mode 1 on B200)
 *   "owned_max_taps" auto rule: mode 3 for bf16 problems with num_query*L*P*4 <= this * spatial_size (default 4) and a
 *                    grad_value of at least 64 MB (for fp32 both strategies write grad_value once and measure the same)
 * Returns the previous value, or -1 for an unknown key.  Results do not depend on the knobs beyond fp rounding. */
int msda_set_tuning(const char *key, int value);

/* out[b,q,m,:] = sum_{l,p} attn[b,q,m,l,p] * bilinear(value_l[b,:,m,:], loc[b,q,m,l,p]) */
int msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                 const void *sampling_loc, const void *attn_weight, void *output, const msda_dims *dims,
                 int dtype, unsigned flags, void *cuda_stream);

/* Bytes of device scratch msda_backward needs for this problem: 0 for f32/f64; N*S*M*D*4 for bf16 (fp32 image of
 * grad_value; 0 with MSDA_FLAG_ALIGNED16 when the owned backward applies); N*S*M*D*8 + 16 with
 * MSDA_FLAG_DETERMINISTIC.  The workspace must be 16-byte aligned. */
size_t msda_backward_workspace_bytes(const msda_dims *dims, int dtype, unsigned flags);

/* The strategy msda_backward will use for this problem when every tensor is 16-byte aligned (0 = invalid dims):
 *   1  row kernel: every tap is a global vector red into a zero-filled grad_value (or workspace);
 *   2  row kernel for the fine levels + msda_bwd_binned: coarse levels aggregated in shared memory, flushed once;
 *   3  row kernel for grad_sampling_loc / grad_attn_weight + msda_bwd_owned: every grad_value line is written exactly
 *      once by its owner -- no zero-fill (MSDA_FLAG_ZERO_GRAD_VALUE costs nothing), no workspace, no fold;
 *   4  msda_bwd_planes: the row algorithm in one kernel whose coarse-level taps are integer shared-memory atomics into
 *      per-(image, head) fixed-point planes flushed once per work item; the other levels leave as vector reds.
 * A pure function of (dims, dtype, flags) and the msda_set_tuning knobs. */
int msda_backward_strategy(const msda_dims *dims, int dtype, unsigned flags);

/* grad_value += scatter(w*attn*grad_out); grad_sampling_loc, grad_attn_weight = analytic gradients. */
int msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                  const void *sampling_loc, const void *attn_weight, const void *grad_output,
                  void *grad_value, void *grad_sampling_loc, void *grad_attn_weight, const msda_dims *dims,
                  int dtype, unsigned flags, void *workspace, size_t workspace_bytes, void *cuda_stream);

/*
 * Fused module path -- SURVEY.md section 8f-1.  Replaces, in MSDeformAttn.forward
 * (models/ops/modules/ms_deform_attn.py:96-111), the elementwise chain between the Linears and the core op:
 * softmax over the L*P logits, offsets / normaliser + reference points, and the padding-mask fill.
 *   sampling_offsets (N, Lq, M, L, P, 2) fp32  raw output of the sampling_offsets Linear
 *   attn_logits      (N, Lq, M, L*P)     fp32  raw output of the attention_weights Linear (pre-softmax)
 *   reference_points (N, Lq, L, ref_dim) fp32  ref_dim 2: (x, y);  ref_dim 4: (cx, cy, w, h)
 * msda_fused_backward writes grad_offsets / grad_logits (same shapes) and accumulates grad_value exactly like
 * msda_backward (same flags, same workspace rule: msda_backward_workspace_bytes).  The gradient of the reference
 * points is a reduction of grad_offsets the caller can do when it is needed.
 * msda_fused_supported() says whether a specialisation exists; callers fall back to msda_forward/backward otherwise.
 * msda_mask_rows zeroes, in place, the rows r of a (n_rows, row_bytes) array whose mask byte is non-zero
 * (value.masked_fill(padding_mask[..., None], 0) and its gradient, reference :96-97).
 */
int msda_fused_supported(const msda_dims *dims, int dtype, int ref_dim);
int msda_mask_rows(void *data, const unsigned char *mask, int64_t n_rows, int64_t row_bytes, void *cuda_stream);
int msda_fused_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_offsets, const void *attn_logits, const void *reference_points,
                       int ref_dim, void *output, const msda_dims *dims, int dtype, unsigned flags,
                       void *cuda_stream);
int msda_fused_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const void *sampling_offsets, const void *attn_logits, const void *reference_points,
                        int ref_dim, const void *grad_output, void *grad_value, void *grad_offsets,
                        void *grad_logits, const msda_dims *dims, int dtype, unsigned flags, void *workspace,
                        size_t workspace_bytes, void *cuda_stream);

/* Same, with the valid-ratio scaling of the decoder layer (models/detection/det_module.py:323-328) inside the kernels:
 * with valid_ratios (N, L, 2) fp32 [w-ratio, h-ratio] != NULL, reference_points is the UN-EXPANDED (N, Lq, ref_dim)
 * tensor and level l uses reference_points * valid_ratios[:, l] (both halves of a 4-d box are scaled);
 * valid_ratios == NULL is msda_fused_forward / msda_fused_backward. */
int msda_fused_forward_vr(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                          const void *sampling_offsets, const void *attn_logits, const void *reference_points,
                          const void *valid_ratios, int ref_dim, void *output, const msda_dims *dims, int dtype,
                          unsigned flags, void *cuda_stream);
int msda_fused_backward_vr(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                           const void *sampling_offsets, const void *attn_logits, const void *reference_points,
                           const void *valid_ratios, int ref_dim, const void *grad_output, void *grad_value,
                           void *grad_offsets, void *grad_logits, const msda_dims *dims, int dtype, unsigned flags,
                           void *workspace, size_t workspace_bytes, void *cuda_stream);

/*
 * Decoder-layer epilogue -- SURVEY.md section 8f-2.  Replaces, after each attention / FFN block of
 * DeformableTransformerDecoderLayer (models/detection/det_module.py:316-318, 331-333, 337-339),
 *     tgt = tgt + dropout(tgt2);  tgt = norm(tgt)          (nn.Dropout + add + nn.LayerNorm: 3 launches, 5 in backward)
 * with one kernel each way.  All tensors fp32, rows x channels contiguous, channels in {128, 256, 384, 512}:
 *   forward : y = LayerNorm(x + keep * keep_scale * z) * gamma + beta.   keep: rows x channels bytes (non-zero = kept),
 *             NULL = dropout off (then keep_scale should be 1).  h_saved / mean / rstd (rows x channels, rows, rows) are
 *             what backward needs; pass all three NULL for inference.
 *   backward: grad_x, grad_z (rows x channels), grad_gamma, grad_beta (channels; deterministic two-stage reduction
 *             through `workspace`, msda_add_dropout_ln_workspace_bytes bytes, 16-byte aligned).
 */
int msda_add_dropout_ln_supported(int64_t channels);
size_t msda_add_dropout_ln_workspace_bytes(int64_t rows, int64_t channels);
int msda_add_dropout_ln_forward(const void *x, const void *z, const unsigned char *keep, float keep_scale,
                                const void *gamma, const void *beta, float eps, void *y, void *h_saved, void *mean,
                                void *rstd, int64_t rows, int64_t channels, void *cuda_stream);
int msda_add_dropout_ln_backward(const void *grad_y, const void *h_saved, const void *mean, const void *rstd,
                                 const unsigned char *keep, float keep_scale, const void *gamma, void *grad_x,
                                 void *grad_z, void *grad_gamma, void *grad_beta, void *workspace,
                                 size_t workspace_bytes, int64_t rows, int64_t channels, void *cuda_stream);

/*
 * Level packing -- SURVEY.md section 8f-3.  Replaces the flatten/transpose/cat of prepare_od_inputs
 * (models/detection/det_module.py:146-155): `level_ptrs[l]` is the contiguous NCHW tensor (N, C, H_l, W_l) of level l,
 * `level_hw[l]` = H_l*W_l (host ints), `memory` is (N, sum_l H_l*W_l, C).  unpack = 0: levels -> memory;
 * unpack = 1: memory -> levels (the adjoint, i.e. the backward of the packing).  One launch, tiled transpose.
 */
int msda_pack_levels(void *const *level_ptrs, const int64_t *level_hw, int num_levels, int64_t batch, int64_t channels,
                     void *memory, int dtype, int unpack, void *cuda_stream);

/*
 * GroupNorm epilogue -> packed memory -- SURVEY.md section 8f-3, second half.  Detector.input_proj
 * (models/detection/detector.py:39-44, 64) ends in GroupNorm(num_groups, C) on each level's conv output, and
 * prepare_od_inputs then re-lays the normalised maps out as memory.  This call does both: `level_ptrs[l]` is the fp32 NCHW
 * CONV output of level l (pre-normalisation), gamma/beta the level's GroupNorm affine (fp32, C each); `memory`
 * (N, sum_l H_l*W_l, C) receives GroupNorm(x) in the op's layout as f32 or bf16 (out_dtype); `stats` (L, N, G, 2) fp32
 * receives mean / rstd (what a GroupNorm backward needs).  Two launches: statistics, then normalise + affine + transpose.
 */
int msda_pack_levels_groupnorm(void *const *level_ptrs, const int64_t *level_hw, int num_levels, int64_t batch,
                               int64_t channels, int num_groups, void *const *gamma_ptrs, void *const *beta_ptrs, float eps,
                               void *memory, int out_dtype, void *stats, void *cuda_stream);

/*
 * Measurement aid for bench.py: one launch of a microbenchmark with the kernels' access pattern and none of their
 * arithmetic -- uniform random 128-byte lines inside `scratch` (make it L2-resident, e.g. one image of value),
 * four lines per warp instruction.  which = 0: LDG.E.128 gather stream (ceiling of the forward's tap gather);
 * which = 1: REDG.E.ADD.F32x4 stream (ceiling of the backward's grad_value scatter; scratch is accumulated into).
 * *lines_out = lines moved by the launch; time it with CUDA events on the same stream.
 */
int msda_probe_ceiling(int which, void *scratch, size_t scratch_bytes, int64_t *lines_out, void *cuda_stream);

/*
 * Host-buffer path (what a caller without device tensors uses; bench.py's "e2e" leg).
 * A session owns device buffers and streams sized for `max_dims`; msda_host_forward_backward copies
 * the inputs host->device image-chunk by image-chunk, runs forward+backward, and copies the four
 * results back, overlapping copies with kernels.  Host buffers should be page-locked for full
 * PCIe bandwidth.  spatial_shapes / level_start_index are HOST pointers here.
 * msda_host_forward_backward returns after all results have landed in the host buffers.  msda_host_submit enqueues
 * the same work and returns at once, so consecutive calls (the six layers of a decoder, the next batch) keep the
 * copy / kernel / copy pipeline full across call boundaries; msda_host_wait blocks until everything submitted so far
 * has landed.  Host buffers handed to msda_host_submit must stay valid and untouched until msda_host_wait returns.
 */
typedef struct msda_host_session msda_host_session;

int msda_host_session_create(msda_host_session **session, const msda_dims *max_dims, int dtype, int device,
                             int images_per_chunk);
void msda_host_session_destroy(msda_host_session *session);
int msda_host_submit(msda_host_session *session, const void *value, const int64_t *spatial_shapes,
                     const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight,
                     const void *grad_output, void *output, void *grad_value, void *grad_sampling_loc,
                     void *grad_attn_weight, const msda_dims *dims, unsigned flags);
int msda_host_wait(msda_host_session *session);
int msda_host_forward_backward(msda_host_session *session, const void *value, const int64_t *spatial_shapes,
                               const int64_t *level_start_index, const void *sampling_loc,
                               const void *attn_weight, const void *grad_output, void *output,
                               void *grad_value, void *grad_sampling_loc, void *grad_attn_weight,
                               const msda_dims *dims, unsigned flags);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_H_ */
