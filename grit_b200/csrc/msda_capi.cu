// msda_capi.cu -- host side of the C ABI declared in include/msda.h.
//
// Replaces the reference's host launchers (models/ops/src/cuda/ms_deform_attn_cuda.cu:20-153 and
// ms_deform_im2col_cuda.cuh:923-1327): argument checks, kernel selection by (dtype, D, L, P),
// launch on the caller's stream, errors returned instead of printed.  No device allocation, no
// synchronisation, no host read of spatial_shapes.
#include "../../include/msda.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "msda_kernels_fused.cuh"
#include "msda_kernels_layer.cuh"
#include "msda_kernels_staged.cuh"
#include "msda_kernels_planes.cuh"

#include <atomic>
#include <type_traits>

namespace {

thread_local char tl_error[512] = "";
thread_local char tl_kernel[128] = "";
thread_local int64_t tl_launches = 0;

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return MSDA_OK;
    return fail(MSDA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

const char *dtype_name(int dtype)
{
    return dtype == MSDA_F32 ? "f32" : dtype == MSDA_F64 ? "f64" : dtype == MSDA_BF16 ? "bf16" : "?";
}

size_t dtype_size(int dtype) { return dtype == MSDA_F32 ? 4 : dtype == MSDA_F64 ? 8 : 2; }

int check_dims(const msda_dims *d, int dtype)
{
    if (!d) return fail(MSDA_ERR_INVALID_ARGUMENT, "dims is null");
    if (dtype != MSDA_F32 && dtype != MSDA_F64 && dtype != MSDA_BF16)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    if (d->batch < 0 || d->num_query < 0 || d->spatial_size < 0)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "negative batch/num_query/spatial_size");
    if (d->num_heads <= 0 || d->channels <= 0 || d->num_levels <= 0 || d->num_point <= 0)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "num_heads, channels, num_levels, num_point must be positive");
    if (d->num_levels > 1024 || d->num_point > (1 << 20) || d->channels > (1 << 24) || d->num_heads > (1 << 20))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "dimension out of range");
    // per-image element count must fit 31 bits (64-bit image bases are used on top of it)
    const int64_t per_image = d->spatial_size * d->num_heads * d->channels;
    if (d->spatial_size > 0 && per_image / d->spatial_size != d->num_heads * d->channels)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "per-image size overflows");
    if (d->batch * d->num_query * d->num_heads > ((int64_t)1 << 40))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "batch*num_query*num_heads too large");
    if (d->spatial_size >= ((int64_t)1 << 31) || d->num_query >= ((int64_t)1 << 31) || d->batch >= ((int64_t)1 << 31))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "batch, num_query and spatial_size must be below 2^31");
    return MSDA_OK;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kWarps = 8;

// Process-wide A/B knobs (msda_set_tuning).  Defaults are the measured best on B200 (profiles/); they exist for
// benchmarking and tests, results do not depend on them beyond fp rounding.
std::atomic<int> g_variant{0};        // forward: 0 = auto (default) | 5 = lean row kernel | 3 = persistent shared-memory-staged
std::atomic<int> g_v3_threads{1024};  // staged forward CTA size: 512, 768 or 1024
std::atomic<int> g_warps{4};          // row kernels, warps per CTA for D=32 L=P=4: 4 or 8
std::atomic<int> g_bf16_x4{1};        // row forward, bf16 D=32 L=P=4: 8-byte lane chunks (0 = 16-byte chunks, the A/B alternative)
std::atomic<int> g_hoist{0};          // row forward: issue all tap loads of a row before consuming any
std::atomic<int> g_bwd_mode{0};       // backward: 0 auto | 1 row kernel only | 2 row + binned coarse levels | 3 owned (sparse)
                                      //           | 4 planes (coarse levels in shared-memory int32 fixed point)
std::atomic<int> g_planes_rows{0};        // planes backward: query rows per work item (CTA); 0 = 256 for small CTAs, 1024 otherwise
std::atomic<int> g_planes_threads{768};  // planes backward CTA size: 256 (x4 per SM), 512, 768 or 1024 (one per SM)
std::atomic<int> g_planes_auto{1};        // auto: 0 = never choose the planes backward by default
std::atomic<int> g_planes_budget{1 << 30};  // planes backward: cap on the plane bytes (default: all the shared memory)
std::atomic<int> g_staged_rows{1024};     // staged forward: query rows per work item (CTA)
std::atomic<int> g_staged_persistent{0};  // staged forward A/B: 1 = one CTA per SM walking the items round-robin
std::atomic<int> g_staged_auto{0};        // auto: 1 = let the rule below choose the staged forward; 0 (default) = never:
                                          // since the row kernel's general path stopped zero-filling (1.31 -> 1.24 ms)
                                          // it beats the staged forward on every shape measured (0.59 vs 0.60 ms at
                                          // 384x640, 1.22 vs 1.38 ms at 800x1333; profiles/r02_fwd_general_path_ab.txt)
std::atomic<int> g_staged_min_rows{200};  // staged rule: D=32, coarse levels fit, num_heads*num_query / #SMs >= this
std::atomic<int> g_bin_min_rows{0};      // auto: binned coarse levels when num_query >= this; 0 = never (measured slower
                                         // than the plain row kernel on B200)
std::atomic<int> g_owned_max_taps{4};    // auto: owned when taps per value pixel (Lq*L*P*4 / S) <= this

struct Geometry {
    int64_t rows;
    unsigned grid;
};

int geometry(const msda_dims *d, Geometry *g)
{
    g->rows = d->batch * d->num_query * d->num_heads;
    const int64_t blocks = (g->rows + kWarps - 1) / kWarps;
    if (blocks > 0x7fffffffLL) return fail(MSDA_ERR_INVALID_ARGUMENT, "grid too large");
    g->grid = (unsigned)blocks;
    return MSDA_OK;
}

bool vec_eligible(const msda_dims *d, int dtype, unsigned flags)
{
    if (flags & MSDA_FLAG_FORCE_GENERIC) return false;
    if (dtype != MSDA_F32 && dtype != MSDA_BF16) return false;
    if (d->batch * d->num_query * d->num_heads >= ((int64_t)1 << 31)) return false;  // 32-bit row index
    if (d->num_heads * d->num_query >= ((int64_t)1 << 31)) return false;
    if (d->spatial_size >= ((int64_t)1 << 27)) return false;  // pixel index is packed as pix*16 + tap mask
    if (d->batch > 65535) return false;                        // gridDim.y
    return d->spatial_size * d->num_heads * d->channels < ((int64_t)1 << 31);
}

struct DeviceInfo {
    int sms = 0;
    int max_smem_optin = 0;
    int smem_per_sm = 0;
};

const DeviceInfo &device_info()
{
    thread_local DeviceInfo cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    DeviceInfo &d = cache[dev & 63];
    if (d.sms == 0) {
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&d.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    }
    return d;
}

// ---- specialisation tables -----------------------------------------------------------------------
#define MSDA_FOR_EACH_SPEC(X) \
    X(32, 4, 4)               \
    X(64, 4, 4)               \
    X(16, 4, 4)               \
    X(128, 4, 4)              \
    X(32, 4, 8)               \
    X(64, 4, 8)               \
    X(32, 1, 4)               \
    X(64, 1, 4)               \
    X(32, 1, 8)               \
    X(64, 1, 8)               \
    X(32, 3, 4)               \
    X(64, 3, 4)               \
    X(32, 5, 4)               \
    X(64, 5, 4)

// shapes with the staged forward, the binned / owned backward and the fused module kernels
#define MSDA_FOR_EACH_FLAGSHIP_SPEC(X) \
    X(32, 4, 4)                        \
    X(64, 4, 4)

template <typename T>
const char *tname();
template <>
const char *tname<float>() { return "f32"; }
template <>
const char *tname<__nv_bfloat16>() { return "bf16"; }

// ---- row kernels (msda_kernels_v5.cuh) ---------------------------------------------------------------
template <typename T, int DD, int LL, int PP, int W>
void fwd_v5_launch(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                   const void *attn, void *out, cudaStream_t st)
{
    const unsigned rpi = (unsigned)(d->num_query * d->num_heads);
    const dim3 grid((rpi + W - 1) / W, (unsigned)d->batch);
    constexpr bool kFlagship = DD == 32 && LL == 4 && PP == 4;  // the hoisted A/B variant exists for this shape only
    bool hoist = false;
    if constexpr (kFlagship) hoist = g_hoist.load() != 0;
    if constexpr (kFlagship) {
        if (hoist) {
            msda::msda_fwd_v5<T, DD, LL, PP, W, true><<<grid, W * 32, 0, st>>>(
                (const T *)value, shapes, lsi, (const float *)loc, (const float *)attn, (T *)out,
                (int)d->spatial_size, (int)d->num_heads, rpi);
        }
    }
    // bf16 at D=32: 8-byte lane chunks give the kernel the fp32 kernel's shape (8 lanes per tap, 4 taps per load instruction)
    bool x4 = false;
    if constexpr (kFlagship && W == 4 && sizeof(T) == 2) x4 = !hoist && g_bf16_x4.load() != 0;
    if constexpr (kFlagship && W == 4 && sizeof(T) == 2) {
        if (x4)
            msda::msda_fwd_v5<T, DD, LL, PP, W, false, msda::ChunkBf16x4><<<grid, W * 32, 0, st>>>(
                (const T *)value, shapes, lsi, (const float *)loc, (const float *)attn, (T *)out,
                (int)d->spatial_size, (int)d->num_heads, rpi);
    }
    if (!hoist && !x4)
        msda::msda_fwd_v5<T, DD, LL, PP, W, false><<<grid, W * 32, 0, st>>>(
            (const T *)value, shapes, lsi, (const float *)loc, (const float *)attn, (T *)out, (int)d->spatial_size,
            (int)d->num_heads, rpi);
    snprintf(tl_kernel, sizeof(tl_kernel), "fwd_v5<%s,D%d,L%d,P%d,w%d%s%s>", tname<T>(), DD, LL, PP, W,
             hoist ? ",hoisted" : "", x4 ? ",x4" : "");
}

template <typename T>
struct BwdChunk {
    using type = msda::Chunk<T>;
};
template <>
struct BwdChunk<__nv_bfloat16> {
    using type = msda::ChunkBf16x4;
};

// skip_budget > 0: levels whose fp32 plane fits that many bytes get their grad_value from another kernel
template <typename T, int DD, int LL, int PP, int W>
void bwd_v5_launch(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                   const void *attn, const void *gout, void *gv_acc, const float *det_scale, void *gloc, void *gattn,
                   int skip_budget, cudaStream_t st)
{
    const unsigned rpi = (unsigned)(d->num_query * d->num_heads);
    const dim3 grid((rpi + W - 1) / W, (unsigned)d->batch);
    using CH = typename BwdChunk<T>::type;
    constexpr bool kHasSkip = (DD == 32 || DD == 64) && LL == 4 && PP == 4 && W == 4;
    if (det_scale)
        msda::msda_bwd_v5<T, CH, msda::AccFix64, DD, LL, PP, W><<<grid, W * 32, 0, st>>>(
            (const T *)value, shapes, lsi, (const float *)loc, (const float *)attn, (const T *)gout,
            (unsigned long long *)gv_acc, det_scale, (float *)gloc, (float *)gattn, (int)d->spatial_size,
            (int)d->num_heads, rpi, 0);
    else if (skip_budget > 0) {
        if constexpr (kHasSkip)
            msda::msda_bwd_v5<T, CH, msda::AccF32, DD, LL, PP, W, true><<<grid, W * 32, 0, st>>>(
                (const T *)value, shapes, lsi, (const float *)loc, (const float *)attn, (const T *)gout,
                (float *)gv_acc, nullptr, (float *)gloc, (float *)gattn, (int)d->spatial_size, (int)d->num_heads, rpi,
                skip_budget);
    } else
        msda::msda_bwd_v5<T, CH, msda::AccF32, DD, LL, PP, W><<<grid, W * 32, 0, st>>>(
            (const T *)value, shapes, lsi, (const float *)loc, (const float *)attn, (const T *)gout, (float *)gv_acc,
            nullptr, (float *)gloc, (float *)gattn, (int)d->spatial_size, (int)d->num_heads, rpi, 0);
    snprintf(tl_kernel, sizeof(tl_kernel), "bwd_v5<%s,D%d,L%d,P%d,w%d%s>", tname<T>(), DD, LL, PP, W,
             det_scale ? ",deterministic" : "");
}

template <int DD, int LL, int PP, int E>
constexpr bool v5_ok()
{
    constexpr int LPT = DD / E, G = 32 / LPT, LP = LL * PP, PPG = LP / G;
    return DD % E == 0 && 32 % LPT == 0 && LP % G == 0 && LP <= 32 && PPG >= 1 && PPG <= LPT;
}

template <typename T>
bool launch_fwd_v5(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                   const void *attn, void *out, cudaStream_t st)
{
    constexpr int E = msda::Chunk<T>::E;
    const bool w8 = g_warps.load() >= 8;
#define X(DD, LL, PP)                                                                                  \
    if constexpr (v5_ok<DD, LL, PP, E>()) {                                                            \
        if (d->channels == (DD) && d->num_levels == (LL) && d->num_point == (PP)) {                   \
            if constexpr ((DD) == 32 && (LL) == 4 && (PP) == 4) {                                      \
                if (w8) {                                                                              \
                    fwd_v5_launch<T, DD, LL, PP, 8>(d, value, shapes, lsi, loc, attn, out, st);        \
                    return true;                                                                       \
                }                                                                                      \
            }                                                                                          \
            fwd_v5_launch<T, DD, LL, PP, 4>(d, value, shapes, lsi, loc, attn, out, st);                \
            return true;                                                                               \
        }                                                                                              \
    }
    MSDA_FOR_EACH_SPEC(X)
#undef X
    return false;
}

template <typename T>
bool launch_bwd_v5(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                   const void *attn, const void *gout, void *gv_acc, const float *det_scale, void *gloc, void *gattn,
                   int skip_budget, cudaStream_t st)
{
    constexpr int E = BwdChunk<T>::type::E;
    const bool w8 = g_warps.load() >= 8 && skip_budget == 0;
#define X(DD, LL, PP)                                                                                             \
    if constexpr (v5_ok<DD, LL, PP, E>()) {                                                                       \
        if (d->channels == (DD) && d->num_levels == (LL) && d->num_point == (PP)) {                              \
            if constexpr ((DD) == 32 && (LL) == 4 && (PP) == 4) {                                                 \
                if (w8) {                                                                                         \
                    bwd_v5_launch<T, DD, LL, PP, 8>(d, value, shapes, lsi, loc, attn, gout, gv_acc, det_scale,    \
                                                    gloc, gattn, 0, st);                                          \
                    return true;                                                                                  \
                }                                                                                                 \
            }                                                                                                     \
            bwd_v5_launch<T, DD, LL, PP, 4>(d, value, shapes, lsi, loc, attn, gout, gv_acc, det_scale, gloc,      \
                                            gattn, skip_budget, st);                                              \
            return true;                                                                                          \
        }                                                                                                         \
    }
    MSDA_FOR_EACH_SPEC(X)
#undef X
    return false;
}

// ---- persistent shared-memory-staged forward (msda_kernels_staged.cuh) ---------------------------------
template <typename K>
int optin_smem(K kernel, int bytes)
{
    return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes),
                      "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
}

template <typename T, int DD, int LL, int PP, int TH>
int fwd_staged_launch(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                      const void *attn, void *out, cudaStream_t st)
{
    auto kernel = msda::msda_fwd_v3<T, DD, LL, PP, TH>;
    const int smem = device_info().max_smem_optin - 1024;  // leave room for the kernel's static shared memory
    if (smem <= 0) return fail(MSDA_ERR_CUDA, "device reports no opt-in shared memory");
    if (int rc = optin_smem(kernel, smem)) return rc;
    // items = batch x chunks x heads, ~1024 rows each: small enough that the block scheduler's granularity costs little
    // when it balances SMs of unequal speed (a grid of only ~4 items per SM loses up to a fifth to wave quantisation),
    // large enough that staging the planes (one 169 KB read of the coarse levels per item) stays at a few percent
    int rows_per_item = g_staged_rows.load();
    if (rows_per_item < 64) rows_per_item = 64;
    int64_t chunks = (d->num_query + rows_per_item - 1) / rows_per_item;
    if (chunks < 1) chunks = 1;
    const int64_t items = d->batch * chunks * d->num_heads;
    if (items > 0x7fffffffLL) return fail(MSDA_ERR_INVALID_ARGUMENT, "grid too large");
    const int64_t grid = g_staged_persistent.load() && items > device_info().sms ? device_info().sms : items;
    kernel<<<(unsigned)grid, TH, smem, st>>>((const T *)value, shapes, lsi, (const float *)loc, (const float *)attn,
                                             (T *)out, (int)d->batch, (int)d->spatial_size, (int)d->num_heads,
                                             (int)d->num_query, smem / (int)sizeof(T), (int)chunks);
    snprintf(tl_kernel, sizeof(tl_kernel), "fwd_staged<%s,D%d,L%d,P%d,t%d>", tname<T>(), DD, LL, PP, TH);
    return MSDA_OK;
}

// returns -1 when no staged specialisation matches, else a status code
template <typename T>
int launch_fwd_staged(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                      const void *attn, void *out, cudaStream_t st)
{
    const int th = g_v3_threads.load();
#define X(DD, LL, PP)                                                                                         \
    if (d->channels == (DD) && d->num_levels == (LL) && d->num_point == (PP))                                \
        return th >= 1024  ? fwd_staged_launch<T, DD, LL, PP, 1024>(d, value, shapes, lsi, loc, attn, out, st) \
               : th >= 768 ? fwd_staged_launch<T, DD, LL, PP, 768>(d, value, shapes, lsi, loc, attn, out, st)  \
                           : fwd_staged_launch<T, DD, LL, PP, 512>(d, value, shapes, lsi, loc, attn, out, st);
    MSDA_FOR_EACH_FLAGSHIP_SPEC(X)
#undef X
    return -1;
}

// ---- grad_value by on-SM aggregation (msda_kernels_binned.cuh) -------------------------------------------
enum { BWD_ROW = 1, BWD_BINNED = 2, BWD_OWNED = 3, BWD_PLANES = 4 };

bool flagship_spec(const msda_dims *d)
{
#define X(DD, LL, PP) \
    if (d->channels == (DD) && d->num_levels == (LL) && d->num_point == (PP)) return true;
    MSDA_FOR_EACH_FLAGSHIP_SPEC(X)
#undef X
    return false;
}

struct BinnedPlan {
    int smem_bytes, region_bytes, acc_budget, bins_cap;
};

// Shared-memory split of msda_bwd_binned for channel width D and P points (see the layout comment at the kernel).
bool binned_plan(int D, int P, BinnedPlan *bp)
{
    const int total = (device_info().max_smem_optin - 1024) & ~15;
    const int min_tile = (msda::kBinThreads / P) * (D * 4 + P * 16);
    // region + 4 * (2 * bins_cap + 4) <= total with bins_cap = (region - min_tile) / (2 D) + 2
    long long region = ((long long)(total - 32) * D + 4LL * min_tile) / (D + 4);
    region &= ~15LL;
    for (;; region -= 16) {
        if (region <= min_tile) return false;
        const long long budget = region - min_tile;
        const long long bins = budget / (2 * D) + 2;
        if (region + 4 * (2 * bins + 4) <= total) {
            bp->smem_bytes = total, bp->region_bytes = (int)region, bp->acc_budget = (int)budget, bp->bins_cap = (int)bins;
            return true;
        }
    }
}

struct OwnedPlan {
    int smem_bytes, chunks, chunk_pixels, tq;
};

bool owned_plan(const msda_dims *d, OwnedPlan *op)
{
    const int total = (device_info().max_smem_optin - 1024) & ~15;
    const int64_t S = d->spatial_size, Lq = d->num_query;
    const int64_t per_row = d->channels * 4 + d->num_levels * d->num_point * 32;
    int64_t chunks = (8LL * device_info().sms + d->batch * d->num_heads - 1) / (d->batch * d->num_heads);
    if (chunks < 1) chunks = 1;
    if (chunks > S) chunks = S > 0 ? S : 1;
    for (;; chunks *= 2) {
        const int64_t cp = (S + chunks - 1) / chunks;
        const int64_t cur_bytes = ((cp + 1) * 4 + 15) & ~15LL;
        const int64_t rows_fit = (total - cur_bytes) / per_row;
        const int64_t want = Lq < 32 ? (Lq > 0 ? Lq : 1) : 32;
        if (total > cur_bytes && rows_fit >= want) {
            op->smem_bytes = total;
            op->chunks = (int)chunks, op->chunk_pixels = (int)cp;
            op->tq = (int)(rows_fit < Lq ? rows_fit : (Lq > 0 ? Lq : 1));
            return true;
        }
        if (cp <= 1) return false;
    }
}

// The backward strategy for this problem.  Pure function of (dims, dtype, flags, knobs): msda_backward_workspace_bytes
// relies on that.
int choose_bwd_mode(const msda_dims *d, int dtype, unsigned flags)
{
    if ((flags & MSDA_FLAG_DETERMINISTIC) || !vec_eligible(d, dtype, flags) || !flagship_spec(d)) return BWD_ROW;
    if (d->batch * d->num_query * d->num_heads == 0 || d->spatial_size == 0) return BWD_ROW;
    const int forced = g_bwd_mode.load();
    if (forced == BWD_ROW) return BWD_ROW;
    BinnedPlan bp;
    OwnedPlan op;
    const bool can_bin = binned_plan((int)d->channels, (int)d->num_point, &bp);
    const bool can_own = owned_plan(d, &op);
    if (forced == BWD_BINNED) return can_bin ? BWD_BINNED : BWD_ROW;
    if (forced == BWD_OWNED) return can_own ? BWD_OWNED : BWD_ROW;
    if (forced == BWD_PLANES) return BWD_PLANES;
    // owned pays where the row path's zero-fill / workspace / fold traffic dominates: sparse problems in bf16 (fp32
    // image of grad_value zero-filled, accumulated, read back, folded) that are large enough to be bandwidth- rather
    // than launch-bound.  For fp32 both strategies write grad_value once and measure the same.
    const int64_t taps = d->num_query * d->num_levels * d->num_point * 4;
    const int64_t value_bytes = d->batch * d->spatial_size * d->num_heads * d->channels * 2;
    if (can_own && dtype == MSDA_BF16 && value_bytes >= ((int64_t)64 << 20) &&
        taps <= (int64_t)g_owned_max_taps.load() * d->spatial_size)
        return BWD_OWNED;
    const int bin_min = g_bin_min_rows.load();
    if (can_bin && bin_min > 0 && d->num_levels >= 2 && d->num_query >= bin_min) return BWD_BINNED;
    // planes (coarse levels' gradient accumulated in shared memory as int32 fixed point) beats the row kernel on every
    // dense D=32 shape INSIDE a long step, where the GPU runs at its power cap (bench.py, six layers back to back;
    // profiles/r02_planes_bench_ab.txt): 800x1333 fp32 step 30.2 -> 29.4 ms (uniform) / 31.5 -> 29.0 ms (detector-like),
    // bf16 60.9 -> 56.2 ms, 384x640 13.7 -> 12.3 ms, with one 768-thread CTA per SM.  Four 256-thread CTAs per SM are
    // faster when the kernel is timed alone (3.33 vs 3.51 ms) but draw more power: the clocks of the whole step drop
    // (1815 vs 1940 MHz) and the step is slower (29.9 ms).  Dense = more taps than the owned strategy's domain, and
    // enough rows to fill the machine.
    if (g_planes_auto.load() && d->channels == 32 && taps > (int64_t)g_owned_max_taps.load() * d->spatial_size &&
        d->num_heads * d->num_query >= (int64_t)g_staged_min_rows.load() * device_info().sms)
        return BWD_PLANES;
    return BWD_ROW;
}

template <typename T>
int launch_bwd_binned(const msda_dims *d, const BinnedPlan &bp, const int64_t *shapes, const int64_t *lsi,
                      const void *loc, const void *attn, const void *gout, float *gv_acc, cudaStream_t st)
{
#define X(DD, LL, PP)                                                                                              \
    if (d->channels == (DD) && d->num_levels == (LL) && d->num_point == (PP)) {                                   \
        auto kernel = msda::msda_bwd_binned<T, DD, LL, PP>;                                                        \
        if (int rc = optin_smem(kernel, bp.smem_bytes)) return rc;                                                 \
        kernel<<<device_info().sms, msda::kBinThreads, bp.smem_bytes, st>>>(                                       \
            shapes, lsi, (const float *)loc, (const float *)attn, (const T *)gout, gv_acc, (int)d->batch,          \
            (int)d->spatial_size, (int)d->num_heads, (int)d->num_query, bp.acc_budget, bp.region_bytes, bp.bins_cap); \
        return MSDA_OK;                                                                                            \
    }
    MSDA_FOR_EACH_FLAGSHIP_SPEC(X)
#undef X
    return fail(MSDA_ERR_UNSUPPORTED, "no binned backward for this shape");
}

template <typename T>
int launch_bwd_owned(const msda_dims *d, const OwnedPlan &op, const int64_t *shapes, const int64_t *lsi, const void *loc,
                     const void *attn, const void *gout, void *grad_value, int accumulate, cudaStream_t st)
{
    const int64_t items = d->batch * d->num_heads * op.chunks;
    const int64_t cap = (int64_t)device_info().sms;
    const unsigned grid = (unsigned)(items < cap ? items : cap);
#define X(DD, LL, PP)                                                                                              \
    if (d->channels == (DD) && d->num_levels == (LL) && d->num_point == (PP)) {                                   \
        auto kernel = msda::msda_bwd_owned<T, DD, LL, PP>;                                                         \
        if (int rc = optin_smem(kernel, op.smem_bytes)) return rc;                                                 \
        kernel<<<grid, msda::kBinThreads, op.smem_bytes, st>>>(                                                    \
            shapes, lsi, (const float *)loc, (const float *)attn, (const T *)gout, (T *)grad_value, (int)d->batch, \
            (int)d->spatial_size, (int)d->num_heads, (int)d->num_query, op.chunks, op.chunk_pixels, op.tq,         \
            accumulate);                                                                                           \
        return MSDA_OK;                                                                                            \
    }
    MSDA_FOR_EACH_FLAGSHIP_SPEC(X)
#undef X
    return fail(MSDA_ERR_UNSUPPORTED, "no owned backward for this shape");
}

// ---- coarse levels in shared-memory fixed-point planes (msda_kernels_planes.cuh) ------------------------------------
// Query chunks per (image, head) of the planes backward for a CTA of TH threads.
int64_t planes_chunks(const msda_dims *d, int TH)
{
    int rows_per_item = g_planes_rows.load();
    if (rows_per_item <= 0) rows_per_item = TH <= 256 ? 256 : 1024;  // measured best for each CTA size
    if (rows_per_item < 64) rows_per_item = 64;
    int64_t chunks = (d->num_query + rows_per_item - 1) / rows_per_item;
    if (chunks < 1) chunks = 1;
    // One CTA per SM and items of ~130 us: when an image's item count is not a multiple of the SM count, the tail of
    // every image runs beside the head of the next one and TWO images' value / grad_value maps compete for the L2 for most
    // of the launch (800x1333: 176 items per image on 148 SMs -> 4.2 GB of DRAM traffic instead of 2.7 GB, 3.58 vs 3.46 ms).
    // So the chunk count is moved, within [0.55, 1.8] x the target, to the value that fills whole waves best (37 chunks
    // x 8 heads = 2 x 148 there).  Only when an image has at least half a wave of items; the knob "planes_rows" > 0 is
    // taken literally.
    if (TH > 256 && g_planes_rows.load() <= 0) {
        const int64_t slots = device_info().sms, heads = d->num_heads;
        if (chunks * heads * 2 >= slots) {
            int64_t best = chunks;
            double best_fill = 0.0;
            const int64_t lo = (chunks * 55 + 99) / 100, hi = chunks * 18 / 10;
            for (int64_t c = lo < 1 ? 1 : lo; c <= hi && c <= d->num_query; ++c) {
                const int64_t n = c * heads, waves = (n + slots - 1) / slots;
                const double fill = (double)n / (double)(waves * slots);
                const bool closer = (c > chunks ? c - chunks : chunks - c) < (best > chunks ? best - chunks : chunks - best);
                if (fill > best_fill + 1e-9 || (fill > best_fill - 1e-9 && closer)) best = c, best_fill = fill;
            }
            chunks = best;
        }
    }
    return chunks;
}

template <typename T, int DD, int LL, int PP, int TH>
int bwd_planes_launch(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                      const void *attn, const void *gout, float *gv_acc, void *gloc, void *gattn, cudaStream_t st)
{
    using CH = typename BwdChunk<T>::type;
    auto kernel = msda::msda_bwd_planes<T, CH, DD, LL, PP, TH>;
    int smem = (device_info().max_smem_optin - 2048 - (TH / 32) * DD * 4) & ~15;  // minus the static shared memory
    if (smem <= 0) return fail(MSDA_ERR_CUDA, "device reports no opt-in shared memory");
    const int cap = g_planes_budget.load();
    // Small CTAs share the SM (1024 / TH of them, the occupancy of the row kernel): each takes its share of the SM's shared
    // memory -- 55 KB at TH = 256 -- and the device-side plan puts the smallest levels that fit there (800x1333: level 3,
    // a quarter of the taps; 384x640: levels 2 + 3, half of them).  One big CTA per SM holds more levels but runs at the
    // mercy of instruction latency.
    if (TH <= 256) {
        const int per_cta = device_info().smem_per_sm / (1024 / TH) - 1024 /* reserved per CTA */ - (TH / 32) * DD * 4 - 256;
        if (per_cta < smem) smem = per_cta & ~15;
        // Ask only for what the planes will need: what a CTA does not take stays L1 (800x1333: 3.31 ms with 36 KB per
        // CTA, 3.52 ms with the full 55 KB).  The host does not read spatial_shapes, so the need is estimated from S for
        // the usual 4:1 pyramid -- the k smallest of four levels hold S/85, S/17, S/4 pixels -- plus 15 %; a pyramid that
        // needs more than the estimate simply keeps fewer levels on chip (the device-side plan decides).
        const double px1 = (double)d->spatial_size / 85.0;
        int64_t want = 0;
        for (double mult : {1.0, 5.0, 21.0}) {
            const int64_t bytes = (int64_t)(px1 * mult * 1.15 + 8.0) * DD * 4;
            if (bytes <= smem) want = bytes;
        }
        if (want > 0 && want < smem) smem = (int)((want + 15) & ~15);
        if (cap >= 0 && cap < smem) smem = (cap + 15) & ~15;
    }
    if (int rc = optin_smem(kernel, smem)) return rc;
    const int64_t chunks = planes_chunks(d, TH);
    const int64_t items = d->batch * chunks * d->num_heads;
    if (items > 0x7fffffffLL) return fail(MSDA_ERR_INVALID_ARGUMENT, "grid too large");
    kernel<<<(unsigned)items, TH, smem, st>>>((const T *)value, shapes, lsi, (const float *)loc, (const float *)attn,
                                              (const T *)gout, gv_acc, (float *)gloc, (float *)gattn, (int)d->batch,
                                              (int)d->spatial_size, (int)d->num_heads, (int)d->num_query,
                                              (cap >= 0 && cap < smem ? cap : smem) / 4, (int)chunks,
                                              (const float *)nullptr, (const float *)nullptr);
    snprintf(tl_kernel, sizeof(tl_kernel), "bwd_planes<%s,D%d,L%d,P%d,t%d>", tname<T>(), DD, LL, PP, TH);
    return MSDA_OK;
}

// returns -1 when no specialisation matches, else a status code
template <typename T>
int launch_bwd_planes(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                      const void *attn, const void *gout, float *gv_acc, void *gloc, void *gattn, cudaStream_t st)
{
    const int th = g_planes_threads.load();
#define X(DD, LL, PP)                                                                                                  \
    if (d->channels == (DD) && d->num_levels == (LL) && d->num_point == (PP))                                         \
        return th >= 1024                                                                                              \
                   ? bwd_planes_launch<T, DD, LL, PP, 1024>(d, value, shapes, lsi, loc, attn, gout, gv_acc, gloc, gattn, st) \
               : th >= 896                                                                                             \
                   ? bwd_planes_launch<T, DD, LL, PP, 896>(d, value, shapes, lsi, loc, attn, gout, gv_acc, gloc, gattn, st)  \
               : th >= 768                                                                                             \
                   ? bwd_planes_launch<T, DD, LL, PP, 768>(d, value, shapes, lsi, loc, attn, gout, gv_acc, gloc, gattn, st)  \
               : th >= 512                                                                                             \
                   ? bwd_planes_launch<T, DD, LL, PP, 512>(d, value, shapes, lsi, loc, attn, gout, gv_acc, gloc, gattn, st)  \
                   : bwd_planes_launch<T, DD, LL, PP, 256>(d, value, shapes, lsi, loc, attn, gout, gv_acc, gloc, gattn, st);
    MSDA_FOR_EACH_FLAGSHIP_SPEC(X)
#undef X
    return -1;
}

// ---- generic kernels -----------------------------------------------------------------------------------------------
template <typename T, typename C>
void launch_fwd_generic(const msda_dims *d, const Geometry &g, const void *value, const int64_t *shapes,
                        const int64_t *lsi, const void *loc, const void *attn, void *out, cudaStream_t st,
                        const char *name)
{
    const size_t smem = 3 * sizeof(int) * (size_t)d->num_levels;
    msda::msda_fwd_generic<T, C, kWarps><<<g.grid, kWarps * 32, smem, st>>>(
        (const T *)value, shapes, lsi, (const C *)loc, (const C *)attn, (T *)out, d->spatial_size,
        (int)d->num_heads, (int)d->channels, (int)d->num_levels, d->num_query, (int)d->num_point, g.rows);
    snprintf(tl_kernel, sizeof(tl_kernel), "fwd_generic<%s>", name);
}

template <typename T, typename C>
void launch_bwd_generic(const msda_dims *d, const Geometry &g, const void *value, const int64_t *shapes,
                        const int64_t *lsi, const void *loc, const void *attn, const void *gout, void *gv_acc,
                        const float *det_scale, void *gloc, void *gattn, cudaStream_t st, const char *name)
{
    const size_t smem = 3 * sizeof(int) * (size_t)d->num_levels;
    if constexpr (std::is_same<C, float>::value) {
        if (det_scale) {
            msda::msda_bwd_generic<T, C, unsigned long long, kWarps><<<g.grid, kWarps * 32, smem, st>>>(
                (const T *)value, shapes, lsi, (const C *)loc, (const C *)attn, (const T *)gout,
                (unsigned long long *)gv_acc, det_scale, (C *)gloc, (C *)gattn, d->spatial_size, (int)d->num_heads,
                (int)d->channels, (int)d->num_levels, d->num_query, (int)d->num_point, g.rows);
            snprintf(tl_kernel, sizeof(tl_kernel), "bwd_generic<%s,deterministic>", name);
            return;
        }
    }
    msda::msda_bwd_generic<T, C, C, kWarps><<<g.grid, kWarps * 32, smem, st>>>(
        (const T *)value, shapes, lsi, (const C *)loc, (const C *)attn, (const T *)gout, (C *)gv_acc, nullptr,
        (C *)gloc, (C *)gattn, d->spatial_size, (int)d->num_heads, (int)d->channels, (int)d->num_levels, d->num_query,
        (int)d->num_point, g.rows);
    snprintf(tl_kernel, sizeof(tl_kernel), "bwd_generic<%s>", name);
}

}  // namespace

extern "C" {

int msda_abi_version(void) { return MSDA_ABI_VERSION; }

const char *msda_last_error(void) { return tl_error; }

const char *msda_last_kernel(void) { return tl_kernel; }

int msda_set_tuning(const char *key, int value)
{
    std::atomic<int> *knob = nullptr;
    if (key && !strcmp(key, "variant")) knob = &g_variant;
    if (key && !strcmp(key, "warps")) knob = &g_warps;
    if (key && !strcmp(key, "v3_threads")) knob = &g_v3_threads;
    if (key && !strcmp(key, "hoist")) knob = &g_hoist;
    if (key && !strcmp(key, "bf16_x4")) knob = &g_bf16_x4;
    if (key && !strcmp(key, "bwd_mode")) knob = &g_bwd_mode;
    if (key && !strcmp(key, "bin_min_rows")) knob = &g_bin_min_rows;
    if (key && !strcmp(key, "staged_min_rows")) knob = &g_staged_min_rows;
    if (key && !strcmp(key, "staged_auto")) knob = &g_staged_auto;
    if (key && !strcmp(key, "staged_rows")) knob = &g_staged_rows;
    if (key && !strcmp(key, "staged_persistent")) knob = &g_staged_persistent;
    if (key && !strcmp(key, "owned_max_taps")) knob = &g_owned_max_taps;
    if (key && !strcmp(key, "planes_rows")) knob = &g_planes_rows;
    if (key && !strcmp(key, "planes_threads")) knob = &g_planes_threads;
    if (key && !strcmp(key, "planes_budget")) knob = &g_planes_budget;
    if (key && !strcmp(key, "planes_auto")) knob = &g_planes_auto;
    if (!knob) return -1;
    return knob->exchange(value);
}

int64_t msda_launch_count(int reset)
{
    const int64_t n = tl_launches;
    if (reset) tl_launches = 0;
    return n;
}

int msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                 const void *sampling_loc, const void *attn_weight, void *output, const msda_dims *dims, int dtype,
                 unsigned flags, void *cuda_stream)
{
    tl_error[0] = 0;
    if (int rc = check_dims(dims, dtype)) return rc;
    Geometry g;
    if (int rc = geometry(dims, &g)) return rc;
    if (g.rows == 0) return MSDA_OK;  // nothing to write: output has zero elements
    if (!value && dims->spatial_size > 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "value is null");
    if (!spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    cudaStream_t st = (cudaStream_t)cuda_stream;

    bool done = false;
    if (vec_eligible(dims, dtype, flags) && aligned16(value) && aligned16(output) &&
        (reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0) {
        // The staged forward paid off when at least three of the four levels of a head fit in shared memory, i.e. >= 3/4
        // of the taps come from the SM (profiles/r02_staged_ab.txt: 0.60 vs 0.62 ms at 384x640 fp32) -- until the row
        // kernel's general path lost its zero-fill (0.59 ms there now), so the rule is opt-in ("staged_auto").  The host
        // does not read spatial_shapes, so the rule assumes the usual 4:1 pyramid: all levels but the finest hold ~S/4
        // pixels.
        const int variant = g_variant.load();
        const int64_t coarse_bytes = dims->spatial_size / 4 * dims->channels * (int64_t)dtype_size(dtype);
        const bool staged_auto = variant == 0 && g_staged_auto.load() && dims->channels == 32 && dims->num_levels >= 3 &&
                                 coarse_bytes <= device_info().max_smem_optin - 2048 &&
                                 dims->num_heads * dims->num_query >= (int64_t)g_staged_min_rows.load() * device_info().sms;
        if (variant == 3 || staged_auto) {
            const int rc = dtype == MSDA_F32
                               ? launch_fwd_staged<float>(dims, value, spatial_shapes, level_start_index, sampling_loc,
                                                          attn_weight, output, st)
                               : launch_fwd_staged<__nv_bfloat16>(dims, value, spatial_shapes, level_start_index,
                                                                  sampling_loc, attn_weight, output, st);
            if (rc > 0) return rc;
            done = rc == 0;
        }
        if (!done)
            done = dtype == MSDA_F32 ? launch_fwd_v5<float>(dims, value, spatial_shapes, level_start_index,
                                                           sampling_loc, attn_weight, output, st)
                                     : launch_fwd_v5<__nv_bfloat16>(dims, value, spatial_shapes, level_start_index,
                                                                   sampling_loc, attn_weight, output, st);
    }
    if (!done) {
        if (dtype == MSDA_F32)
            launch_fwd_generic<float, float>(dims, g, value, spatial_shapes, level_start_index, sampling_loc,
                                             attn_weight, output, st, "f32");
        else if (dtype == MSDA_F64)
            launch_fwd_generic<double, double>(dims, g, value, spatial_shapes, level_start_index, sampling_loc,
                                               attn_weight, output, st, "f64");
        else
            launch_fwd_generic<__nv_bfloat16, float>(dims, g, value, spatial_shapes, level_start_index, sampling_loc,
                                                     attn_weight, output, st, "bf16");
    }
    ++tl_launches;
    return check_cuda(cudaPeekAtLastError(), "msda_forward launch");
}

int msda_backward_strategy(const msda_dims *dims, int dtype, unsigned flags)
{
    if (check_dims(dims, dtype) != MSDA_OK) return 0;
    return choose_bwd_mode(dims, dtype, flags);
}

size_t msda_backward_workspace_bytes(const msda_dims *dims, int dtype, unsigned flags)
{
    if (!dims) return 0;
    const size_t n_value = (size_t)(dims->batch * dims->spatial_size * dims->num_heads * dims->channels);
    if ((flags & MSDA_FLAG_DETERMINISTIC) && dtype != MSDA_F64)
        return n_value * sizeof(long long) + 16;  // int64 fixed-point accumulators + {max|attn|, max|g|, 2^k, 2^-k}
    if (dtype == MSDA_BF16) {
        // the owned backward writes bf16 grad_value directly; it needs 16-byte aligned tensors, which the caller vouches
        // for with MSDA_FLAG_ALIGNED16 (otherwise the answer stays conservative)
        if ((flags & MSDA_FLAG_ALIGNED16) && check_dims(dims, dtype) == MSDA_OK &&
            choose_bwd_mode(dims, dtype, flags) == BWD_OWNED)
            return 0;
        return n_value * sizeof(float);  // fp32 accumulation image of grad_value
    }
    return 0;
}

}  // extern "C" (helpers below need internal linkage)

namespace {

// grad_value accumulation target: grad_value itself (f32/f64), or the caller's workspace (bf16: fp32 image;
// deterministic: int64 fixed point + scale tail) that a fold kernel turns into grad_value afterwards.
struct AccPlan {
    void *gv_acc = nullptr;
    const float *det_scale = nullptr;
    unsigned *tail = nullptr;
    bool via_workspace = false;
    bool det = false;
};

// attn == nullptr means "attention weights are a softmax output, bounded by 1" (fused path)
int acc_begin(const msda_dims *dims, const Geometry &g, int dtype, unsigned flags, void *grad_value, void *workspace,
              size_t workspace_bytes, const void *attn, const void *grad_output, cudaStream_t st, AccPlan *plan)
{
    const int64_t n_value = dims->batch * dims->spatial_size * dims->num_heads * dims->channels;
    plan->det = (flags & MSDA_FLAG_DETERMINISTIC) != 0;
    plan->via_workspace = plan->det || dtype == MSDA_BF16;
    plan->gv_acc = grad_value;
    if (!plan->via_workspace) return MSDA_OK;
    const size_t need = msda_backward_workspace_bytes(dims, dtype, flags);
    if (!workspace || workspace_bytes < need)
        return fail(MSDA_ERR_WORKSPACE, "this backward needs a %zu-byte workspace, got %zu", need, workspace_bytes);
    if (!aligned16(workspace)) return fail(MSDA_ERR_WORKSPACE, "workspace must be 16-byte aligned");
    if (dtype == MSDA_BF16 && !plan->det && !aligned16(grad_value))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "bf16 grad_value must be 16-byte aligned (vector fold)");
    if (int rc = check_cuda(cudaMemsetAsync(workspace, 0, need, st), "memset workspace")) return rc;
    plan->gv_acc = workspace;
    if (plan->det) {
        plan->tail = reinterpret_cast<unsigned *>(static_cast<char *>(workspace) + (size_t)n_value * sizeof(long long));
        plan->det_scale = reinterpret_cast<const float *>(plan->tail) + 2;
        const int64_t n_attn = attn ? g.rows * dims->num_levels * dims->num_point : 0;
        const int64_t n_gout = g.rows * dims->channels;
        const int blocks = device_info().sms * 8;
        if (dtype == MSDA_F32)
            msda::msda_det_absmax<float><<<blocks, 256, 0, st>>>((const float *)attn, n_attn,
                                                                 (const float *)grad_output, n_gout, plan->tail);
        else
            msda::msda_det_absmax<__nv_bfloat16><<<blocks, 256, 0, st>>>(
                (const float *)attn, n_attn, (const __nv_bfloat16 *)grad_output, n_gout, plan->tail);
        // one element of grad_value receives at most one tap of each (query, level, point) of its image and head
        const double worst = (double)dims->num_query * (double)dims->num_levels * (double)dims->num_point;
        msda::msda_det_scale<<<1, 1, 0, st>>>(plan->tail, worst, attn ? 0.f : 1.f);
        tl_launches += 2;
        if (int rc = check_cuda(cudaPeekAtLastError(), "deterministic pre-pass launch")) return rc;
    }
    return MSDA_OK;
}

int acc_end(const msda_dims *dims, int dtype, unsigned flags, void *grad_value, void *workspace, const AccPlan &plan,
            cudaStream_t st)
{
    if (!plan.via_workspace) return MSDA_OK;
    const int64_t n_value = dims->batch * dims->spatial_size * dims->num_heads * dims->channels;
    const int threads = 256;
    const int accumulate = (flags & MSDA_FLAG_ZERO_GRAD_VALUE) ? 0 : 1;
    if (plan.det) {
        const int blocks = device_info().sms * 16;
        if (dtype == MSDA_F32)
            msda::msda_det_fold<float><<<blocks, threads, 0, st>>>((const long long *)workspace,
                                                                  (const float *)plan.tail, (float *)grad_value,
                                                                  n_value, accumulate);
        else
            msda::msda_det_fold<__nv_bfloat16><<<blocks, threads, 0, st>>>(
                (const long long *)workspace, (const float *)plan.tail, (__nv_bfloat16 *)grad_value, n_value,
                accumulate);
    } else {
        int64_t blocks = (n_value / 8 + threads - 1) / threads;
        if (blocks < 1) blocks = 1;
        if (blocks > device_info().sms * 16) blocks = device_info().sms * 16;
        msda::msda_fold_workspace_bf16<<<(unsigned)blocks, threads, 0, st>>>(
            (const float *)workspace, (__nv_bfloat16 *)grad_value, n_value, accumulate);
    }
    ++tl_launches;
    return check_cuda(cudaPeekAtLastError(), "fold launch");
}

// zero-fills shared by the plain and the fused backward; returns 1 when there is nothing left to launch
int backward_prologue(const msda_dims *dims, const Geometry &g, int dtype, unsigned flags, void *grad_value,
                      void *grad_pts2, void *grad_pts1, cudaStream_t st, int *rc_out, bool writes_everything = false)
{
    const int64_t n_value = dims->batch * dims->spatial_size * dims->num_heads * dims->channels;
    const bool via_workspace = (flags & MSDA_FLAG_DETERMINISTIC) || dtype == MSDA_BF16;
    *rc_out = MSDA_OK;
    if (n_value > 0 && !grad_value) {
        *rc_out = fail(MSDA_ERR_INVALID_ARGUMENT, "grad_value is null");
        return 1;
    }
    if ((flags & MSDA_FLAG_ZERO_GRAD_VALUE) && n_value > 0 && ((!via_workspace && !writes_everything) || g.rows == 0))
        if ((*rc_out = check_cuda(cudaMemsetAsync(grad_value, 0, (size_t)n_value * dtype_size(dtype), st),
                                  "memset grad_value")))
            return 1;
    if (g.rows == 0) return 1;
    if (n_value == 0) {  // no value pixels at all: every sample is out of range, all gradients are zero
        const size_t el = dtype == MSDA_F64 ? 8 : 4;
        const size_t pts = (size_t)g.rows * dims->num_levels * dims->num_point;
        if (!grad_pts2 || !grad_pts1) {
            *rc_out = fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
            return 1;
        }
        cudaMemsetAsync(grad_pts2, 0, pts * 2 * el, st);
        *rc_out = check_cuda(cudaMemsetAsync(grad_pts1, 0, pts * el, st), "memset gradients");
        return 1;
    }
    return 0;
}

}  // namespace

extern "C" {

int msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                  const void *sampling_loc, const void *attn_weight, const void *grad_output, void *grad_value,
                  void *grad_sampling_loc, void *grad_attn_weight, const msda_dims *dims, int dtype, unsigned flags,
                  void *workspace, size_t workspace_bytes, void *cuda_stream)
{
    tl_error[0] = 0;
    if (int rc = check_dims(dims, dtype)) return rc;
    const bool det = (flags & MSDA_FLAG_DETERMINISTIC) != 0;
    if (det && dtype == MSDA_F64)
        return fail(MSDA_ERR_UNSUPPORTED, "deterministic backward supports f32 and bf16 (int64 fixed point cannot "
                                          "carry fp64 precision)");
    Geometry g;
    if (int rc = geometry(dims, &g)) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;

    const bool vec_ok = vec_eligible(dims, dtype, flags) && aligned16(value) && aligned16(grad_output) &&
                        (reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0 &&
                        (reinterpret_cast<uintptr_t>(grad_sampling_loc) & 7u) == 0;
    int mode = vec_ok ? choose_bwd_mode(dims, dtype, flags) : BWD_ROW;
    if (mode == BWD_OWNED && !aligned16(grad_value)) mode = BWD_ROW;
    BinnedPlan bp{};
    OwnedPlan op{};
    if (mode == BWD_BINNED && !binned_plan((int)dims->channels, (int)dims->num_point, &bp)) mode = BWD_ROW;
    if (mode == BWD_OWNED && !owned_plan(dims, &op)) mode = BWD_ROW;

    int rc0 = MSDA_OK;
    if (backward_prologue(dims, g, dtype, flags, grad_value, grad_sampling_loc, grad_attn_weight, st, &rc0,
                          mode == BWD_OWNED))
        return rc0;
    if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_output ||
        !grad_sampling_loc || !grad_attn_weight)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");

    if (mode == BWD_OWNED) {
        // every line of grad_value is written once by its owner: no zero-fill, no workspace, no fold
        bool ok = dtype == MSDA_F32
                      ? launch_bwd_v5<float>(dims, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                             grad_output, grad_value, nullptr, grad_sampling_loc, grad_attn_weight,
                                             0x7fffffff, st)
                      : launch_bwd_v5<__nv_bfloat16>(dims, value, spatial_shapes, level_start_index, sampling_loc,
                                                     attn_weight, grad_output, grad_value, nullptr, grad_sampling_loc,
                                                     grad_attn_weight, 0x7fffffff, st);
        if (!ok) return fail(MSDA_ERR_UNSUPPORTED, "owned backward: no row kernel for this shape");
        const int accumulate = (flags & MSDA_FLAG_ZERO_GRAD_VALUE) ? 0 : 1;
        const int rc = dtype == MSDA_F32
                           ? launch_bwd_owned<float>(dims, op, spatial_shapes, level_start_index, sampling_loc,
                                                     attn_weight, grad_output, grad_value, accumulate, st)
                           : launch_bwd_owned<__nv_bfloat16>(dims, op, spatial_shapes, level_start_index, sampling_loc,
                                                             attn_weight, grad_output, grad_value, accumulate, st);
        if (rc) return rc;
        tl_launches += 2;
        snprintf(tl_kernel, sizeof(tl_kernel), "bwd_v5+owned<%s,D%d,L%d,P%d>", dtype_name(dtype), (int)dims->channels,
                 (int)dims->num_levels, (int)dims->num_point);
        return check_cuda(cudaPeekAtLastError(), "msda_backward launch");
    }

    AccPlan plan;
    if (int rc = acc_begin(dims, g, dtype, flags, grad_value, workspace, workspace_bytes, attn_weight, grad_output, st,
                           &plan))
        return rc;
    void *gv_acc = plan.gv_acc;
    const float *det_scale = plan.det_scale;

    bool done = false;
    if (mode == BWD_PLANES && vec_ok && aligned16(gv_acc) && !det_scale) {
        const int rc = dtype == MSDA_F32
                           ? launch_bwd_planes<float>(dims, value, spatial_shapes, level_start_index, sampling_loc,
                                                      attn_weight, grad_output, (float *)gv_acc, grad_sampling_loc,
                                                      grad_attn_weight, st)
                           : launch_bwd_planes<__nv_bfloat16>(dims, value, spatial_shapes, level_start_index,
                                                              sampling_loc, attn_weight, grad_output, (float *)gv_acc,
                                                              grad_sampling_loc, grad_attn_weight, st);
        if (rc > 0) return rc;
        done = rc == 0;
    }
    if (!done && vec_ok && aligned16(gv_acc)) {
        const int skip = mode == BWD_BINNED ? bp.acc_budget : 0;
        done = dtype == MSDA_F32
                   ? launch_bwd_v5<float>(dims, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                          grad_output, gv_acc, det_scale, grad_sampling_loc, grad_attn_weight, skip, st)
                   : launch_bwd_v5<__nv_bfloat16>(dims, value, spatial_shapes, level_start_index, sampling_loc,
                                                  attn_weight, grad_output, gv_acc, det_scale, grad_sampling_loc,
                                                  grad_attn_weight, skip, st);
        if (done && mode == BWD_BINNED) {
            const int rc = dtype == MSDA_F32
                               ? launch_bwd_binned<float>(dims, bp, spatial_shapes, level_start_index, sampling_loc,
                                                          attn_weight, grad_output, (float *)gv_acc, st)
                               : launch_bwd_binned<__nv_bfloat16>(dims, bp, spatial_shapes, level_start_index,
                                                                  sampling_loc, attn_weight, grad_output,
                                                                  (float *)gv_acc, st);
            if (rc) return rc;
            ++tl_launches;
            const size_t n = strlen(tl_kernel);
            snprintf(tl_kernel + n, sizeof(tl_kernel) - n, "+binned");
        }
    }
    if (!done) {
        if (dtype == MSDA_F32)
            launch_bwd_generic<float, float>(dims, g, value, spatial_shapes, level_start_index, sampling_loc,
                                             attn_weight, grad_output, gv_acc, det_scale, grad_sampling_loc,
                                             grad_attn_weight, st, "f32");
        else if (dtype == MSDA_F64)
            launch_bwd_generic<double, double>(dims, g, value, spatial_shapes, level_start_index, sampling_loc,
                                               attn_weight, grad_output, gv_acc, nullptr, grad_sampling_loc,
                                               grad_attn_weight, st, "f64");
        else
            launch_bwd_generic<__nv_bfloat16, float>(dims, g, value, spatial_shapes, level_start_index, sampling_loc,
                                                     attn_weight, grad_output, gv_acc, det_scale, grad_sampling_loc,
                                                     grad_attn_weight, st, "bf16");
    }
    ++tl_launches;
    if (int rc = check_cuda(cudaPeekAtLastError(), "msda_backward launch")) return rc;
    return acc_end(dims, dtype, flags, grad_value, workspace, plan, st);
}

// ---- fused module path (softmax + sampling-location arithmetic inside the kernels) --------------------------------

#define MSDA_FOR_EACH_FUSED_SPEC(X) MSDA_FOR_EACH_FLAGSHIP_SPEC(X)

int msda_pack_levels(void *const *level_ptrs, const int64_t *level_hw, int num_levels, int64_t batch, int64_t channels,
                     void *memory, int dtype, int unpack, void *cuda_stream)
{
    tl_error[0] = 0;
    if (!level_ptrs || !level_hw || !memory) return fail(MSDA_ERR_INVALID_ARGUMENT, "null pointer");
    if (num_levels < 1 || num_levels > 8) return fail(MSDA_ERR_UNSUPPORTED, "msda_pack_levels supports 1..8 levels");
    if (dtype != MSDA_F32 && dtype != MSDA_BF16 && dtype != MSDA_F64)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    if (batch <= 0 || channels <= 0) return MSDA_OK;
    if (batch > 65535) return fail(MSDA_ERR_UNSUPPORTED, "batch > 65535");
    msda::PackArgs a;
    a.num_levels = num_levels;
    int64_t S = 0, tiles = 0;
    for (int l = 0; l < num_levels; ++l) {
        if (!level_ptrs[l] || level_hw[l] <= 0 || level_hw[l] > 0x7fffffff)
            return fail(MSDA_ERR_INVALID_ARGUMENT, "bad level %d", l);
        a.level[l] = level_ptrs[l];
        a.hw[l] = (int)level_hw[l];
        a.start[l] = (int)S;
        a.tile_start[l] = (int)tiles;
        S += level_hw[l];
        tiles += (level_hw[l] + 31) / 32;
    }
    a.tile_start[num_levels] = (int)tiles;
    if (S > 0x7fffffff || tiles > 0x7fffffff) return fail(MSDA_ERR_INVALID_ARGUMENT, "pyramid too large");
    const dim3 grid((unsigned)tiles, (unsigned)((channels + 31) / 32), (unsigned)batch);
    cudaStream_t st = (cudaStream_t)cuda_stream;
#define LAUNCH(T)                                                                                        \
    (unpack ? msda::msda_pack_levels<T, false><<<grid, 256, 0, st>>>(a, (T *)memory, (int)channels, (int)S) \
            : msda::msda_pack_levels<T, true><<<grid, 256, 0, st>>>(a, (T *)memory, (int)channels, (int)S))
    if (dtype == MSDA_F32)
        LAUNCH(float);
    else if (dtype == MSDA_F64)
        LAUNCH(double);
    else
        LAUNCH(__nv_bfloat16);
#undef LAUNCH
    ++tl_launches;
    return check_cuda(cudaPeekAtLastError(), "msda_pack_levels launch");
}

int msda_probe_ceiling(int which, void *scratch, size_t scratch_bytes, int64_t *lines_out, void *cuda_stream)
{
    tl_error[0] = 0;
    if (!scratch || !aligned16(scratch) || scratch_bytes < (1u << 20) || !lines_out)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_probe_ceiling needs a 16-byte aligned scratch buffer of >= 1 MiB");
    if (scratch_bytes > ((size_t)1 << 36)) scratch_bytes = (size_t)1 << 36;
    const unsigned n_lines = (unsigned)(scratch_bytes / 128);
    const int warps = 4, blocks = device_info().sms * 32, iters = 1024;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (which == 0)
        msda::msda_probe_gather<<<blocks, warps * 32, 0, st>>>((const float *)scratch, (float *)scratch, n_lines, iters);
    else if (which == 1)
        msda::msda_probe_red<<<blocks, warps * 32, 0, st>>>((float *)scratch, n_lines, iters);
    else
        return fail(MSDA_ERR_INVALID_ARGUMENT, "which must be 0 (gather) or 1 (red)");
    *lines_out = (int64_t)blocks * warps * iters * 4;
    return check_cuda(cudaPeekAtLastError(), "msda_probe_ceiling launch");
}

int msda_fused_supported(const msda_dims *dims, int dtype, int ref_dim)
{
    if (!dims || (dtype != MSDA_F32 && dtype != MSDA_BF16) || (ref_dim != 2 && ref_dim != 4)) return 0;
    if (!vec_eligible(dims, dtype, 0)) return 0;
#define X(DD, LL, PP) \
    if (dims->channels == (DD) && dims->num_levels == (LL) && dims->num_point == (PP)) return 1;
    MSDA_FOR_EACH_FUSED_SPEC(X)
#undef X
    return 0;
}

int msda_mask_rows(void *data, const unsigned char *mask, int64_t n_rows, int64_t row_bytes, void *cuda_stream)
{
    tl_error[0] = 0;
    if (n_rows <= 0 || row_bytes <= 0) return MSDA_OK;
    if (!data || !mask) return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    if (row_bytes % 16 != 0 || !aligned16(data) || row_bytes > 0x7fffffff)
        return fail(MSDA_ERR_UNSUPPORTED, "msda_mask_rows needs 16-byte aligned rows");
    int64_t blocks = (n_rows + 7) / 8;
    if (blocks > device_info().sms * 32) blocks = device_info().sms * 32;
    msda::msda_mask_rows<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>((float *)data, mask, n_rows,
                                                                                         (int)row_bytes);
    ++tl_launches;
    return check_cuda(cudaPeekAtLastError(), "msda_mask_rows launch");
}

}  // extern "C"

namespace {

template <typename T, int DD, int LL, int PP, int RD>
void fused_fwd_launch(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi,
                      const void *offs, const void *logits, const void *ref, const void *vratio, void *out,
                      cudaStream_t st)
{
    constexpr int W = 4;
    const unsigned rpi = (unsigned)(d->num_query * d->num_heads);
    const dim3 grid((rpi + W - 1) / W, (unsigned)d->batch);
    bool x4 = false;
    if constexpr (sizeof(T) == 2 && DD == 32) x4 = g_bf16_x4.load() != 0;  // 8-byte lane chunks, as in fwd_v5_launch
    if constexpr (sizeof(T) == 2 && DD == 32) {
        if (x4)
            msda::msda_fwd_fused<T, DD, LL, PP, W, RD, msda::ChunkBf16x4><<<grid, W * 32, 0, st>>>(
                (const T *)value, shapes, lsi, (const float *)offs, (const float *)logits, (const float *)ref,
                (const float *)vratio, (T *)out, (int)d->spatial_size, (int)d->num_heads, (int)d->num_query, rpi);
    }
    if (!x4)
        msda::msda_fwd_fused<T, DD, LL, PP, W, RD><<<grid, W * 32, 0, st>>>(
            (const T *)value, shapes, lsi, (const float *)offs, (const float *)logits, (const float *)ref,
            (const float *)vratio, (T *)out, (int)d->spatial_size, (int)d->num_heads, (int)d->num_query, rpi);
    snprintf(tl_kernel, sizeof(tl_kernel), "fwd_fused<%s,D%d,L%d,P%d,ref%d%s>", tname<T>(), DD, LL, PP, RD,
             x4 ? ",x4" : "");
}

template <typename T, int DD, int LL, int PP, int RD>
void fused_bwd_launch(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi,
                      const void *offs, const void *logits, const void *ref, const void *vratio, const void *gout,
                      void *gv_acc,
                      const float *det_scale, void *goffs, void *glogits, cudaStream_t st)
{
    constexpr int W = 4;
    const unsigned rpi = (unsigned)(d->num_query * d->num_heads);
    const dim3 grid((rpi + W - 1) / W, (unsigned)d->batch);
    using CH = typename BwdChunk<T>::type;
    if (det_scale)
        msda::msda_bwd_fused<T, CH, msda::AccFix64, DD, LL, PP, W, RD><<<grid, W * 32, 0, st>>>(
            (const T *)value, shapes, lsi, (const float *)offs, (const float *)logits, (const float *)ref,
            (const float *)vratio, (const T *)gout, (unsigned long long *)gv_acc, det_scale, (float *)goffs, (float *)glogits,
            (int)d->spatial_size, (int)d->num_heads, (int)d->num_query, rpi);
    else
        msda::msda_bwd_fused<T, CH, msda::AccF32, DD, LL, PP, W, RD><<<grid, W * 32, 0, st>>>(
            (const T *)value, shapes, lsi, (const float *)offs, (const float *)logits, (const float *)ref,
            (const float *)vratio, (const T *)gout, (float *)gv_acc, nullptr, (float *)goffs, (float *)glogits, (int)d->spatial_size,
            (int)d->num_heads, (int)d->num_query, rpi);
    snprintf(tl_kernel, sizeof(tl_kernel), "bwd_fused<%s,D%d,L%d,P%d,ref%d%s>", tname<T>(), DD, LL, PP, RD,
             det_scale ? ",deterministic" : "");
}

// planes backward of the fused module path (D = 32 only: the auto rule's domain); one 768-thread CTA per SM
template <typename T, int LL, int PP, int RD>
int fused_bwd_planes_launch(const msda_dims *d, const void *value, const int64_t *shapes, const int64_t *lsi,
                            const void *offs, const void *logits, const void *ref, const void *vratio, const void *gout,
                            float *gv_acc, void *goffs, void *glogits, cudaStream_t st)
{
    constexpr int DD = 32, TH = 768;
    using CH = typename BwdChunk<T>::type;
    auto kernel = msda::msda_bwd_planes<T, CH, DD, LL, PP, TH, RD>;
    const int smem = (device_info().max_smem_optin - 2048 - (TH / 32) * DD * 4) & ~15;
    if (smem <= 0) return fail(MSDA_ERR_CUDA, "device reports no opt-in shared memory");
    if (int rc = optin_smem(kernel, smem)) return rc;
    const int64_t chunks = planes_chunks(d, TH);
    const int64_t items = d->batch * chunks * d->num_heads;
    if (items > 0x7fffffffLL) return fail(MSDA_ERR_INVALID_ARGUMENT, "grid too large");
    kernel<<<(unsigned)items, TH, smem, st>>>((const T *)value, shapes, lsi, (const float *)offs, (const float *)logits,
                                              (const T *)gout, gv_acc, (float *)goffs, (float *)glogits, (int)d->batch,
                                              (int)d->spatial_size, (int)d->num_heads, (int)d->num_query, smem / 4,
                                              (int)chunks, (const float *)ref, (const float *)vratio);
    snprintf(tl_kernel, sizeof(tl_kernel), "bwd_planes_fused<%s,D%d,L%d,P%d,ref%d,t%d>", tname<T>(), DD, LL, PP, RD, TH);
    return MSDA_OK;
}

}  // namespace

extern "C" {

int msda_fused_forward_vr(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                          const void *sampling_offsets, const void *attn_logits, const void *reference_points,
                          const void *valid_ratios, int ref_dim, void *output, const msda_dims *dims, int dtype,
                          unsigned flags, void *cuda_stream)
{
    tl_error[0] = 0;
    (void)flags;
    if (int rc = check_dims(dims, dtype)) return rc;
    if (!msda_fused_supported(dims, dtype, ref_dim))
        return fail(MSDA_ERR_UNSUPPORTED, "no fused specialisation for D=%lld L=%lld P=%lld dtype=%s ref_dim=%d",
                    (long long)dims->channels, (long long)dims->num_levels, (long long)dims->num_point,
                    dtype_name(dtype), ref_dim);
    Geometry g;
    if (int rc = geometry(dims, &g)) return rc;
    if (g.rows == 0) return MSDA_OK;
    if (!value || !spatial_shapes || !level_start_index || !sampling_offsets || !attn_logits || !reference_points ||
        !output)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    if (!aligned16(value) || !aligned16(output) || !aligned16(reference_points) ||
        (reinterpret_cast<uintptr_t>(sampling_offsets) & 7u) || (reinterpret_cast<uintptr_t>(valid_ratios) & 7u))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "fused kernels need 16-byte aligned value/output/reference_points");
    cudaStream_t st = (cudaStream_t)cuda_stream;
#define ARGS dims, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points, valid_ratios, output, st
#define X(DD, LL, PP)                                                                                 \
    if (dims->channels == (DD) && dims->num_levels == (LL) && dims->num_point == (PP)) {             \
        if (dtype == MSDA_F32)                                                                        \
            ref_dim == 2 ? fused_fwd_launch<float, DD, LL, PP, 2>(ARGS)                               \
                         : fused_fwd_launch<float, DD, LL, PP, 4>(ARGS);                              \
        else                                                                                          \
            ref_dim == 2 ? fused_fwd_launch<__nv_bfloat16, DD, LL, PP, 2>(ARGS)                       \
                         : fused_fwd_launch<__nv_bfloat16, DD, LL, PP, 4>(ARGS);                      \
    }
    MSDA_FOR_EACH_FUSED_SPEC(X)
#undef X
#undef ARGS
    ++tl_launches;
    return check_cuda(cudaPeekAtLastError(), "msda_fused_forward launch");
}

int msda_fused_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_offsets, const void *attn_logits, const void *reference_points,
                       int ref_dim, void *output, const msda_dims *dims, int dtype, unsigned flags, void *cuda_stream)
{
    return msda_fused_forward_vr(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                                 reference_points, nullptr, ref_dim, output, dims, dtype, flags, cuda_stream);
}

int msda_fused_backward_vr(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                           const void *sampling_offsets, const void *attn_logits, const void *reference_points,
                           const void *valid_ratios, int ref_dim, const void *grad_output, void *grad_value,
                           void *grad_offsets, void *grad_logits, const msda_dims *dims, int dtype, unsigned flags,
                           void *workspace, size_t workspace_bytes, void *cuda_stream)
{
    tl_error[0] = 0;
    if (int rc = check_dims(dims, dtype)) return rc;
    if (!msda_fused_supported(dims, dtype, ref_dim))
        return fail(MSDA_ERR_UNSUPPORTED, "no fused specialisation for this shape/dtype");
    Geometry g;
    if (int rc = geometry(dims, &g)) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc0 = MSDA_OK;
    if (backward_prologue(dims, g, dtype, flags, grad_value, grad_offsets, grad_logits, st, &rc0)) return rc0;
    if (!value || !spatial_shapes || !level_start_index || !sampling_offsets || !attn_logits || !reference_points ||
        !grad_output || !grad_offsets || !grad_logits)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    if (!aligned16(value) || !aligned16(grad_output) || !aligned16(grad_value) || !aligned16(workspace) ||
        !aligned16(reference_points) || (reinterpret_cast<uintptr_t>(sampling_offsets) & 7u) ||
        (reinterpret_cast<uintptr_t>(grad_offsets) & 7u) || (reinterpret_cast<uintptr_t>(valid_ratios) & 7u))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "fused kernels need 16-byte aligned tensors");
    AccPlan plan;
    if (int rc = acc_begin(dims, g, dtype, flags, grad_value, workspace, workspace_bytes, nullptr, grad_output, st,
                           &plan))
        return rc;
    // dense D=32 problems: the planes backward (same rule as msda_backward), here with the fused point source
    if (!plan.det && dims->channels == 32 && dims->num_levels == 4 && dims->num_point == 4 &&
        choose_bwd_mode(dims, dtype, flags) == BWD_PLANES) {
#define PARGS                                                                                                   \
    dims, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points, valid_ratios, \
        grad_output, (float *)plan.gv_acc, grad_offsets, grad_logits, st
        int rc = MSDA_OK;
        if (dtype == MSDA_F32)
            rc = ref_dim == 2 ? fused_bwd_planes_launch<float, 4, 4, 2>(PARGS) : fused_bwd_planes_launch<float, 4, 4, 4>(PARGS);
        else
            rc = ref_dim == 2 ? fused_bwd_planes_launch<__nv_bfloat16, 4, 4, 2>(PARGS)
                              : fused_bwd_planes_launch<__nv_bfloat16, 4, 4, 4>(PARGS);
#undef PARGS
        if (rc) return rc;
        ++tl_launches;
        if (int rc2 = check_cuda(cudaPeekAtLastError(), "msda_fused_backward launch")) return rc2;
        return acc_end(dims, dtype, flags, grad_value, workspace, plan, st);
    }
#define ARGS                                                                                                    \
    dims, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points, valid_ratios, \
        grad_output, plan.gv_acc, plan.det_scale, grad_offsets, grad_logits, st
#define X(DD, LL, PP)                                                                                 \
    if (dims->channels == (DD) && dims->num_levels == (LL) && dims->num_point == (PP)) {             \
        if (dtype == MSDA_F32)                                                                        \
            ref_dim == 2 ? fused_bwd_launch<float, DD, LL, PP, 2>(ARGS)                               \
                         : fused_bwd_launch<float, DD, LL, PP, 4>(ARGS);                              \
        else                                                                                          \
            ref_dim == 2 ? fused_bwd_launch<__nv_bfloat16, DD, LL, PP, 2>(ARGS)                       \
                         : fused_bwd_launch<__nv_bfloat16, DD, LL, PP, 4>(ARGS);                      \
    }
    MSDA_FOR_EACH_FUSED_SPEC(X)
#undef X
#undef ARGS
    ++tl_launches;
    if (int rc = check_cuda(cudaPeekAtLastError(), "msda_fused_backward launch")) return rc;
    return acc_end(dims, dtype, flags, grad_value, workspace, plan, st);
}

int msda_fused_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const void *sampling_offsets, const void *attn_logits, const void *reference_points,
                        int ref_dim, const void *grad_output, void *grad_value, void *grad_offsets, void *grad_logits,
                        const msda_dims *dims, int dtype, unsigned flags, void *workspace, size_t workspace_bytes,
                        void *cuda_stream)
{
    return msda_fused_backward_vr(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                                  reference_points, nullptr, ref_dim, grad_output, grad_value, grad_offsets,
                                  grad_logits, dims, dtype, flags, workspace, workspace_bytes, cuda_stream);
}

// ---- decoder-layer epilogue: y = LayerNorm(x + dropout(z)) (msda_kernels_layer.cuh) -----------------------------------

static int ln_blocks(int64_t rows)
{
    int64_t b = (rows + 7) / 8;
    return (int)(b < 1 ? 1 : (b > 1024 ? 1024 : b));
}

size_t msda_add_dropout_ln_workspace_bytes(int64_t rows, int64_t channels)
{
    if (rows <= 0 || channels <= 0) return 0;
    return (size_t)ln_blocks(rows) * 2 * (size_t)channels * sizeof(float);
}

int msda_add_dropout_ln_supported(int64_t channels) { return channels > 0 && channels % 128 == 0 && channels <= 512; }

int msda_add_dropout_ln_forward(const void *x, const void *z, const unsigned char *keep, float keep_scale,
                                const void *gamma, const void *beta, float eps, void *y, void *h_saved, void *mean,
                                void *rstd, int64_t rows, int64_t channels, void *cuda_stream)
{
    tl_error[0] = 0;
    if (rows < 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "negative row count");
    if (!msda_add_dropout_ln_supported(channels))
        return fail(MSDA_ERR_UNSUPPORTED, "add+dropout+LayerNorm supports channels in {128, 256, 384, 512}, got %lld",
                    (long long)channels);
    if (rows == 0) return MSDA_OK;
    if (!x || !z || !gamma || !beta || !y) return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    if ((h_saved != nullptr) != (mean != nullptr) || (mean != nullptr) != (rstd != nullptr))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "h_saved, mean and rstd must be given together (training) or all be null");
    if (!aligned16(x) || !aligned16(z) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta) || !aligned16(h_saved) ||
        (reinterpret_cast<uintptr_t>(keep) & 3u))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "add+dropout+LayerNorm needs 16-byte aligned tensors");
    const unsigned grid = (unsigned)((rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)cuda_stream;
#define LAUNCH(V)                                                                                                     \
    msda::msda_add_dropout_ln_fwd<V><<<grid, 256, 0, st>>>((const float *)x, (const float *)z, keep, keep_scale,      \
                                                           (const float *)gamma, (const float *)beta, eps, (float *)y, \
                                                           (float *)h_saved, (float *)mean, (float *)rstd, rows)
    switch (channels / 128) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        default: LAUNCH(4); break;
    }
#undef LAUNCH
    ++tl_launches;
    snprintf(tl_kernel, sizeof(tl_kernel), "add_dropout_ln_fwd<C%d>", (int)channels);
    return check_cuda(cudaPeekAtLastError(), "msda_add_dropout_ln_forward launch");
}

int msda_add_dropout_ln_backward(const void *grad_y, const void *h_saved, const void *mean, const void *rstd,
                                 const unsigned char *keep, float keep_scale, const void *gamma, void *grad_x,
                                 void *grad_z, void *grad_gamma, void *grad_beta, void *workspace,
                                 size_t workspace_bytes, int64_t rows, int64_t channels, void *cuda_stream)
{
    tl_error[0] = 0;
    if (rows < 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "negative row count");
    if (!msda_add_dropout_ln_supported(channels))
        return fail(MSDA_ERR_UNSUPPORTED, "add+dropout+LayerNorm supports channels in {128, 256, 384, 512}, got %lld",
                    (long long)channels);
    if (!grad_gamma || !grad_beta) return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (rows == 0) {
        cudaMemsetAsync(grad_gamma, 0, (size_t)channels * 4, st);
        return check_cuda(cudaMemsetAsync(grad_beta, 0, (size_t)channels * 4, st), "memset");
    }
    if (!grad_y || !h_saved || !mean || !rstd || !gamma || !grad_x || !grad_z)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    const size_t need = msda_add_dropout_ln_workspace_bytes(rows, channels);
    if (!workspace || workspace_bytes < need)
        return fail(MSDA_ERR_WORKSPACE, "add+dropout+LayerNorm backward needs a %zu-byte workspace, got %zu", need,
                    workspace_bytes);
    if (!aligned16(grad_y) || !aligned16(h_saved) || !aligned16(grad_x) || !aligned16(grad_z) || !aligned16(gamma) ||
        !aligned16(workspace) || (reinterpret_cast<uintptr_t>(keep) & 3u))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "add+dropout+LayerNorm needs 16-byte aligned tensors");
    const int blocks = ln_blocks(rows);
#define LAUNCH(V)                                                                                                    \
    msda::msda_add_dropout_ln_bwd<V><<<blocks, 256, 0, st>>>((const float *)grad_y, (const float *)h_saved,          \
                                                             (const float *)mean, (const float *)rstd, keep,          \
                                                             keep_scale, (const float *)gamma, (float *)grad_x,       \
                                                             (float *)grad_z, (float *)workspace, rows)
    switch (channels / 128) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        default: LAUNCH(4); break;
    }
#undef LAUNCH
    msda::msda_ln_param_grads<<<(unsigned)((channels + 127) / 128), 128, 0, st>>>(
        (const float *)workspace, blocks, (int)channels, (float *)grad_gamma, (float *)grad_beta);
    tl_launches += 2;
    snprintf(tl_kernel, sizeof(tl_kernel), "add_dropout_ln_bwd<C%d>", (int)channels);
    return check_cuda(cudaPeekAtLastError(), "msda_add_dropout_ln_backward launch");
}

// ---- GroupNorm epilogue -> packed memory (msda_kernels_layer.cuh) ---------------------------------------------------------
int msda_pack_levels_groupnorm(void *const *level_ptrs, const int64_t *level_hw, int num_levels, int64_t batch,
                               int64_t channels, int num_groups, void *const *gamma_ptrs, void *const *beta_ptrs, float eps,
                               void *memory, int out_dtype, void *stats, void *cuda_stream)
{
    tl_error[0] = 0;
    if (!level_ptrs || !level_hw || !gamma_ptrs || !beta_ptrs || !memory || !stats)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "null pointer");
    if (num_levels < 1 || num_levels > 8) return fail(MSDA_ERR_UNSUPPORTED, "msda_pack_levels_groupnorm supports 1..8 levels");
    if (out_dtype != MSDA_F32 && out_dtype != MSDA_BF16)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "memory dtype must be f32 or bf16");
    if (num_groups <= 0 || channels <= 0 || channels % num_groups != 0)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "channels must be a positive multiple of num_groups");
    if (batch <= 0) return MSDA_OK;
    if (batch > 65535 || num_groups > 65535) return fail(MSDA_ERR_UNSUPPORTED, "batch / num_groups > 65535");
    msda::GnPackArgs a;
    a.num_levels = num_levels;
    int64_t S = 0, tiles = 0;
    for (int l = 0; l < num_levels; ++l) {
        if (!level_ptrs[l] || !gamma_ptrs[l] || !beta_ptrs[l] || level_hw[l] <= 0 || level_hw[l] > 0x7fffffff)
            return fail(MSDA_ERR_INVALID_ARGUMENT, "bad level %d", l);
        a.level[l] = (const float *)level_ptrs[l];
        a.gamma[l] = (const float *)gamma_ptrs[l], a.beta[l] = (const float *)beta_ptrs[l];
        a.hw[l] = (int)level_hw[l];
        a.start[l] = (int)S;
        a.tile_start[l] = (int)tiles;
        S += level_hw[l];
        tiles += (level_hw[l] + 31) / 32;
    }
    a.tile_start[num_levels] = (int)tiles;
    if (S > 0x7fffffff || tiles > 0x7fffffff) return fail(MSDA_ERR_INVALID_ARGUMENT, "pyramid too large");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    msda::msda_gn_stats<<<dim3((unsigned)num_groups, (unsigned)batch, (unsigned)num_levels), 256, 0, st>>>(
        a, (int)batch, (int)channels, num_groups, eps, (float *)stats);
    const dim3 grid((unsigned)tiles, (unsigned)((channels + 31) / 32), (unsigned)batch);
    if (out_dtype == MSDA_F32)
        msda::msda_pack_levels_gn<float><<<grid, 256, 0, st>>>(a, (const float *)stats, (float *)memory, (int)batch,
                                                               (int)channels, num_groups, (int)S);
    else
        msda::msda_pack_levels_gn<__nv_bfloat16><<<grid, 256, 0, st>>>(a, (const float *)stats, (__nv_bfloat16 *)memory,
                                                                       (int)batch, (int)channels, num_groups, (int)S);
    tl_launches += 2;
    snprintf(tl_kernel, sizeof(tl_kernel), "pack_levels_gn<%s>", dtype_name(out_dtype));
    return check_cuda(cudaPeekAtLastError(), "msda_pack_levels_groupnorm launch");
}

// ---- host-buffer session ---------------------------------------------------------------------------

struct msda_host_session {
    static constexpr int kMaxSlots = 8;
    int n_slots;  // pipeline depth: chunks in flight (H2D of one, kernels of another, D2H of a third, ...)
    msda_dims max_dims;
    int dtype;
    int device;
    int chunk;  // images per chunk
    int next_slot;  // round-robin cursor; persists across submits so consecutive calls keep the pipeline full
    int meta_levels;  // number of levels currently uploaded (0 = none)
    int64_t h_shapes[2 * 64], h_lsi[64];  // host copy of the uploaded level metadata
    int64_t *d_shapes, *d_lsi;
    cudaStream_t stream[kMaxSlots];
    struct Slot {
        char *value, *loc, *attn, *gout, *out, *gvalue, *gloc, *gattn, *ws;
    } slot[kMaxSlots];
    size_t ws_bytes;
};

static void session_free(msda_host_session *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    for (int i = 0; i < msda_host_session::kMaxSlots; ++i) {
        if (s->stream[i]) cudaStreamDestroy(s->stream[i]);
        char *ptrs[] = {s->slot[i].value, s->slot[i].loc,    s->slot[i].attn, s->slot[i].gout, s->slot[i].out,
                        s->slot[i].gvalue, s->slot[i].gloc,  s->slot[i].gattn, s->slot[i].ws};
        for (char *p : ptrs)
            if (p) cudaFree(p);
    }
    if (s->d_shapes) cudaFree(s->d_shapes);
    if (s->d_lsi) cudaFree(s->d_lsi);
    delete s;
}

int msda_host_session_create(msda_host_session **session, const msda_dims *max_dims, int dtype, int device,
                             int images_per_chunk)
{
    tl_error[0] = 0;
    if (!session) return fail(MSDA_ERR_INVALID_ARGUMENT, "session out-pointer is null");
    *session = nullptr;
    if (int rc = check_dims(max_dims, dtype)) return rc;
    if (images_per_chunk <= 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "images_per_chunk must be positive");
    if (max_dims->num_levels > 64) return fail(MSDA_ERR_UNSUPPORTED, "host sessions support at most 64 levels");
    if (int rc = check_cuda(cudaSetDevice(device), "cudaSetDevice")) return rc;
    auto *s = new msda_host_session();
    memset(s, 0, sizeof(*s));
    s->max_dims = *max_dims;
    s->dtype = dtype;
    s->device = device;
    s->chunk = images_per_chunk;
    s->n_slots = 4;
    if (const char *env = getenv("MSDA_HOST_SLOTS")) {
        const int v = atoi(env);
        if (v >= 1 && v <= msda_host_session::kMaxSlots) s->n_slots = v;
    }
    const size_t ev = dtype_size(dtype), el = dtype == MSDA_F64 ? 8 : 4;
    const msda_dims &d = *max_dims;
    const size_t c = (size_t)images_per_chunk;
    const size_t n_val = c * d.spatial_size * d.num_heads * d.channels;
    const size_t n_pts = c * d.num_query * d.num_heads * d.num_levels * d.num_point;
    const size_t n_out = c * d.num_query * d.num_heads * d.channels;
    msda_dims cd = d;
    cd.batch = images_per_chunk;
    // sized for the most demanding mode (deterministic) so any flags accepted by msda_backward work through a session
    s->ws_bytes = msda_backward_workspace_bytes(&cd, dtype, dtype == MSDA_F64 ? 0 : MSDA_FLAG_DETERMINISTIC);
    cudaError_t e = cudaSuccess;
    auto alloc = [&](char **p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc((void **)p, bytes ? bytes : 16);
    };
    alloc((char **)&s->d_shapes, sizeof(int64_t) * 2 * d.num_levels);
    alloc((char **)&s->d_lsi, sizeof(int64_t) * d.num_levels);
    for (int i = 0; i < s->n_slots; ++i) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream[i], cudaStreamNonBlocking);
        auto &k = s->slot[i];
        alloc(&k.value, n_val * ev), alloc(&k.gvalue, n_val * ev);
        alloc(&k.loc, n_pts * 2 * el), alloc(&k.gloc, n_pts * 2 * el);
        alloc(&k.attn, n_pts * el), alloc(&k.gattn, n_pts * el);
        alloc(&k.gout, n_out * ev), alloc(&k.out, n_out * ev);
        alloc(&k.ws, s->ws_bytes);
    }
    if (e != cudaSuccess) {
        session_free(s);
        return check_cuda(e, "msda_host_session_create");
    }
    *session = s;
    return MSDA_OK;
}

void msda_host_session_destroy(msda_host_session *session) { session_free(session); }

int msda_host_submit(msda_host_session *s, const void *value, const int64_t *spatial_shapes,
                     const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight,
                     const void *grad_output, void *output, void *grad_value, void *grad_sampling_loc,
                     void *grad_attn_weight, const msda_dims *dims, unsigned flags)
{
    tl_error[0] = 0;
    if (!s) return fail(MSDA_ERR_INVALID_ARGUMENT, "session is null");
    if (int rc = check_dims(dims, s->dtype)) return rc;
    const msda_dims &mx = s->max_dims;
    if (dims->spatial_size > mx.spatial_size || dims->num_heads * dims->channels > mx.num_heads * mx.channels ||
        dims->num_query * dims->num_heads * dims->num_levels * dims->num_point >
            mx.num_query * mx.num_heads * mx.num_levels * mx.num_point ||
        dims->num_query * dims->num_heads * dims->channels > mx.num_query * mx.num_heads * mx.channels ||
        dims->spatial_size * dims->num_heads * dims->channels > mx.spatial_size * mx.num_heads * mx.channels ||
        dims->num_levels > mx.num_levels)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "problem exceeds the session's max_dims");
    if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_output || !output ||
        !grad_value || !grad_sampling_loc || !grad_attn_weight)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "null host pointer");
    if (int rc = check_cuda(cudaSetDevice(s->device), "cudaSetDevice")) return rc;

    const size_t ev = dtype_size(s->dtype), el = s->dtype == MSDA_F64 ? 8 : 4;
    const size_t img_val = (size_t)(dims->spatial_size * dims->num_heads * dims->channels) * ev;
    const size_t img_pts = (size_t)(dims->num_query * dims->num_heads * dims->num_levels * dims->num_point) * el;
    const size_t img_out = (size_t)(dims->num_query * dims->num_heads * dims->channels) * ev;
    const int K = s->n_slots;
    const int L = (int)dims->num_levels;

    // Level metadata lives on the device for the whole session.  It is uploaded only when it changes; work already in
    // flight may still be reading the old copy, so a change drains the pipeline first (rare: one pyramid per model).
    bool meta_changed = s->meta_levels != L;
    for (int i = 0; i < L && !meta_changed; ++i)
        meta_changed = s->h_shapes[2 * i] != spatial_shapes[2 * i] || s->h_shapes[2 * i + 1] != spatial_shapes[2 * i + 1] ||
                       s->h_lsi[i] != level_start_index[i];
    cudaError_t e = cudaSuccess;
    auto track = [&](cudaError_t ei) {
        if (e == cudaSuccess) e = ei;
    };
    if (meta_changed) {
        for (int i = 0; i < K; ++i) track(cudaStreamSynchronize(s->stream[i]));
        memcpy(s->h_shapes, spatial_shapes, sizeof(int64_t) * 2 * L);
        memcpy(s->h_lsi, level_start_index, sizeof(int64_t) * L);
        s->meta_levels = L;
        track(cudaMemcpyAsync(s->d_shapes, s->h_shapes, sizeof(int64_t) * 2 * L, cudaMemcpyHostToDevice, s->stream[0]));
        track(cudaMemcpyAsync(s->d_lsi, s->h_lsi, sizeof(int64_t) * L, cudaMemcpyHostToDevice, s->stream[0]));
        track(cudaStreamSynchronize(s->stream[0]));
        if (e != cudaSuccess) return check_cuda(e, "msda_host_submit (level metadata)");
    }

    int rc = MSDA_OK;
    for (int64_t b0 = 0; b0 < dims->batch && rc == MSDA_OK; b0 += s->chunk) {
        const int64_t nb = (dims->batch - b0 < s->chunk) ? dims->batch - b0 : s->chunk;
        const int slot = s->next_slot;
        s->next_slot = (slot + 1) % K;
        auto &k = s->slot[slot];
        cudaStream_t st = s->stream[slot];  // stream order serialises successive uses of the slot's buffers
        msda_dims cd = *dims;
        cd.batch = nb;
        track(cudaMemcpyAsync(k.value, (const char *)value + b0 * img_val, nb * img_val, cudaMemcpyHostToDevice, st));
        track(cudaMemcpyAsync(k.loc, (const char *)sampling_loc + b0 * img_pts * 2, nb * img_pts * 2,
                              cudaMemcpyHostToDevice, st));
        track(cudaMemcpyAsync(k.attn, (const char *)attn_weight + b0 * img_pts, nb * img_pts, cudaMemcpyHostToDevice, st));
        track(cudaMemcpyAsync(k.gout, (const char *)grad_output + b0 * img_out, nb * img_out, cudaMemcpyHostToDevice, st));
        rc = msda_forward(k.value, s->d_shapes, s->d_lsi, k.loc, k.attn, k.out, &cd, s->dtype, flags, st);
        if (rc) break;
        rc = msda_backward(k.value, s->d_shapes, s->d_lsi, k.loc, k.attn, k.gout, k.gvalue, k.gloc, k.gattn, &cd,
                           s->dtype, flags | MSDA_FLAG_ZERO_GRAD_VALUE, k.ws, s->ws_bytes, st);
        if (rc) break;
        track(cudaMemcpyAsync((char *)output + b0 * img_out, k.out, nb * img_out, cudaMemcpyDeviceToHost, st));
        track(cudaMemcpyAsync((char *)grad_value + b0 * img_val, k.gvalue, nb * img_val, cudaMemcpyDeviceToHost, st));
        track(cudaMemcpyAsync((char *)grad_sampling_loc + b0 * img_pts * 2, k.gloc, nb * img_pts * 2,
                              cudaMemcpyDeviceToHost, st));
        track(cudaMemcpyAsync((char *)grad_attn_weight + b0 * img_pts, k.gattn, nb * img_pts, cudaMemcpyDeviceToHost, st));
    }
    if (rc) return rc;
    return check_cuda(e, "msda_host_submit");
}

int msda_host_wait(msda_host_session *s)
{
    tl_error[0] = 0;
    if (!s) return fail(MSDA_ERR_INVALID_ARGUMENT, "session is null");
    if (int rc = check_cuda(cudaSetDevice(s->device), "cudaSetDevice")) return rc;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < s->n_slots; ++i) {
        const cudaError_t ei = cudaStreamSynchronize(s->stream[i]);
        if (e == cudaSuccess) e = ei;
    }
    return check_cuda(e, "msda_host_wait");
}

int msda_host_forward_backward(msda_host_session *s, const void *value, const int64_t *spatial_shapes,
                               const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight,
                               const void *grad_output, void *output, void *grad_value, void *grad_sampling_loc,
                               void *grad_attn_weight, const msda_dims *dims, unsigned flags)
{
    const int rc = msda_host_submit(s, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                    output, grad_value, grad_sampling_loc, grad_attn_weight, dims, flags);
    const int rc_wait = msda_host_wait(s);  // drain even after a failed submit: earlier chunks may be in flight
    return rc ? rc : rc_wait;
}

}  // extern "C"
