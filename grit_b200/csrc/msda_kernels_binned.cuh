// msda_kernels_binned.cuh -- grad_value of the COARSE pyramid levels by on-SM pre-aggregation (sm_100a).
//
// Why: in the row-parallel backward (msda_bwd_v5) every bilinear tap leaves the SM as one 128-byte
// `red.global.add.v4.f32`; the SM's egress path moves one such line per ~5 cycles, which is the whole cost of
// that kernel (profiles/r01_ncu_v5_fwd_bwd.txt: 601.8 M red sectors, l1tex->xbar request path 90.7 % busy).  The
// coarse levels of one (image, head) pair are tiny -- at 800x1333 levels 2+3 are 1323 pixels, 169 KB of fp32
// gradient -- yet receive half of all taps, each pixel thousands of times.  sm_100a has no native fp32
// shared-memory atomic (atomicAdd(float) on shared memory is an LDS/FADD/ATOMS.CAS loop, measured slower than the
// global reds in round 1), so the accumulation is organised so that it needs none:
//
//   * a work item is (image, head, coarse level, slice of the queries); one persistent 1024-thread CTA per SM
//     keeps the level's fp32 gradient plane of that (image, head) in shared memory for the whole item;
//   * the item's rows are walked in tiles.  Per tile: (1) every sample POINT of the level is resolved and counted
//     into the bin of its top-left tap with a native integer shared-memory atomic (ATOMS.ADD returns the rank
//     inside the bin); (2) a block-wide exclusive scan turns counts into offsets; (3) the points' records
//     {row, attention weight, fractional offsets} are scattered into bin order -- a counting sort, integer
//     atomics only; (4) every bin is OWNED by one lane group, which reads its records, multiplies the rows'
//     grad_output (staged in shared memory per tile) and accumulates the four taps' contributions in registers,
//     then adds them to the shared-memory plane with plain loads/stores.  Bins are processed in four colour
//     phases (parity of the bin's row and column), inside a phase the 2x2 pixel footprints of all bins are
//     disjoint, so no two owners touch the same pixel and no atomics are needed;
//   * when the item ends the plane is flushed with one vector red per non-zero 16-byte chunk: a coarse pixel
//     leaves the SM once per item instead of once per tap.
//
// Which levels take this path is decided ON THE DEVICE from spatial_shapes (no host sync) by level_is_binned();
// msda_bwd_v5 evaluates the same rule and skips its reds for exactly those levels.
// Semantics (validity window, -0.5 shift, per-tap zero padding) are resolve_taps(), shared with every other kernel;
// replaces the grad_value part of ms_deform_attn_col2im_bilinear (ms_deform_im2col_cuda.cuh:87-159).
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kBinThreads = 1024;
constexpr int kBinMaxPpt = 4;      // sample points one thread resolves per tile
constexpr int kBinMaxLevels = 8;

// A level is binned when its fp32 gradient plane for one head fits the accumulator budget (the host derives the
// budget from the opt-in shared memory size: plane + one minimal tile of staging + bin counters must fit).
__device__ __forceinline__ bool level_is_binned(int H, int W, int D, int acc_budget_bytes)
{
    return H > 0 && W > 0 && (long long)H * W * D * 4 <= (long long)acc_budget_bytes;
}

// Sample point -> cell of the level: top-left tap (r0, c0) with r0 in [-1, H-1], c0 in [-1, W-1], fractional offsets,
// and whether the point is inside the (-1, size) window at all.  Same arithmetic as resolve_taps().
struct Cell {
    int r0, c0;
    float lh, lw;
    bool live;
};
__device__ __forceinline__ Cell resolve_cell(float x, float y, int H, int W)
{
    Cell c;
    float h_im = fmaf(y, (float)H, -0.5f);
    float w_im = fmaf(x, (float)W, -0.5f);
    c.live = h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;  // false for NaN
    if (!c.live) h_im = w_im = 0.f;
    const float hf = floorf(h_im), wf = floorf(w_im);
    c.r0 = (int)hf, c.c0 = (int)wf;
    c.lh = h_im - hf, c.lw = w_im - wf;
    return c;
}

// Bit l set = level l is binned.  Evaluated identically by msda_bwd_v5 (which then skips its reds for those
// levels) and msda_bwd_binned (which produces them).
template <int L>
__device__ __forceinline__ unsigned binned_level_mask(const int (&sH)[L], const int (&sW)[L], int D, int acc_budget_bytes)
{
    unsigned mask = 0;
#pragma unroll
    for (int l = 0; l < L; ++l)
        if (level_is_binned(sH[l], sW[l], D, acc_budget_bytes)) mask |= 1u << l;
    return mask;
}

// Exclusive scan of cnt[0..n) into off[0..n], off[n] = total; with RESET cnt is zeroed afterwards (off may alias cnt
// when RESET is false).  All kBinThreads threads call it.
template <bool RESET>
__device__ __forceinline__ void block_exclusive_scan(unsigned *cnt, unsigned *off, int n, unsigned *warp_sums)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int per = (n + kBinThreads - 1) / kBinThreads;
    const int b0 = t * per;
    unsigned local = 0;
    for (int i = 0; i < per; ++i)
        if (b0 + i < n) local += cnt[b0 + i];
    unsigned incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned w = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += v;
        }
        warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    unsigned run = incl - local + (warp > 0 ? warp_sums[warp - 1] : 0u);
    for (int i = 0; i < per; ++i)
        if (b0 + i < n) {
            const unsigned c = cnt[b0 + i];
            if (RESET) cnt[b0 + i] = 0;
            off[b0 + i] = run;
            run += c;
        }
    if (t == kBinThreads - 1) off[n] = warp_sums[31];
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ float4 load4_as_float(const T *p);
template <>
__device__ __forceinline__ float4 load4_as_float<float>(const float *p)
{
    return __ldg(reinterpret_cast<const float4 *>(p));
}
template <>
__device__ __forceinline__ float4 load4_as_float<__nv_bfloat16>(const __nv_bfloat16 *p)
{
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                       __uint_as_float(v.y & 0xffff0000u));
}

__device__ __forceinline__ void fma4(float4 &acc, float w, const float4 &g)
{
    acc.x = fmaf(w, g.x, acc.x), acc.y = fmaf(w, g.y, acc.y), acc.z = fmaf(w, g.z, acc.z), acc.w = fmaf(w, g.w, acc.w);
}

__device__ __forceinline__ void smem_accumulate4(float *p, const float4 &a)
{
    float4 v = *reinterpret_cast<float4 *>(p);
    v.x += a.x, v.y += a.y, v.z += a.z, v.w += a.w;
    *reinterpret_cast<float4 *>(p) = v;
}

// Dynamic shared memory layout (bytes), all 16-byte aligned:
//   [region : region_bytes][cnt : bins_cap * 4][off : (bins_cap + 4) * 4]
// region = this item's plane (H*W*D*4) followed by the tile staging: grad_output rows (tq*D*4) and point records
// (tq*P*16), tq = as many rows as fit (a multiple of kBinThreads/P, at most kBinMaxPpt of them).
// acc_budget = region_bytes - (kBinThreads/P) * (D*4 + P*16), so a binned level always leaves room for one minimal
// tile; bins_cap = acc_budget / (2*D) + 2  >=  (H+1)*(W+1) for every level whose plane fits the budget.
template <typename T, int D, int L, int P>
__global__ void __launch_bounds__(kBinThreads, 1)
msda_bwd_binned(const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi, const float *__restrict__ loc,
                const float *__restrict__ attn, const T *__restrict__ grad_out, float *__restrict__ gv_acc, int N, int S,
                int M, int Lq, int acc_budget, int region_bytes, int bins_cap)
{
    static_assert(D % 4 == 0 && (D / 4) <= 32 && 32 % (D / 4) == 0 && L <= kBinMaxLevels, "unsupported");
    constexpr int LPT = D / 4;                  // lanes per pixel line (one float4 each)
    constexpr int NGROUPS = kBinThreads / LPT;  // bin owners per CTA
    constexpr int LP = L * P;
    constexpr int ROWS_PER_PASS = kBinThreads / P;  // rows whose points of one level fill the CTA once
    static_assert(kBinThreads % P == 0, "P must divide the CTA size");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *acc_s = reinterpret_cast<float *>(smem_raw);
    unsigned *cnt = reinterpret_cast<unsigned *>(smem_raw + region_bytes);
    unsigned *off = cnt + bins_cap;
    __shared__ int sH[L], sW[L], sStart[L];
    __shared__ unsigned warp_sums[32];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    // ---- plan: the binned levels and how the queries are sliced (identical in every CTA) ----------------------
    int lev[L];
    int nb = 0;
#pragma unroll
    for (int l = 0; l < L; ++l)
        if (level_is_binned(sH[l], sW[l], D, acc_budget)) lev[nb++] = l;
    if (nb == 0) return;
    const long long groups = (long long)N * M * nb;
    int slices = (int)((4LL * gridDim.x + groups - 1) / groups);  // aim at >= 4 items per CTA
    const int max_slices = (Lq + ROWS_PER_PASS - 1) / ROWS_PER_PASS;
    slices = max(1, min(slices, max_slices));
    const int rows_per_slice = (Lq + slices - 1) / slices;
    const long long n_items = groups * slices;

    const int t = threadIdx.x;
    const int gid = t / LPT, sub = t % LPT;
    const int MD = M * D;

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        // item -> (n, slice, m, level): the level index moves fastest so the two level-items that read the same
        // grad_output rows run at about the same time (the second read hits L2)
        int li = (int)(item % nb);
        long long rest = item / nb;
        const int m = (int)(rest % M);
        rest /= M;
        const int sl = (int)(rest % slices);
        const int n = (int)(rest / slices);
        int l = 0;
#pragma unroll
        for (int k = 0; k < L; ++k)
            if (k == li) l = lev[k];
        const int H = sH[l], W = sW[l], start = sStart[l];
        const int npix = H * W, W1 = W + 1, nbins = (H + 1) * W1;
        const int qs = sl * rows_per_slice, qe = min(qs + rows_per_slice, Lq);
        // rows per tile: as many as the region holds behind this level's plane
        int tq = (region_bytes - npix * D * 4) / (D * 4 + P * 16);
        tq = min(tq / ROWS_PER_PASS, kBinMaxPpt) * ROWS_PER_PASS;  // >= ROWS_PER_PASS by the definition of acc_budget
        float *g_s = acc_s + npix * D;
        float4 *rec_s = reinterpret_cast<float4 *>(g_s + (size_t)tq * D);

        for (int i = t; i < npix * LPT; i += kBinThreads) reinterpret_cast<float4 *>(acc_s)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = t; i < nbins; i += kBinThreads) cnt[i] = 0;
        __syncthreads();

        for (int q0 = qs; q0 < qe; q0 += tq) {
            const int rows = min(tq, qe - q0);
            // ---- stage grad_output rows of this head (fp32) ------------------------------------------------------
            for (int i = t; i < rows * LPT; i += kBinThreads) {
                const int ql = i / LPT, c = i % LPT;
                const int64_t row = ((int64_t)n * Lq + q0 + ql) * M + m;
                reinterpret_cast<float4 *>(g_s)[i] = load4_as_float<T>(grad_out + row * D + c * 4);
            }
            // ---- (1) resolve and count ------------------------------------------------------------------------------
            int my_bin[kBinMaxPpt];
            unsigned my_rank[kBinMaxPpt];
            float my_a[kBinMaxPpt], my_lh[kBinMaxPpt], my_lw[kBinMaxPpt];
#pragma unroll
            for (int k = 0; k < kBinMaxPpt; ++k) {
                my_bin[k] = -1;
                const int i = t + k * kBinThreads;
                const int ql = i / P, p = i % P;
                if (ql < rows) {
                    const int64_t row = ((int64_t)n * Lq + q0 + ql) * M + m;
                    const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + row * LP + l * P + p);
                    const Cell cell = resolve_cell(xy.x, xy.y, H, W);
                    if (cell.live) {  // a skipped point never reads its weight, like the reference
                        my_a[k] = __ldg(attn + row * LP + l * P + p);
                        my_lh[k] = cell.lh, my_lw[k] = cell.lw;
                        my_bin[k] = (cell.r0 + 1) * W1 + (cell.c0 + 1);
                        my_rank[k] = atomicAdd(&cnt[my_bin[k]], 1u);
                    }
                }
            }
            __syncthreads();
            // ---- (2) counts -> offsets ------------------------------------------------------------------------------
            block_exclusive_scan<true>(cnt, off, nbins, warp_sums);
            // ---- (3) scatter the records into bin order -------------------------------------------------------------
#pragma unroll
            for (int k = 0; k < kBinMaxPpt; ++k)
                if (my_bin[k] >= 0) {
                    const int ql = (t + k * kBinThreads) / P;
                    rec_s[off[my_bin[k]] + my_rank[k]] = make_float4(__int_as_float(ql), my_a[k], my_lh[k], my_lw[k]);
                }
            __syncthreads();
            // ---- (4) owners: four colour phases ---------------------------------------------------------------------
#pragma unroll 1
            for (int colour = 0; colour < 4; ++colour) {
                const int cr = colour >> 1, cc = colour & 1;
                const int nr = (H + 2 - cr) >> 1, nc = (W + 2 - cc) >> 1;  // bin rows rb in [0,H], cols cb in [0,W] of this parity
                for (int j = gid; j < nr * nc; j += NGROUPS) {
                    const int rb = 2 * (j / nc) + cr, cb = 2 * (j % nc) + cc;
                    const int bin = rb * W1 + cb;
                    const unsigned beg = off[bin], end = off[bin + 1];
                    if (beg == end) continue;
                    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
                    for (unsigned i = beg; i < end; ++i) {
                        const float4 r = rec_s[i];
                        const float4 g = *reinterpret_cast<const float4 *>(g_s + __float_as_int(r.x) * D + sub * 4);
                        const float a = r.y, lh = r.z, lw = r.w;
                        const float hh = 1.f - lh, hw = 1.f - lw;
                        const float ah = a * hh, al = a * lh;
                        fma4(a0, ah * hw, g);
                        fma4(a1, ah * lw, g);
                        fma4(a2, al * hw, g);
                        fma4(a3, al * lw, g);
                    }
                    // footprint: pixels (rb-1, cb-1), (rb-1, cb), (rb, cb-1), (rb, cb); taps outside the map are dropped
                    float *p0 = acc_s + ((rb - 1) * W + (cb - 1)) * D + sub * 4;
                    const bool top = rb >= 1, bot = rb < H, lef = cb >= 1, rig = cb < W;
                    if (top && lef) smem_accumulate4(p0, a0);
                    if (top && rig) smem_accumulate4(p0 + D, a1);
                    if (bot && lef) smem_accumulate4(p0 + W * D, a2);
                    if (bot && rig) smem_accumulate4(p0 + W * D + D, a3);
                }
                __syncthreads();
            }
        }

        // ---- flush the plane: one vector red per non-zero chunk ------------------------------------------------------
        float *dst = gv_acc + ((int64_t)n * S * M + m) * D + (int64_t)start * MD;
        for (int i = t; i < npix * LPT; i += kBinThreads) {
            const int pix = i / LPT, c = i % LPT;
            const float4 v = reinterpret_cast<const float4 *>(acc_s)[i];
            if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
                red_add_f32x4(dst + (int64_t)pix * MD + c * 4, v.x, v.y, v.z, v.w);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// msda_bwd_owned -- grad_value for SPARSE problems (few queries over a large map: GRIT's decoder, Lq = 150).
//
// There the row-parallel backward spends its time on memory that no sample touches: grad_value is dense, so it is
// zero-filled (and, for bf16, accumulated in an fp32 workspace that is zero-filled, read back and folded) although
// the 9 600 taps of an (image, head) pair land on a fraction of its 22 223 pixels.  Here every pixel line of
// grad_value is written exactly ONCE, by the lane group that owns it, with a plain store -- no zero-fill, no
// workspace, no fold, no floating-point atomics:
//   * a work item is (image, head, chunk of the pixel range); the CTA resolves ALL taps of that (image, head) --
//     cheap, there are few -- and keeps those that land in its chunk;
//   * counting sort by pixel in shared memory (native integer atomics): count, block scan, scatter {row, weight};
//   * owners walk the chunk's pixels: sum weight * grad_output[row] over the pixel's records (grad_output rows staged
//     in shared memory as fp32), then store the line in the output dtype (bf16 included: fp32 accumulation, one
//     rounding), or read-add-store when accumulating into a caller's non-zero grad_value / across row tiles.
// Queries are walked in tiles of `tq` rows when they do not all fit in shared memory.
// Dynamic shared memory: [gtile : tq*D*4][rec : tq*L*P*4 * 8][cur : (chunk_pixels + 1) * 4].
template <typename T>
__device__ __forceinline__ void store4_from_float(T *p, const float4 &v);
template <>
__device__ __forceinline__ void store4_from_float<float>(float *p, const float4 &v)
{
    *reinterpret_cast<float4 *>(p) = v;
}
template <>
__device__ __forceinline__ void store4_from_float<__nv_bfloat16>(__nv_bfloat16 *p, const float4 &v)
{
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const unsigned *>(&lo);
    u.y = *reinterpret_cast<const unsigned *>(&hi);
    *reinterpret_cast<uint2 *>(p) = u;
}
template <typename T>
__device__ __forceinline__ float4 load4_plain(const T *p);
template <>
__device__ __forceinline__ float4 load4_plain<float>(const float *p)
{
    return *reinterpret_cast<const float4 *>(p);
}
template <>
__device__ __forceinline__ float4 load4_plain<__nv_bfloat16>(const __nv_bfloat16 *p)
{
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                       __uint_as_float(v.y & 0xffff0000u));
}

template <typename T, int D, int L, int P>
__global__ void __launch_bounds__(kBinThreads, 1)
msda_bwd_owned(const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi, const float *__restrict__ loc,
               const float *__restrict__ attn, const T *__restrict__ grad_out, T *__restrict__ grad_value, int N, int S,
               int M, int Lq, int chunks, int chunk_pixels, int tq, int accumulate)
{
    static_assert(D % 4 == 0 && (D / 4) <= 32 && 32 % (D / 4) == 0 && L <= kBinMaxLevels, "unsupported");
    constexpr int LPT = D / 4;
    constexpr int NGROUPS = kBinThreads / LPT;
    constexpr int LP = L * P;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *g_s = reinterpret_cast<float *>(smem_raw);
    float2 *rec_s = reinterpret_cast<float2 *>(smem_raw + (size_t)tq * D * 4);
    unsigned *cur = reinterpret_cast<unsigned *>(smem_raw + (size_t)tq * D * 4 + (size_t)tq * LP * 4 * 8);
    __shared__ int sH[L], sW[L], sStart[L];
    __shared__ unsigned warp_sums[32];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    const int t = threadIdx.x;
    const int gid = t / LPT, sub = t % LPT;
    const int MD = M * D;
    const long long n_items = (long long)N * M * chunks;

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int c = (int)(item % chunks);
        const int m = (int)((item / chunks) % M);
        const int n = (int)(item / ((long long)chunks * M));
        const int p0 = c * chunk_pixels, p1 = min(p0 + chunk_pixels, S);
        const int np = p1 - p0;
        if (np <= 0) continue;
        T *gv = grad_value + ((int64_t)n * S * M + m) * D + (int64_t)p0 * MD + sub * 4;

        for (int q0 = 0, tile = 0; q0 < Lq || tile == 0; q0 += tq, ++tile) {
            const int rows = max(0, min(tq, Lq - q0));
            for (int i = t; i < np + 1; i += kBinThreads) cur[i] = 0;
            for (int i = t; i < rows * LPT; i += kBinThreads) {
                const int ql = i / LPT, ch = i % LPT;
                const int64_t row = ((int64_t)n * Lq + q0 + ql) * M + m;
                reinterpret_cast<float4 *>(g_s)[i] = load4_as_float<T>(grad_out + row * D + ch * 4);
            }
            __syncthreads();
            // ---- pass A: count the taps that land in this chunk ------------------------------------------------------
            for (int i = t; i < rows * LP; i += kBinThreads) {
                const int ql = i / LP, pt = i % LP, l = pt / P;
                const int64_t row = ((int64_t)n * Lq + q0 + ql) * M + m;
                const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + row * LP + pt);
                const int H = sH[l], W = sW[l];
                const Cell cell = resolve_cell(xy.x, xy.y, H, W);
                if (!cell.live) continue;
                const int base = sStart[l] + cell.r0 * W + cell.c0 - p0;
                const bool top = cell.r0 >= 0, bot = cell.r0 + 1 < H, lef = cell.c0 >= 0, rig = cell.c0 + 1 < W;
                if (top && lef && (unsigned)base < (unsigned)np) atomicAdd(&cur[base], 1u);
                if (top && rig && (unsigned)(base + 1) < (unsigned)np) atomicAdd(&cur[base + 1], 1u);
                if (bot && lef && (unsigned)(base + W) < (unsigned)np) atomicAdd(&cur[base + W], 1u);
                if (bot && rig && (unsigned)(base + W + 1) < (unsigned)np) atomicAdd(&cur[base + W + 1], 1u);
            }
            __syncthreads();
            block_exclusive_scan<false>(cur, cur, np, warp_sums);
            // ---- pass B: scatter {row, weight}; cur[p] ends at the end of pixel p's segment ----------------------------
            for (int i = t; i < rows * LP; i += kBinThreads) {
                const int ql = i / LP, pt = i % LP, l = pt / P;
                const int64_t row = ((int64_t)n * Lq + q0 + ql) * M + m;
                const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + row * LP + pt);
                const int H = sH[l], W = sW[l];
                const Cell cell = resolve_cell(xy.x, xy.y, H, W);
                if (!cell.live) continue;
                const float a = __ldg(attn + row * LP + pt);
                const int base = sStart[l] + cell.r0 * W + cell.c0 - p0;
                const bool top = cell.r0 >= 0, bot = cell.r0 + 1 < H, lef = cell.c0 >= 0, rig = cell.c0 + 1 < W;
                const float hh = 1.f - cell.lh, hw = 1.f - cell.lw;
                const float ah = a * hh, al = a * cell.lh;
                const float qf = __int_as_float(ql);
                if (top && lef && (unsigned)base < (unsigned)np) rec_s[atomicAdd(&cur[base], 1u)] = make_float2(qf, ah * hw);
                if (top && rig && (unsigned)(base + 1) < (unsigned)np)
                    rec_s[atomicAdd(&cur[base + 1], 1u)] = make_float2(qf, ah * cell.lw);
                if (bot && lef && (unsigned)(base + W) < (unsigned)np)
                    rec_s[atomicAdd(&cur[base + W], 1u)] = make_float2(qf, al * hw);
                if (bot && rig && (unsigned)(base + W + 1) < (unsigned)np)
                    rec_s[atomicAdd(&cur[base + W + 1], 1u)] = make_float2(qf, al * cell.lw);
            }
            __syncthreads();
            // ---- owners: one lane group per pixel line ---------------------------------------------------------------------
            const bool first = tile == 0;
            for (int p = gid; p < np; p += NGROUPS) {
                const unsigned beg = p > 0 ? cur[p - 1] : 0u, end = cur[p];
                if (beg == end && (!first || accumulate)) continue;  // nothing to add to what is already there
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (unsigned i = beg; i < end; ++i) {
                    const float2 r = rec_s[i];
                    fma4(acc, r.y, *reinterpret_cast<const float4 *>(g_s + __float_as_int(r.x) * D + sub * 4));
                }
                T *dst = gv + (int64_t)p * MD;
                if (!first || accumulate) {
                    const float4 old = load4_plain<T>(dst);
                    acc.x += old.x, acc.y += old.y, acc.z += old.z, acc.w += old.w;
                }
                store4_from_float<T>(dst, acc);
            }
            __syncthreads();
        }
    }
}

}  // namespace msda
