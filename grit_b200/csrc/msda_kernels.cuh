// msda_kernels.cuh -- sm_100a device code for multi-scale deformable attention.
//
// Replaces the reference's models/ops/src/cuda/ms_deform_im2col_cuda.cuh (1 forward + 6 backward
// kernels, one thread per output scalar, scalar taps, scalar atomics).  Design here:
//
//   * one WARP per (image, query, head) row; the D channels of a head are spread over `LPT` lanes,
//     each lane owning one 16-byte chunk (4 fp32 / 8 bf16), so a bilinear tap is one coalesced
//     vector load per lane and a 128-byte line per tap at D=32 fp32;
//   * the 32/LPT lane groups of the warp work on different sample points at the same time, the
//     point loop is fully unrolled so all tap loads of a row are in flight together;
//   * forward: group partial sums are combined with xor-shuffles, lanes of group 0 store the row;
//   * backward: grad_value goes out as vector reductions (REDG.E.ADD.F32x4), the three per-point
//     scalars (grad_attn, grad_loc.x, grad_loc.y) are reduced over the LPT lanes by shuffles and
//     written once per row as one coalesced store each -- no shared memory, no __syncthreads in
//     the loop, no zero-fill of grad_loc / grad_attn needed;
//   * spatial_shapes / level_start_index stay on the device (no host sync): each CTA copies them
//     into shared memory once.
//
// Semantics (validity window, -0.5 shift, per-tap zero padding, gradient formulas) follow
// ms_deform_im2col_cuda.cuh:33-159 and :272-296; see oracle/msda_oracle_impl.h for the restatement
// the tests compare against.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

// ------------------------------------------------------------------------------------------------
// 16-byte lane chunks
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Chunk;

template <>
struct Chunk<float> {
    static constexpr int E = 4;
    __device__ __forceinline__ static void load(const float *p, float (&r)[4])
    {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        r[0] = v.x, r[1] = v.y, r[2] = v.z, r[3] = v.w;
    }
    // same load as an ordered (volatile) PTX statement: the compiler may not sink it below later volatile asm,
    // which is how the hoisted kernels keep a whole row's taps in flight
    __device__ __forceinline__ static void load_ordered(const float *p, float (&r)[4])
    {
        asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3])
                     : "l"(p));
    }
    __device__ __forceinline__ static void load_shared(const float *p, float (&r)[4])
    {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        r[0] = v.x, r[1] = v.y, r[2] = v.z, r[3] = v.w;
    }
    __device__ __forceinline__ static void store(float *p, const float (&r)[4])
    {
        *reinterpret_cast<float4 *>(p) = make_float4(r[0], r[1], r[2], r[3]);
    }
};

template <>
struct Chunk<__nv_bfloat16> {
    static constexpr int E = 8;
    __device__ __forceinline__ static void unpack(const uint4 &v, float (&r)[8])
    {
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[2 * i] = __uint_as_float(w[i] << 16);
            r[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ static void load(const __nv_bfloat16 *p, float (&r)[8])
    {
        unpack(__ldg(reinterpret_cast<const uint4 *>(p)), r);
    }
    __device__ __forceinline__ static void load_ordered(const __nv_bfloat16 *p, float (&r)[8])
    {
        uint4 v;
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        unpack(v, r);
    }
    __device__ __forceinline__ static void load_shared(const __nv_bfloat16 *p, float (&r)[8])
    {
        unpack(*reinterpret_cast<const uint4 *>(p), r);
    }
    __device__ __forceinline__ static void store(__nv_bfloat16 *p, const float (&r)[8])
    {
        uint4 v;
        unsigned *w = reinterpret_cast<unsigned *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
            w[i] = *reinterpret_cast<const unsigned *>(&h);
        }
        *reinterpret_cast<uint4 *>(p) = v;
    }
};

// 8-byte bf16 chunk (4 channels per lane).  The bf16 backward uses it so that a lane owns 4 channels = ONE 16-byte
// fp32 `red` per tap and a tap's 128-byte fp32 gradient line leaves the SM as one request (with the 16-byte bf16
// chunk every lane would issue two half-sector reds per tap: measured 1.6x slower).
struct ChunkBf16x4 {
    static constexpr int E = 4;
    __device__ __forceinline__ static void load(const __nv_bfloat16 *p, float (&r)[4])
    {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        r[0] = __uint_as_float(v.x << 16), r[1] = __uint_as_float(v.x & 0xffff0000u);
        r[2] = __uint_as_float(v.y << 16), r[3] = __uint_as_float(v.y & 0xffff0000u);
    }
};

// fire-and-forget vector reduction into global memory (REDG.E.ADD.F32x4 on sm_90+)
__device__ __forceinline__ void red_add_f32x4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

template <int E>
__device__ __forceinline__ void red_add_chunk(float *p, const float (&g)[E], float s)
{
#pragma unroll
    for (int i = 0; i < E; i += 4) red_add_f32x4(p + i, s * g[i], s * g[i + 1], s * g[i + 2], s * g[i + 3]);
}

// One sample point resolved against its level: tap offsets (in pixels), validity and weights.
struct Taps {
    int pix;          // start + r0*W + c0  (pixel index of the top-left tap inside the image)
    int W;            // row pitch in pixels
    bool live;        // inside the (-1, size) window; dead points contribute nothing and get zero gradients
    bool tl, tr, bl, br;
    float lh, lw, hh, hw;
};

__device__ __forceinline__ Taps resolve_taps(float x, float y, int H, int W, int start)
{
    Taps t;
    float h_im = fmaf(y, (float)H, -0.5f);
    float w_im = fmaf(x, (float)W, -0.5f);
    t.live = h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;  // false for NaN
    if (!t.live) h_im = w_im = 0.f;  // keeps every derived quantity finite; all four taps end up invalid
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int r0 = (int)hf, c0 = (int)wf;
    t.lh = h_im - hf, t.lw = w_im - wf;
    t.hh = 1.f - t.lh, t.hw = 1.f - t.lw;
    const bool top = t.live && r0 >= 0, bot = t.live && r0 + 1 < H;
    const bool lef = c0 >= 0, rig = c0 + 1 < W;
    t.tl = top && lef, t.tr = top && rig, t.bl = bot && lef, t.br = bot && rig;
    t.pix = start + r0 * W + c0;
    t.W = W;
    return t;
}

template <int L>
__device__ __forceinline__ void stage_levels(const int64_t *shapes, const int64_t *lsi, int (&sH)[L], int (&sW)[L],
                                             int (&sStart)[L])
{
    if (threadIdx.x < L) {
        sH[threadIdx.x] = (int)shapes[2 * threadIdx.x];
        sW[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
        sStart[threadIdx.x] = (int)lsi[threadIdx.x];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Specialised forward: compile-time D, L, P; T in {float, bf16}; loc/attn fp32.
// ------------------------------------------------------------------------------------------------
template <typename T, int D, int L, int P, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
msda_fwd_vec(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
             const float *__restrict__ loc, const float *__restrict__ attn, T *__restrict__ out, int S, int M, int Lq,
             int64_t rows)
{
    constexpr int E = Chunk<T>::E;
    constexpr int LPT = D / E;   // lanes per tap
    constexpr int G = 32 / LPT;  // sample points in flight per warp
    constexpr int LP = L * P;
    constexpr int PPG = LP / G;  // points each lane group walks through
    static_assert(D % E == 0 && 32 % LPT == 0 && LP % G == 0, "unsupported specialisation");

    __shared__ int sH[L], sW[L], sStart[L];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    const int lane = threadIdx.x & 31;
    const int g = lane / LPT, sub = lane % LPT;
    const unsigned urow = blockIdx.x * WARPS + (threadIdx.x >> 5);  // (b*Lq + q)*M + m, < 2^31 (host-checked)
    if (urow >= (unsigned)rows) return;
    const int64_t row = urow;
    const int m = (int)(urow % (unsigned)M);
    const int64_t b = urow / ((unsigned)M * (unsigned)Lq);
    const int MD = M * D;
    const T *vbase = value + (b * S * M + m) * (int64_t)D + sub * E;
    const float2 *lrow = reinterpret_cast<const float2 *>(loc) + row * LP;
    const float *arow = attn + row * LP;

    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;

#pragma unroll
    for (int it = 0; it < PPG; ++it) {
        const int pt = it * G + g;
        const int l = pt / P;
        const float2 xy = __ldg(lrow + pt);
        const Taps t = resolve_taps(xy.x, xy.y, sH[l], sW[l], sStart[l]);
        const float a = t.live ? __ldg(arow + pt) : 0.f;  // the reference never reads attn of a skipped point
        const T *p0 = vbase + (int64_t)t.pix * MD;
        const T *p1 = p0 + (int64_t)t.W * MD;
        float v0[E], v1[E], v2[E], v3[E];
#pragma unroll
        for (int e = 0; e < E; ++e) v0[e] = v1[e] = v2[e] = v3[e] = 0.f;
        if (t.tl) Chunk<T>::load(p0, v0);
        if (t.tr) Chunk<T>::load(p0 + MD, v1);
        if (t.bl) Chunk<T>::load(p1, v2);
        if (t.br) Chunk<T>::load(p1 + MD, v3);
        const float w0 = a * t.hh * t.hw, w1 = a * t.hh * t.lw, w2 = a * t.lh * t.hw, w3 = a * t.lh * t.lw;
#pragma unroll
        for (int e = 0; e < E; ++e)
            acc[e] = fmaf(w0, v0[e], fmaf(w1, v1[e], fmaf(w2, v2[e], fmaf(w3, v3[e], acc[e]))));
    }

#pragma unroll
    for (int off = LPT; off < 32; off <<= 1) {
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
    }
    if (g == 0) Chunk<T>::store(out + row * D + sub * E, acc);
}

// ------------------------------------------------------------------------------------------------
// Specialised backward.  grad_value accumulates in fp32 (`gv_acc`): for T=float that IS grad_value,
// for T=bf16 it is the caller's fp32 workspace which msda_f32_to_bf16_accumulate folds back.
// ------------------------------------------------------------------------------------------------
template <typename T, int D, int L, int P, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
msda_bwd_vec(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
             const float *__restrict__ loc, const float *__restrict__ attn, const T *__restrict__ grad_out,
             float *__restrict__ gv_acc, float *__restrict__ grad_loc, float *__restrict__ grad_attn, int S, int M,
             int Lq, int64_t rows)
{
    constexpr int E = Chunk<T>::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    constexpr int PPG = LP / G;
    static_assert(D % E == 0 && 32 % LPT == 0 && LP % G == 0 && PPG <= LPT, "unsupported specialisation");

    __shared__ int sH[L], sW[L], sStart[L];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    const int lane = threadIdx.x & 31;
    const int g = lane / LPT, sub = lane % LPT;
    const unsigned urow = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (urow >= (unsigned)rows) return;
    const int64_t row = urow;
    const int m = (int)(urow % (unsigned)M);
    const int64_t b = urow / ((unsigned)M * (unsigned)Lq);
    const int MD = M * D;
    const int64_t img = (b * S * M + m) * (int64_t)D + sub * E;
    const T *vbase = value + img;
    float *gbase = gv_acc + img;
    const float2 *lrow = reinterpret_cast<const float2 *>(loc) + row * LP;
    const float *arow = attn + row * LP;

    float go[E];
    Chunk<T>::load(grad_out + row * D + sub * E, go);

    float keep_a = 0.f, keep_x = 0.f, keep_y = 0.f;  // results of point (sub*G + g), for sub < PPG

#pragma unroll
    for (int it = 0; it < PPG; ++it) {
        const int pt = it * G + g;
        const int l = pt / P;
        const float2 xy = __ldg(lrow + pt);
        const int H = sH[l], W = sW[l];
        const Taps t = resolve_taps(xy.x, xy.y, H, W, sStart[l]);
        const float a = t.live ? __ldg(arow + pt) : 0.f;
        const int64_t o0 = (int64_t)t.pix * MD, o1 = o0 + (int64_t)W * MD;
        float v0[E], v1[E], v2[E], v3[E];
#pragma unroll
        for (int e = 0; e < E; ++e) v0[e] = v1[e] = v2[e] = v3[e] = 0.f;
        if (t.tl) Chunk<T>::load(vbase + o0, v0);
        if (t.tr) Chunk<T>::load(vbase + o0 + MD, v1);
        if (t.bl) Chunk<T>::load(vbase + o1, v2);
        if (t.br) Chunk<T>::load(vbase + o1 + MD, v3);

        if (t.tl) red_add_chunk<E>(gbase + o0, go, a * t.hh * t.hw);
        if (t.tr) red_add_chunk<E>(gbase + o0 + MD, go, a * t.hh * t.lw);
        if (t.bl) red_add_chunk<E>(gbase + o1, go, a * t.lh * t.hw);
        if (t.br) red_add_chunk<E>(gbase + o1 + MD, go, a * t.lh * t.lw);

        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;  // <grad_out, tap_i> over this lane's channels
#pragma unroll
        for (int e = 0; e < E; ++e) {
            d0 = fmaf(go[e], v0[e], d0);
            d1 = fmaf(go[e], v1[e], d1);
            d2 = fmaf(go[e], v2[e], d2);
            d3 = fmaf(go[e], v3[e], d3);
        }
        float s_a = t.hh * (t.hw * d0 + t.lw * d1) + t.lh * (t.hw * d2 + t.lw * d3);
        float s_x = t.hh * (d1 - d0) + t.lh * (d3 - d2);
        float s_y = t.hw * (d2 - d0) + t.lw * (d3 - d1);
#pragma unroll
        for (int off = 1; off < LPT; off <<= 1) {
            s_a += __shfl_xor_sync(0xffffffffu, s_a, off);
            s_x += __shfl_xor_sync(0xffffffffu, s_x, off);
            s_y += __shfl_xor_sync(0xffffffffu, s_y, off);
        }
        if (sub == it) {
            keep_a = s_a;
            keep_x = (float)W * a * s_x;
            keep_y = (float)H * a * s_y;
        }
    }
    if (sub < PPG) {
        const int pt = sub * G + g;
        reinterpret_cast<float2 *>(grad_loc)[row * LP + pt] = make_float2(keep_x, keep_y);
        grad_attn[row * LP + pt] = keep_a;
    }
}

// ------------------------------------------------------------------------------------------------
// Generic kernels: any D, L, P; T in {float, double, bf16}; C = compute/loc type (float or double).
// Warp per row, lanes stride over channels, scalar taps.  This is the path gradcheck (fp64) and
// odd channel counts (30, 71, ...) take.
// ------------------------------------------------------------------------------------------------
template <typename C, typename T>
__device__ __forceinline__ C to_c(T v)
{
    return (C)v;
}
template <>
__device__ __forceinline__ float to_c<float, __nv_bfloat16>(__nv_bfloat16 v)
{
    return __bfloat162float(v);
}
template <typename T, typename C>
__device__ __forceinline__ T from_c(C v)
{
    return (T)v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_c<__nv_bfloat16, float>(float v)
{
    return __float2bfloat16_rn(v);
}

template <typename C>
struct TapsG {
    int64_t pix;
    int W;
    bool live;
    bool tl, tr, bl, br;
    C lh, lw, hh, hw;
};

template <typename C>
__device__ __forceinline__ TapsG<C> resolve_taps_g(C x, C y, int H, int W, int start)
{
    TapsG<C> t;
    C h_im = y * (C)H - (C)0.5;
    C w_im = x * (C)W - (C)0.5;
    t.live = h_im > (C)-1 && w_im > (C)-1 && h_im < (C)H && w_im < (C)W;
    if (!t.live) h_im = w_im = (C)0;
    const C hf = floor(h_im), wf = floor(w_im);
    const int r0 = (int)hf, c0 = (int)wf;
    t.lh = h_im - hf, t.lw = w_im - wf;
    t.hh = (C)1 - t.lh, t.hw = (C)1 - t.lw;
    const bool top = t.live && r0 >= 0, bot = t.live && r0 + 1 < H;
    const bool lef = c0 >= 0, rig = c0 + 1 < W;
    t.tl = top && lef, t.tr = top && rig, t.bl = bot && lef, t.br = bot && rig;
    t.pix = (int64_t)start + (int64_t)r0 * W + c0;
    t.W = W;
    return t;
}

template <typename T, typename C, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
msda_fwd_generic(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
                 const C *__restrict__ loc, const C *__restrict__ attn, T *__restrict__ out, int64_t S, int M, int D,
                 int L, int64_t Lq, int P, int64_t rows)
{
    extern __shared__ int s_meta[];  // H[L], W[L], start[L]
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        s_meta[i] = (int)shapes[2 * i];
        s_meta[L + i] = (int)shapes[2 * i + 1];
        s_meta[2 * L + i] = (int)lsi[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int m = (int)(row % M);
    const int64_t b = row / ((int64_t)M * Lq);
    const int64_t MD = (int64_t)M * D;
    const T *vimg = value + (b * S * M + m) * (int64_t)D;
    const C *lrow = loc + row * L * P * 2;
    const C *arow = attn + row * L * P;

    for (int c = lane; c < D; c += 32) {
        C acc = 0;
        for (int l = 0; l < L; ++l) {
            const int H = s_meta[l], W = s_meta[L + l], start = s_meta[2 * L + l];
            for (int p = 0; p < P; ++p) {
                const int k = l * P + p;
                const TapsG<C> t = resolve_taps_g<C>(lrow[2 * k], lrow[2 * k + 1], H, W, start);
                const C a = t.live ? arow[k] : (C)0;
                const T *p0 = vimg + t.pix * MD + c;
                const T *p1 = p0 + (int64_t)W * MD;
                const C v0 = t.tl ? to_c<C, T>(p0[0]) : (C)0;
                const C v1 = t.tr ? to_c<C, T>(p0[MD]) : (C)0;
                const C v2 = t.bl ? to_c<C, T>(p1[0]) : (C)0;
                const C v3 = t.br ? to_c<C, T>(p1[MD]) : (C)0;
                acc += a * (t.hh * (t.hw * v0 + t.lw * v1) + t.lh * (t.hw * v2 + t.lw * v3));
            }
        }
        out[row * D + c] = from_c<T, C>(acc);
    }
}

// A = accumulation element of gv_acc: double for fp64, float otherwise (bf16 goes through the workspace).
template <typename C>
__device__ __forceinline__ void acc_add(C *p, C v, float)
{
    atomicAdd(p, v);
}
// deterministic mode: power-of-two scaled int64 fixed point (see AccFix64 in msda_kernels_v5.cuh)
__device__ __forceinline__ void acc_add(unsigned long long *p, float v, float scale)
{
    const long long q = __float2ll_rn(v * scale);
    if (q != 0) atomicAdd(p, (unsigned long long)q);
}

template <typename T, typename C, typename A, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
msda_bwd_generic(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
                 const C *__restrict__ loc, const C *__restrict__ attn, const T *__restrict__ grad_out,
                 A *__restrict__ gv_acc, const float *__restrict__ det_scale, C *__restrict__ grad_loc,
                 C *__restrict__ grad_attn, int64_t S, int M, int D, int L, int64_t Lq, int P, int64_t rows)
{
    const float fx_scale = det_scale ? __ldg(det_scale) : 1.f;
    extern __shared__ int s_meta[];
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        s_meta[i] = (int)shapes[2 * i];
        s_meta[L + i] = (int)shapes[2 * i + 1];
        s_meta[2 * L + i] = (int)lsi[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int m = (int)(row % M);
    const int64_t b = row / ((int64_t)M * Lq);
    const int64_t MD = (int64_t)M * D;
    const int64_t img = (b * S * M + m) * (int64_t)D;
    const C *lrow = loc + row * L * P * 2;
    const C *arow = attn + row * L * P;
    const T *grow = grad_out + row * D;

    for (int l = 0; l < L; ++l) {
        const int H = s_meta[l], W = s_meta[L + l], start = s_meta[2 * L + l];
        for (int p = 0; p < P; ++p) {
            const int k = l * P + p;
            const TapsG<C> t = resolve_taps_g<C>(lrow[2 * k], lrow[2 * k + 1], H, W, start);
            const C a = t.live ? arow[k] : (C)0;
            const C w0 = t.hh * t.hw, w1 = t.hh * t.lw, w2 = t.lh * t.hw, w3 = t.lh * t.lw;
            const int64_t o0 = img + t.pix * MD, o1 = o0 + (int64_t)W * MD;
            C s_a = 0, s_x = 0, s_y = 0;
            for (int c = lane; c < D; c += 32) {
                const C gch = to_c<C, T>(grow[c]);
                const C v0 = t.tl ? to_c<C, T>(value[o0 + c]) : (C)0;
                const C v1 = t.tr ? to_c<C, T>(value[o0 + MD + c]) : (C)0;
                const C v2 = t.bl ? to_c<C, T>(value[o1 + c]) : (C)0;
                const C v3 = t.br ? to_c<C, T>(value[o1 + MD + c]) : (C)0;
                const C ga = gch * a;
                if (t.tl) acc_add(gv_acc + o0 + c, w0 * ga, fx_scale);
                if (t.tr) acc_add(gv_acc + o0 + MD + c, w1 * ga, fx_scale);
                if (t.bl) acc_add(gv_acc + o1 + c, w2 * ga, fx_scale);
                if (t.br) acc_add(gv_acc + o1 + MD + c, w3 * ga, fx_scale);
                s_a += gch * (w0 * v0 + w1 * v1 + w2 * v2 + w3 * v3);
                s_x += gch * (t.hh * (v1 - v0) + t.lh * (v3 - v2));
                s_y += gch * (t.hw * (v2 - v0) + t.lw * (v3 - v1));
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                s_a += __shfl_xor_sync(0xffffffffu, s_a, off);
                s_x += __shfl_xor_sync(0xffffffffu, s_x, off);
                s_y += __shfl_xor_sync(0xffffffffu, s_y, off);
            }
            if (lane == 0) {
                grad_attn[row * L * P + k] = s_a;
                grad_loc[(row * L * P + k) * 2] = (C)W * a * s_x;
                grad_loc[(row * L * P + k) * 2 + 1] = (C)H * a * s_y;
            }
        }
    }
}

// bf16 backward epilogue: grad_value(bf16) = [grad_value(bf16) +] workspace(fp32)
__global__ void msda_fold_workspace_bf16(const float *__restrict__ ws, __nv_bfloat16 *__restrict__ gv, int64_t n,
                                         int accumulate)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
        if (i + 8 <= n) {
            const float4 a = *reinterpret_cast<const float4 *>(ws + i);
            const float4 b = *reinterpret_cast<const float4 *>(ws + i + 4);
            float r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            if (accumulate) {
                float old[8];
                Chunk<__nv_bfloat16>::load(gv + i, old);
#pragma unroll
                for (int e = 0; e < 8; ++e) r[e] += old[e];
            }
            Chunk<__nv_bfloat16>::store(gv + i, r);
        } else {
            for (int64_t j = i; j < n; ++j) {
                float r = ws[j];
                if (accumulate) r += __bfloat162float(gv[j]);
                gv[j] = __float2bfloat16_rn(r);
            }
        }
    }
}

}  // namespace msda
