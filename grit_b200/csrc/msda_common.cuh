// msda_common.cuh -- device helpers shared by every kernel file (sm_100a).
//
//   * 16-byte lane chunks (4 fp32 / 8 bf16) and the vector reduction `red.global.add.v4.f32`;
//   * tap resolution with the reference's semantics (ms_deform_im2col_cuda.cuh:33-84, 272-296): pixel coordinate
//     loc*size - 0.5, a point counts only inside the open window (-1, size), each tap is zero outside the map;
//   * the "resolve once" record a resolver lane publishes for its sample point, and the halving shuffle reduction
//     the backward kernels use for their three per-point scalars.
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
    // streaming read of data that is dead after this use: evict-first in L1 and L2 (LDG.E.EF)
    __device__ __forceinline__ static void load_stream(const float *p, float (&r)[4])
    {
        asm volatile("ld.global.cs.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3])
                     : "l"(p));
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
    __device__ __forceinline__ static void store(__nv_bfloat16 *p, const float (&r)[4])
    {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(r[0], r[1]), hi = __floats2bfloat162_rn(r[2], r[3]);
        *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<const unsigned *>(&lo),
                                                   *reinterpret_cast<const unsigned *>(&hi));
    }
    __device__ __forceinline__ static void load_stream(const __nv_bfloat16 *p, float (&r)[4])
    {
        uint2 v;
        asm volatile("ld.global.cs.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
        r[0] = __uint_as_float(v.x << 16), r[1] = __uint_as_float(v.x & 0xffff0000u);
        r[2] = __uint_as_float(v.y << 16), r[3] = __uint_as_float(v.y & 0xffff0000u);
    }
    __device__ __forceinline__ static void load_shared(const __nv_bfloat16 *p, float (&r)[4])
    {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
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

// Hide how a pointer was computed from the optimiser.  The row kernels add a 32-bit element offset per tap to a per-lane
// base pointer; when ptxas can see that the base ends in `+ lane_chunk` it re-derives every tap address from scratch
// (LOP3 + SHF + LEA + LEA.HI.X per tap).  With an opaque base each tap address is one IMAD.WIDE.
template <typename T>
__device__ __forceinline__ const T *opaque_ptr(const T *p)
{
    asm("" : "+l"(p));
    return p;
}
template <typename T>
__device__ __forceinline__ T *opaque_ptr(T *p)
{
    asm("" : "+l"(p));
    return p;
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

// What a resolver lane publishes for its point.
struct Resolved {
    int pm;             // (pixel index of the top-left tap) * 16 + 4-bit tap validity mask (tl=1, tr=2, bl=4, br=8)
    float a, lh, lw;    // attention weight (0 for a dead point) and the fractional offsets
};

__device__ __forceinline__ Resolved resolve_point(float x, float y, int H, int W, int start, const float *attn_ptr)
{
    const Taps t = resolve_taps(x, y, H, W, start);
    Resolved r;
    const int mask = (t.tl ? 1 : 0) | (t.tr ? 2 : 0) | (t.bl ? 4 : 0) | (t.br ? 8 : 0);
    r.pm = t.pix * 16 + mask;
    r.a = t.live ? __ldg(attn_ptr) : 0.f;
    r.lh = t.lh, r.lw = t.lw;
    return r;
}

// same, with the attention weight already in a register (software-prefetched by the persistent kernels)
__device__ __forceinline__ Resolved resolve_point_v(float x, float y, int H, int W, int start, float attn_raw)
{
    const Taps t = resolve_taps(x, y, H, W, start);
    Resolved r;
    const int mask = (t.tl ? 1 : 0) | (t.tr ? 2 : 0) | (t.bl ? 4 : 0) | (t.br ? 8 : 0);
    r.pm = t.pix * 16 + mask;
    r.a = t.live ? attn_raw : 0.f;  // a skipped point ignores its weight (NaN included), like the reference
    r.lh = t.lh, r.lw = t.lw;
    return r;
}

// Halving reduction of NV = 3*PPG values over the LPT lanes of a group.  On return lane `sub` holds, in
// vals[0..2], the group-wide sums of iteration  it = sub / (LPT / PPG)  (all lanes of that sub-range agree).
template <int PPG, int LPT>
__device__ __forceinline__ void group_reduce3(float (&vals)[3 * PPG], int sub)
{
    int n = PPG;
#pragma unroll
    for (int width = LPT / 2; width >= 1; width >>= 1) {
        if (n > 1) {
            const int half = 3 * n / 2;
            const bool upper = (sub & width) != 0;
#pragma unroll
            for (int i = 0; i < 3 * PPG / 2; ++i) {
                if (i < half) {
                    const float send = upper ? vals[i] : vals[i + half];
                    const float keep = upper ? vals[i + half] : vals[i];
                    vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, width);
                }
            }
            n >>= 1;
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], width);
        }
    }
}

}  // namespace msda
