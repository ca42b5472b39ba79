// msda_kernels_v2.cuh -- second-generation specialised kernels (sm_100a).
//
// Same warp-per-row / lane-group-per-tap layout as msda_kernels.cuh, with three changes that the
// first ncu capture (profiles/r01_*.txt) asked for:
//   1. every sample point is resolved ONCE per warp (lane p resolves point p: coordinates, window
//      test, floor, weights, tap mask) and the result is broadcast to the lane group that gathers
//      it with a handful of shuffles, instead of all LPT lanes of a group redoing the same scalar
//      math -- the v1 forward was 64 % issue-bound at 459 instructions per row;
//   2. the backward reduces its 3*PPG per-point scalars with a halving ("transpose") reduction:
//      3*PPG + ... shuffles instead of 3*PPG*log2(LPT);
//   3. optional HEAD_MAJOR row order: consecutive warps walk the queries of ONE (image, head) pair,
//      so the SM's L1 and the L2 see one head's value planes at a time (coarse levels become
//      L1-resident) instead of all M heads interleaved.
#pragma once

#include "msda_kernels.cuh"

namespace msda {

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

template <bool HEAD_MAJOR>
__device__ __forceinline__ void decode_row(unsigned urow, int M, int Lq, int64_t &b, int &m, int64_t &row)
{
    if (HEAD_MAJOR) {  // urow = (b*M + m)*Lq + q
        const unsigned bm = urow / (unsigned)Lq;
        const unsigned q = urow - bm * (unsigned)Lq;
        b = bm / (unsigned)M;
        m = (int)(bm - (unsigned)b * (unsigned)M);
        row = ((int64_t)b * Lq + q) * M + m;
    } else {  // urow = (b*Lq + q)*M + m, which is the memory row itself
        row = urow;
        m = (int)(urow % (unsigned)M);
        b = urow / ((unsigned)M * (unsigned)Lq);
    }
}

template <typename T, int D, int L, int P, int WARPS, bool HEAD_MAJOR>
__global__ void __launch_bounds__(WARPS * 32)
msda_fwd_v2(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
            const float *__restrict__ loc, const float *__restrict__ attn, T *__restrict__ out, int S, int M, int Lq,
            int64_t rows)
{
    constexpr int E = Chunk<T>::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    constexpr int PPG = LP / G;
    static_assert(D % E == 0 && 32 % LPT == 0 && LP % G == 0 && LP <= 32 && 32 % LP == 0, "unsupported");

    __shared__ int sH[L], sW[L], sStart[L];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    const int lane = threadIdx.x & 31;
    const int g = lane / LPT, sub = lane % LPT;
    const unsigned urow = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (urow >= (unsigned)rows) return;
    int64_t b, row;
    int m;
    decode_row<HEAD_MAJOR>(urow, M, Lq, b, m, row);
    const int MD = M * D;
    const T *vimg = value + (b * S * M + m) * (int64_t)D + sub * E;

    // resolver phase: lane p owns sample point p (lanes >= LP mirror lane p % LP)
    const int rp = lane % LP;
    const int rl = rp / P;
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + row * LP + rp);
    const Resolved mine = resolve_point(xy.x, xy.y, sH[rl], sW[rl], sStart[rl], attn + row * LP + rp);

    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;

#pragma unroll
    for (int it = 0; it < PPG; ++it) {
        const int pt = it * G + g;
        const int pm = __shfl_sync(0xffffffffu, mine.pm, pt);
        const float a = __shfl_sync(0xffffffffu, mine.a, pt);
        const float lh = __shfl_sync(0xffffffffu, mine.lh, pt);
        const float lw = __shfl_sync(0xffffffffu, mine.lw, pt);
        const int pitch = sW[pt / P] * MD;
        const T *p0 = vimg + (int64_t)(pm >> 4) * MD;
        const T *p1 = p0 + pitch;
        float v0[E], v1[E], v2[E], v3[E];
#pragma unroll
        for (int e = 0; e < E; ++e) v0[e] = v1[e] = v2[e] = v3[e] = 0.f;
        if (pm & 1) Chunk<T>::load(p0, v0);
        if (pm & 2) Chunk<T>::load(p0 + MD, v1);
        if (pm & 4) Chunk<T>::load(p1, v2);
        if (pm & 8) Chunk<T>::load(p1 + MD, v3);
        const float ah = a * (1.f - lh), al = a * lh, hw = 1.f - lw;
        const float w0 = ah * hw, w1 = ah * lw, w2 = al * hw, w3 = al * lw;
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

// Persistent flavour of msda_fwd_v2: a fixed grid of resident warps strides over the rows in memory order
// and keeps the NEXT row's location / weight loads in flight while the current row gathers, so the DRAM
// latency of the loc/attn stream is off the critical path and no CTA-launch overhead is paid per 4-8 rows.
template <typename T, int D, int L, int P, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
msda_fwd_v2p(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
             const float *__restrict__ loc, const float *__restrict__ attn, T *__restrict__ out, int S, int M, int Lq,
             int64_t rows)
{
    constexpr int E = Chunk<T>::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    constexpr int PPG = LP / G;
    static_assert(D % E == 0 && 32 % LPT == 0 && LP % G == 0 && LP <= 32 && 32 % LP == 0, "unsupported");

    __shared__ int sH[L], sW[L], sStart[L];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    const int lane = threadIdx.x & 31;
    const int g = lane / LPT, sub = lane % LPT;
    const unsigned stride = gridDim.x * WARPS;
    const int MD = M * D;
    const int rp = lane % LP;
    const int rl = rp / P;
    const int rH = sH[rl], rW = sW[rl], rStart = sStart[rl];
    int pitch[PPG];
#pragma unroll
    for (int it = 0; it < PPG; ++it) pitch[it] = sW[(it * G + g) / P] * MD;

    unsigned urow = blockIdx.x * WARPS + (threadIdx.x >> 5);
    float2 xy_next = make_float2(0.f, 0.f);
    float a_next = 0.f;
    if (urow < (unsigned)rows) {
        xy_next = __ldg(reinterpret_cast<const float2 *>(loc) + (int64_t)urow * LP + rp);
        a_next = __ldg(attn + (int64_t)urow * LP + rp);
    }
    for (; urow < (unsigned)rows; urow += stride) {
        const int64_t row = urow;
        const float2 xy = xy_next;
        const float a_raw = a_next;
        if (urow + stride < (unsigned)rows) {
            xy_next = __ldg(reinterpret_cast<const float2 *>(loc) + (row + stride) * LP + rp);
            a_next = __ldg(attn + (row + stride) * LP + rp);
        }
        const int m = (int)(urow % (unsigned)M);
        const int64_t b = urow / ((unsigned)M * (unsigned)Lq);
        const T *vimg = value + (b * S * M + m) * (int64_t)D + sub * E;
        const Resolved mine = resolve_point_v(xy.x, xy.y, rH, rW, rStart, a_raw);

        float acc[E];
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = 0.f;
#pragma unroll
        for (int it = 0; it < PPG; ++it) {
            const int pt = it * G + g;
            const int pm = __shfl_sync(0xffffffffu, mine.pm, pt);
            const float a = __shfl_sync(0xffffffffu, mine.a, pt);
            const float lh = __shfl_sync(0xffffffffu, mine.lh, pt);
            const float lw = __shfl_sync(0xffffffffu, mine.lw, pt);
            const T *p0 = vimg + (int64_t)(pm >> 4) * MD;
            const T *p1 = p0 + pitch[it];
            float v0[E], v1[E], v2[E], v3[E];
#pragma unroll
            for (int e = 0; e < E; ++e) v0[e] = v1[e] = v2[e] = v3[e] = 0.f;
            if (pm & 1) Chunk<T>::load(p0, v0);
            if (pm & 2) Chunk<T>::load(p0 + MD, v1);
            if (pm & 4) Chunk<T>::load(p1, v2);
            if (pm & 8) Chunk<T>::load(p1 + MD, v3);
            const float ah = a * (1.f - lh), al = a * lh, hw = 1.f - lw;
            const float w0 = ah * hw, w1 = ah * lw, w2 = al * hw, w3 = al * lw;
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

template <typename T, int D, int L, int P, int WARPS, bool HEAD_MAJOR>
__global__ void __launch_bounds__(WARPS * 32)
msda_bwd_v2(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
            const float *__restrict__ loc, const float *__restrict__ attn, const T *__restrict__ grad_out,
            float *__restrict__ gv_acc, float *__restrict__ grad_loc, float *__restrict__ grad_attn, int S, int M,
            int Lq, int64_t rows)
{
    constexpr int E = Chunk<T>::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    constexpr int PPG = LP / G;
    static_assert(D % E == 0 && 32 % LPT == 0 && LP % G == 0 && LP <= 32 && 32 % LP == 0, "unsupported");
    static_assert(PPG <= LPT && (PPG & (PPG - 1)) == 0, "halving reduction needs PPG to be a power of two <= LPT");

    __shared__ int sH[L], sW[L], sStart[L];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    const int lane = threadIdx.x & 31;
    const int g = lane / LPT, sub = lane % LPT;
    const unsigned urow = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (urow >= (unsigned)rows) return;
    int64_t b, row;
    int m;
    decode_row<HEAD_MAJOR>(urow, M, Lq, b, m, row);
    const int MD = M * D;
    const int64_t img = (b * S * M + m) * (int64_t)D + sub * E;
    const T *vimg = value + img;
    float *gimg = gv_acc + img;

    const int rp = lane % LP;
    const int rl = rp / P;
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + row * LP + rp);
    const Resolved mine = resolve_point(xy.x, xy.y, sH[rl], sW[rl], sStart[rl], attn + row * LP + rp);

    float go[E];
    Chunk<T>::load(grad_out + row * D + sub * E, go);

    float part[3 * PPG];  // (s_attn, s_x, s_y) of iteration it at [3*it .. 3*it+2]

#pragma unroll
    for (int it = 0; it < PPG; ++it) {
        const int pt = it * G + g;
        const int pm = __shfl_sync(0xffffffffu, mine.pm, pt);
        const float a = __shfl_sync(0xffffffffu, mine.a, pt);
        const float lh = __shfl_sync(0xffffffffu, mine.lh, pt);
        const float lw = __shfl_sync(0xffffffffu, mine.lw, pt);
        const int pitch = sW[pt / P] * MD;
        const int64_t o0 = (int64_t)(pm >> 4) * MD, o1 = o0 + pitch;
        float v0[E], v1[E], v2[E], v3[E];
#pragma unroll
        for (int e = 0; e < E; ++e) v0[e] = v1[e] = v2[e] = v3[e] = 0.f;
        if (pm & 1) Chunk<T>::load(vimg + o0, v0);
        if (pm & 2) Chunk<T>::load(vimg + o0 + MD, v1);
        if (pm & 4) Chunk<T>::load(vimg + o1, v2);
        if (pm & 8) Chunk<T>::load(vimg + o1 + MD, v3);
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float ah = a * hh, al = a * lh;
        if (pm & 1) red_add_chunk<E>(gimg + o0, go, ah * hw);
        if (pm & 2) red_add_chunk<E>(gimg + o0 + MD, go, ah * lw);
        if (pm & 4) red_add_chunk<E>(gimg + o1, go, al * hw);
        if (pm & 8) red_add_chunk<E>(gimg + o1 + MD, go, al * lw);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            d0 = fmaf(go[e], v0[e], d0);
            d1 = fmaf(go[e], v1[e], d1);
            d2 = fmaf(go[e], v2[e], d2);
            d3 = fmaf(go[e], v3[e], d3);
        }
        part[3 * it + 0] = hh * (hw * d0 + lw * d1) + lh * (hw * d2 + lw * d3);
        part[3 * it + 1] = a * (hh * (d1 - d0) + lh * (d3 - d2));
        part[3 * it + 2] = a * (hw * (d2 - d0) + lw * (d3 - d1));
    }

    group_reduce3<PPG, LPT>(part, sub);
    constexpr int SPAN = LPT / PPG;  // lanes sharing one iteration's result
    if (sub % SPAN == 0) {
        const int it = sub / SPAN;
        const int pt = it * G + g;
        const int l = pt / P;
        reinterpret_cast<float2 *>(grad_loc)[row * LP + pt] =
            make_float2((float)sW[l] * part[1], (float)sH[l] * part[2]);
        grad_attn[row * LP + pt] = part[0];
    }
}

}  // namespace msda
