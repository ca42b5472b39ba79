// msda_kernels_v5.cuh -- "lean" specialised kernels (sm_100a): same algorithm as msda_kernels_v2.cuh
// (warp per row, resolve once, lane group per tap), re-cut after reading the v2 SASS, which spent
//   ~40 instructions/row on two runtime integer divisions (row -> image, head),
//   ~6 instructions per tap on 64-bit address arithmetic,
//   30 CS2R + predicate logic per row zero-filling registers for taps that are almost never out of range,
//   17 branches per row around the predicated `red`s of the backward.
// Changes:
//   * 2-D grid: blockIdx.y is the image, blockIdx.x walks the image's (query, head) rows -> no division
//     (head = row & (M-1) when M is a power of two);
//   * 32-bit intra-image element offsets, one IMAD.WIDE per tap address;
//   * a warp-uniform fast path when all taps of all the row's points are inside the map (the common
//     case): unconditional loads / reds, no zero-fill, no predicates; the general path keeps the
//     per-tap predicates, with predicated `red` issued from inline PTX so no branch is needed.
#pragma once

#include "msda_common.cuh"
#include "msda_kernels_generic.cuh"
#include "msda_kernels_binned.cuh"

namespace msda {

__device__ __forceinline__ void red_add_f32x4_if(float *p, float a, float b, float c, float d, int pred)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p),
        "f"(a), "f"(b), "f"(c), "f"(d), "r"(pred)
        : "memory");
}

template <int E, bool ALL>
__device__ __forceinline__ void red_chunk(float *p, const float (&g)[E], float s, int pred)
{
#pragma unroll
    for (int i = 0; i < E; i += 4) {
        if (ALL)
            red_add_f32x4(p + i, s * g[i], s * g[i + 1], s * g[i + 2], s * g[i + 3]);
        else
            red_add_f32x4_if(p + i, s * g[i], s * g[i + 1], s * g[i + 2], s * g[i + 3], pred);
    }
}

// ---- grad_value accumulation policies ------------------------------------------------------------------------------
// AccF32  : fp32 vector reductions (REDG.E.ADD.F32x4).  Fast; summation order, hence the last bits, vary run to run.
// AccFix64: deterministic.  Every contribution is scaled by a power of two chosen on the device from max|attn| *
//           max|grad_out| and the worst-case number of addends, rounded to int64 and added with integer atomics;
//           integer addition is associative, so the result is bit-reproducible (and is the correctly rounded exact
//           sum up to 2^-k).  Costs 8-byte scalar reds instead of 16-byte vector ones.
struct AccF32 {
    using elem = float;
    template <typename T, int E>
    __device__ __forceinline__ void prepare(const T *, int, int)
    {
    }
    template <int E, bool ALL>
    __device__ __forceinline__ void add(float *p, const float (&g)[E], float s, int pred) const
    {
        red_chunk<E, ALL>(p, g, s, pred);
    }
};

// The 8-byte integer reds have no vector form, so the lane -> channel map is re-cut for them: lane `sub` owns channels
// sub, sub+LPT, sub+2*LPT, ... (instead of E consecutive ones), which makes the LPT lanes of a tap hit LPT consecutive
// 8-byte words per instruction -- whole 32-byte sectors -- instead of one word in each of LPT different sectors.
struct AccFix64 {
    using elem = unsigned long long;
    float scale;       // 2^k
    float gs[8];       // this lane's grad_out values in the strided channel map
    int sub, sub_e, lpt;
    template <typename T, int E>
    __device__ __forceinline__ void prepare(const T *grad_row, int sub_, int lpt_)
    {
        sub = sub_, lpt = lpt_, sub_e = sub_ * E;
#pragma unroll
        for (int e = 0; e < E; ++e) gs[e] = to_c<float, T>(grad_row[sub_ + lpt_ * e]);
    }
    template <int E, bool ALL>
    __device__ __forceinline__ void add(unsigned long long *p, const float (&)[E], float s, int pred) const
    {
        if (ALL || pred) {
            unsigned long long *line = p - sub_e + sub;  // first word of this lane in the tap's line
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const long long q = __float2ll_rn(s * gs[e] * scale);
                if (q != 0) atomicAdd(line + lpt * e, (unsigned long long)q);
            }
        }
    }
};

template <typename T, typename CH, bool ALL>
__device__ __forceinline__ void load_taps(const T *vimg, int o0, int o1, int o2, int o3, int pm, float (&v0)[CH::E],
                                          float (&v1)[CH::E], float (&v2)[CH::E], float (&v3)[CH::E])
{
    if (ALL) {
        CH::load(vimg + o0, v0);
        CH::load(vimg + o1, v1);
        CH::load(vimg + o2, v2);
        CH::load(vimg + o3, v3);
    } else {
#pragma unroll
        for (int e = 0; e < CH::E; ++e) v0[e] = v1[e] = v2[e] = v3[e] = 0.f;
        if (pm & 1) CH::load(vimg + o0, v0);
        if (pm & 2) CH::load(vimg + o1, v1);
        if (pm & 4) CH::load(vimg + o2, v2);
        if (pm & 8) CH::load(vimg + o3, v3);
    }
}

// (Choosing the fast path per ITERATION by a warp vote inside the general body -- 94 / 88 / 77 / 60 % of the iterations
//  of levels 0..3 qualify with uniform locations, against 38 % of whole rows -- was measured slower: 48 instead of 40
//  registers, 10 instead of 12 CTAs per SM, 1.43 vs 1.31 ms.)
template <typename T, int D, int L, int P, bool ALL, typename CH = Chunk<T>>
__device__ __forceinline__ void fwd_row_body(const Resolved &mine, const T *vimg, int MD, const int (&sW)[L], int g,
                                             float (&acc)[CH::E])
{
    constexpr int E = CH::E;
    constexpr int G = 32 / (D / E);
    constexpr int PPG = L * P / G;
#pragma unroll
    for (int it = 0; it < PPG; ++it) {
        const int pt = it * G + g;
        const int pm = __shfl_sync(0xffffffffu, mine.pm, pt);
        const float a = __shfl_sync(0xffffffffu, mine.a, pt);
        const float lh = __shfl_sync(0xffffffffu, mine.lh, pt);
        const float lw = __shfl_sync(0xffffffffu, mine.lw, pt);
        const int o0 = (pm >> 4) * MD, o1 = o0 + MD;
        const int o2 = o0 + sW[pt / P] * MD, o3 = o2 + MD;
        const float ah = a - a * lh, al = a * lh, hw = 1.f - lw;
        const float w0 = ah * hw, w1 = ah * lw, w2 = al * hw, w3 = al * lw;
        if (ALL) {
            float v0[E], v1[E], v2[E], v3[E];
            load_taps<T, CH, true>(vimg, o0, o1, o2, o3, pm, v0, v1, v2, v3);
#pragma unroll
            for (int e = 0; e < E; ++e)
                acc[e] = fmaf(w0, v0[e], fmaf(w1, v1[e], fmaf(w2, v2[e], fmaf(w3, v3[e], acc[e]))));
        } else {
            // General path, fp32: one predicated block per tap -- the load directly followed by its FMAs under the tap's
            // predicate -- so nothing is zero-filled (that was 8 CS2R per iteration, and 62 % of the rows of an 800x1333
            // pyramid take this path with uniform locations because some point of the row touches a border).  Same-box
            // A/B at 800x1333 (profiles/r02_fwd_general_path_ab.txt): zero-fill 1.31 ms, this form 1.24 ms; the four
            // predicated loads first and the FMAs after them 1.34 ms, and the same blocks behind a helper function
            // 1.30 ms (ptxas then spills one register at the 40-register budget).  bf16 keeps the zero-fill form: its
            // predicated blocks measured 3.21 vs 2.68 ms, and predicated raw loads
            // (inline PTX) followed by predicated unpack + FMA blocks 2.81 ms.
            if constexpr (E > 4) {
                float v0[E], v1[E], v2[E], v3[E];
                load_taps<T, CH, false>(vimg, o0, o1, o2, o3, pm, v0, v1, v2, v3);
#pragma unroll
                for (int e = 0; e < E; ++e)
                    acc[e] = fmaf(w0, v0[e], fmaf(w1, v1[e], fmaf(w2, v2[e], fmaf(w3, v3[e], acc[e]))));
                continue;
            }
            {
                if (pm & 8) { float v[E]; CH::load(vimg + o3, v);
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[e] = fmaf(w3, v[e], acc[e]); }
                if (pm & 4) { float v[E]; CH::load(vimg + o2, v);
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[e] = fmaf(w2, v[e], acc[e]); }
                if (pm & 2) { float v[E]; CH::load(vimg + o1, v);
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[e] = fmaf(w1, v[e], acc[e]); }
                if (pm & 1) { float v[E]; CH::load(vimg + o0, v);
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[e] = fmaf(w0, v[e], acc[e]); }
            }
        }
    }
}

// Two-phase variant: phase 1 issues the tap loads of ALL the row's points (16 x LDG.E.128 per lane at D=32, L=P=4)
// before phase 2 consumes any of them, trading registers (the taps of a whole row are live at once) for
// memory-level parallelism: the v2 forward is latency-bound (row latency ~5400 cycles at ~36 warps/SM).
template <typename T, int D, int L, int P, bool ALL>
__device__ __forceinline__ void fwd_row_body_hoisted(const Resolved &mine, const T *vimg, int MD, const int (&sW)[L],
                                                     int g, float (&acc)[Chunk<T>::E])
{
    constexpr int E = Chunk<T>::E;
    constexpr int G = 32 / (D / E);
    constexpr int PPG = L * P / G;
    float v[PPG][4][E];
    float w[PPG][4];
#pragma unroll
    for (int it = 0; it < PPG; ++it) {
        const int pt = it * G + g;
        const int pm = __shfl_sync(0xffffffffu, mine.pm, pt);
        const float a = __shfl_sync(0xffffffffu, mine.a, pt);
        const float lh = __shfl_sync(0xffffffffu, mine.lh, pt);
        const float lw = __shfl_sync(0xffffffffu, mine.lw, pt);
        const int o0 = (pm >> 4) * MD, o1 = o0 + MD;
        const int o2 = o0 + sW[pt / P] * MD, o3 = o2 + MD;
        static_assert(ALL, "the hoisted body is the all-taps-valid fast path");
        Chunk<T>::load_ordered(vimg + o0, v[it][0]);
        Chunk<T>::load_ordered(vimg + o1, v[it][1]);
        Chunk<T>::load_ordered(vimg + o2, v[it][2]);
        Chunk<T>::load_ordered(vimg + o3, v[it][3]);
        const float ah = a - a * lh, al = a * lh, hw = 1.f - lw;
        w[it][0] = ah * hw, w[it][1] = ah * lw, w[it][2] = al * hw, w[it][3] = al * lw;
    }
    // scheduling fence: every loaded register passes through an (empty) volatile asm, so no consumer can be
    // scheduled above it and all loads are issued before the first FMA waits on the scoreboard
#pragma unroll
    for (int it = 0; it < PPG; ++it)
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int e = 0; e < E; e += 4)
                asm volatile("" : "+f"(v[it][t][e]), "+f"(v[it][t][e + 1]), "+f"(v[it][t][e + 2]), "+f"(v[it][t][e + 3]));
#pragma unroll
    for (int it = 0; it < PPG; ++it) {
#pragma unroll
        for (int e = 0; e < E; ++e)
            acc[e] = fmaf(w[it][0], v[it][0][e],
                          fmaf(w[it][1], v[it][1][e], fmaf(w[it][2], v[it][2][e], fmaf(w[it][3], v[it][3][e], acc[e]))));
    }
}

// (40 registers, 12 CTAs per SM.  A 32-register build -- __launch_bounds__(128, 16), 64 resident warps -- spills and
//  measured 1.38 vs 1.31 ms; naming a minimum of ONE CTA per SM makes ptxas spend 92 registers and costs 35 %.)
// CH: the lane chunk.  Chunk<T> = 16 bytes per lane (default); ChunkBf16x4 = 8 bytes (4 bf16 channels) per lane, which at
// D=32 gives bf16 the fp32 kernel's shape -- 8 lanes per tap, 4 taps per load instruction, 4-channel register blocks, the
// 40-register budget and the predicated-block general path -- instead of 4 lanes per tap with 8-channel blocks.
template <typename T, int D, int L, int P, int WARPS, bool HOIST = false, typename CH = Chunk<T>>
__global__ void __launch_bounds__(WARPS * 32, (HOIST ? 1024 : (sizeof(T) == 2 && CH::E == 8) ? 1280 : 1536) / (WARPS * 32))  // fp32 <= 40 registers (48 resident warps), bf16 <= 48 (40 warps)
msda_fwd_v5(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
            const float *__restrict__ loc, const float *__restrict__ attn, T *__restrict__ out, int S, int M,
            unsigned rows_per_image)
{
    constexpr int E = CH::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    static_assert(D % E == 0 && 32 % LPT == 0 && LP % G == 0 && LP <= 32, "unsupported");

    __shared__ int sH[L], sW[L], sStart[L];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    const int lane = threadIdx.x & 31;
    const int g = lane / LPT, sub = lane % LPT;
    const unsigned r = blockIdx.x * WARPS + (threadIdx.x >> 5);  // (q*M + m) inside image blockIdx.y
    if (r >= rows_per_image) return;
    const unsigned m = ((M & (M - 1)) == 0) ? (r & (unsigned)(M - 1)) : (r % (unsigned)M);
    const int MD = M * D;
    const int64_t row = (int64_t)blockIdx.y * rows_per_image + r;
    const T *vimg = opaque_ptr(value + ((int64_t)blockIdx.y * S * M + m) * D + sub * E);

    const int rp = lane % LP;
    const int rl = rp / P;
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + row * LP + rp);
    const Resolved mine = resolve_point(xy.x, xy.y, sH[rl], sW[rl], sStart[rl], attn + row * LP + rp);

    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    if (__all_sync(0xffffffffu, (mine.pm & 15) == 15)) {
        if constexpr (HOIST)
            fwd_row_body_hoisted<T, D, L, P, true>(mine, vimg, MD, sW, g, acc);  // (default chunk only)
        else
            fwd_row_body<T, D, L, P, true, CH>(mine, vimg, MD, sW, g, acc);
    } else {
        fwd_row_body<T, D, L, P, false, CH>(mine, vimg, MD, sW, g, acc);
    }

#pragma unroll
    for (int off = LPT; off < 32; off <<= 1) {
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
    }
    if (g == 0) CH::store(out + row * D + sub * E, acc);
}

template <typename T, typename CH, typename ACC, int D, int L, int P, bool ALL, bool SKIP = false>
__device__ __forceinline__ void bwd_row_body(const ACC &accp, const Resolved &mine, const T *vimg,
                                             typename ACC::elem *gimg, int MD,
                                             const int (&sW)[L], int g, const float (&go)[CH::E],
                                             float (&part)[3 * (L * P / (32 / (D / CH::E)))], unsigned skipmask)
{
    constexpr int E = CH::E;
    constexpr int G = 32 / (D / E);
    constexpr int PPG = L * P / G;
#pragma unroll
    for (int it = 0; it < PPG; ++it) {
        const int pt = it * G + g;
        const int pm = __shfl_sync(0xffffffffu, mine.pm, pt);
        const float a = __shfl_sync(0xffffffffu, mine.a, pt);
        const float lh = __shfl_sync(0xffffffffu, mine.lh, pt);
        const float lw = __shfl_sync(0xffffffffu, mine.lw, pt);
        const int o0 = (pm >> 4) * MD, o1 = o0 + MD;
        const int o2 = o0 + sW[pt / P] * MD, o3 = o2 + MD;
        float v0[E], v1[E], v2[E], v3[E];
        load_taps<T, CH, ALL>(vimg, o0, o1, o2, o3, pm, v0, v1, v2, v3);
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float ah = a * hh, al = a * lh;
        // levels in skipmask get their grad_value from msda_bwd_binned / msda_bwd_owned.  In every specialisation that
        // has a SKIP instantiation the level of an iteration is the same for all lane groups (G divides P), so this is a
        // warp-uniform branch around the reds and their multiplies, not a divergent one.
        if (!SKIP || !((skipmask >> (pt / P)) & 1u)) {
            accp.template add<E, ALL>(gimg + o0, go, ah * hw, pm & 1);
            accp.template add<E, ALL>(gimg + o1, go, ah * lw, pm & 2);
            accp.template add<E, ALL>(gimg + o2, go, al * hw, pm & 4);
            accp.template add<E, ALL>(gimg + o3, go, al * lw, pm & 8);
        }
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            d0 = fmaf(go[e], v0[e], d0);
            d1 = fmaf(go[e], v1[e], d1);
            d2 = fmaf(go[e], v2[e], d2);
            d3 = fmaf(go[e], v3[e], d3);
        }
        const float top = d1 - d0, bot = d3 - d2;  // d/dx of the two tap rows
        part[3 * it + 0] = hh * fmaf(lw, top, d0) + lh * fmaf(lw, bot, d2);
        part[3 * it + 1] = a * fmaf(lh, bot - top, top);
        part[3 * it + 2] = a * fmaf(lw, (d3 - d1) - (d2 - d0), d2 - d0);
    }
}

// Group-wide sums of the 3*PPG per-point scalars.  Power-of-two PPG: halving reduction (msda_kernels_v2.cuh), the
// holder of iteration `it` is lane sub = it * (LPT/PPG).  Other PPG (L*P = 12, 20, ... as in 3- or 5-level pyramids):
// plain butterflies, the holder of iteration `it` is lane sub = it.  Returns whether this lane is a holder; `it` and
// r[0..2] are then its iteration and that iteration's (d/d attn, a*d/dw, a*d/dh).
template <int PPG, int LPT>
__device__ __forceinline__ bool reduce_points(float (&part)[3 * PPG], int sub, int &it, float (&r)[3])
{
    if constexpr ((PPG & (PPG - 1)) == 0) {
        group_reduce3<PPG, LPT>(part, sub);
        constexpr int SPAN = LPT / PPG;
        it = sub / SPAN;
        r[0] = part[0], r[1] = part[1], r[2] = part[2];
        return (sub % SPAN) == 0;
    } else {
#pragma unroll
        for (int width = LPT / 2; width >= 1; width >>= 1)
#pragma unroll
            for (int i = 0; i < 3 * PPG; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], width);
        it = sub < PPG ? sub : 0;
        r[0] = r[1] = r[2] = 0.f;
#pragma unroll
        for (int k = 0; k < PPG; ++k)
            if (k == it) r[0] = part[3 * k], r[1] = part[3 * k + 1], r[2] = part[3 * k + 2];
        return sub < PPG;
    }
}

template <typename T, typename CH, typename ACC, int D, int L, int P, int WARPS, bool SKIP = false>
__global__ void __launch_bounds__(WARPS * 32, 1024 / (WARPS * 32))  // <= 64 registers: 32 resident warps per SM
msda_bwd_v5(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
            const float *__restrict__ loc, const float *__restrict__ attn, const T *__restrict__ grad_out,
            typename ACC::elem *__restrict__ gv_acc, const float *__restrict__ det_scale,
            float *__restrict__ grad_loc, float *__restrict__ grad_attn, int S, int M, unsigned rows_per_image,
            int binned_budget)
{
    constexpr int E = CH::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    constexpr int PPG = LP / G;
    static_assert(D % E == 0 && 32 % LPT == 0 && LP % G == 0 && LP <= 32, "unsupported");
    static_assert(PPG <= LPT, "one holder lane per point of the group is needed");

    __shared__ int sH[L], sW[L], sStart[L];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    const int lane = threadIdx.x & 31;
    const int g = lane / LPT, sub = lane % LPT;
    const unsigned r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= rows_per_image) return;
    const unsigned m = ((M & (M - 1)) == 0) ? (r & (unsigned)(M - 1)) : (r % (unsigned)M);
    const int MD = M * D;
    const int64_t row = (int64_t)blockIdx.y * rows_per_image + r;
    const int64_t img = ((int64_t)blockIdx.y * S * M + m) * D + sub * E;
    // (no opaque_ptr here: the backward is bound by the L1->XBAR request path, not by instruction issue, and measured
    //  7 % SLOWER with the leaner addressing -- 3.78 vs 3.53 ms -- because loads and reds then reach that path in bursts)
    const T *vimg = value + img;
    typename ACC::elem *gimg = gv_acc + img;
    ACC accp;
    if constexpr (sizeof(typename ACC::elem) == 8) accp.scale = __ldg(det_scale);
    accp.template prepare<T, E>(grad_out + row * D, sub, LPT);

    const int rp = lane % LP;
    const int rl = rp / P;
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + row * LP + rp);
    const Resolved mine = resolve_point(xy.x, xy.y, sH[rl], sW[rl], sStart[rl], attn + row * LP + rp);

    float go[E];
    CH::load(grad_out + row * D + sub * E, go);

    // levels whose grad_value another kernel produces (same device-side rule as that kernel; 0 = none)
    const unsigned skipmask = SKIP ? binned_level_mask<L>(sH, sW, D, binned_budget) : 0u;

    float part[3 * PPG];
    if (__all_sync(0xffffffffu, (mine.pm & 15) == 15))
        bwd_row_body<T, CH, ACC, D, L, P, true, SKIP>(accp, mine, vimg, gimg, MD, sW, g, go, part, skipmask);
    else
        bwd_row_body<T, CH, ACC, D, L, P, false, SKIP>(accp, mine, vimg, gimg, MD, sW, g, go, part, skipmask);

    int it;
    float r3[3];
    if (reduce_points<PPG, LPT>(part, sub, it, r3)) {
        const int pt = it * G + g;
        const int l = pt / P;
        reinterpret_cast<float2 *>(grad_loc)[row * LP + pt] = make_float2((float)sW[l] * r3[1], (float)sH[l] * r3[2]);
        grad_attn[row * LP + pt] = r3[0];
    }
}

// ---- deterministic mode helpers ---------------------------------------------------------------------------------------
// workspace tail: [0] = max|attn| bits, [1] = max|grad_out| bits (uint, monotone for non-negative floats),
//                 [2] = scale 2^k (float), [3] = 2^-k (float)
// The maxima are taken on the BIT PATTERNS of |x| (monotone for non-negative floats, and Inf / NaN compare above every
// finite value), so a non-finite input is not silently dropped the way fmaxf would drop a NaN.
template <typename T>
__global__ void msda_det_absmax(const float *__restrict__ attn, int64_t n_attn, const T *__restrict__ gout,
                                int64_t n_gout, unsigned *__restrict__ tail)
{
    unsigned ma = 0u, mg = 0u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_attn; i += stride)
        ma = max(ma, __float_as_uint(fabsf(attn[i])));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_gout; i += stride)
        mg = max(mg, __float_as_uint(fabsf(to_c<float, T>(gout[i]))));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ma = max(ma, __shfl_xor_sync(0xffffffffu, ma, off));
        mg = max(mg, __shfl_xor_sync(0xffffffffu, mg, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(tail + 0, ma);
        atomicMax(tail + 1, mg);
    }
}

// One thread: k = 61 - ceil(log2(worst-case addends)) - exponent(max|attn| * max|grad_out|), so that the int64 sums
// cannot overflow whatever the sampling pattern is.  Non-finite inputs (an overflowed loss scale, a NaN upstream) cannot
// be represented in fixed point: the inverse scale becomes NaN, so the fold writes NaN into all of grad_value -- loud,
// like the float path, instead of a clean-looking gradient.
__global__ void msda_det_scale(unsigned *tail, double worst_addends, float attn_bound)
{
    // attn_bound > 0: the caller knows max|attn| a priori (softmax output <= 1 in the fused path)
    const bool finite = (attn_bound > 0.f || tail[0] <= 0x7f7fffffu) && tail[1] <= 0x7f7fffffu;
    const float ma = attn_bound > 0.f ? attn_bound : __uint_as_float(tail[0]);
    const float mg = __uint_as_float(tail[1]);
    int e_prod = 0, e_n = 0;
    frexp((double)ma * (double)mg, &e_prod);  // product < 2^e_prod
    frexp(worst_addends, &e_n);                // addends < 2^e_n
    int k = 61 - e_n - e_prod;
    if (!finite || !(ma > 0.f) || !(mg > 0.f)) k = 0;  // all-zero inputs: any scale works
    k = max(-100, min(100, k));
    reinterpret_cast<float *>(tail)[2] = ldexpf(1.f, k);
    reinterpret_cast<float *>(tail)[3] = finite ? ldexpf(1.f, -k) : __uint_as_float(0x7fc00000u);
}

template <typename T>
__global__ void msda_det_fold(const long long *__restrict__ acc, const float *__restrict__ tail, T *__restrict__ gv,
                              int64_t n, int accumulate)
{
    const double inv = (double)tail[3];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float r = (float)((double)acc[i] * inv);
        if (accumulate) r += to_c<float, T>(gv[i]);
        gv[i] = from_c<T, float>(r);
    }
}

// ---- on-chip ceilings (microbenchmarks used by bench.py; see profiles/r01_micro_red_gather_ceilings.txt) ---------------
// The op moves ~18x more bytes between L2 and the SMs than its compulsory HBM traffic, so the resources that bound it
// are the L2->L1 gather rate (forward) and the L2 reduction rate (backward).  These two kernels measure those rates
// with the kernels' own access pattern and none of their arithmetic: uniform random 128-byte lines inside an
// L2-resident region, four lines per warp instruction.
__device__ __forceinline__ unsigned probe_hash(unsigned x)
{
    x ^= x >> 16, x *= 0x7feb352dU, x ^= x >> 15, x *= 0x846ca68bU, x ^= x >> 16;
    return x;
}

// Index generation must not be what the probe measures: one LCG step + one multiply-high per line (3 instructions),
// instead of a hash and a runtime modulo (~40 instructions, which made the round-1 probes issue-bound at ~56 B/clk/SM).
__device__ __forceinline__ unsigned probe_next(unsigned &state, unsigned n_lines)
{
    state = state * 1664525u + 1013904223u;
    return __umulhi(state, n_lines);
}

__global__ void msda_probe_red(float *dst, unsigned n_lines, int iters)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    unsigned state = probe_hash(w * 4u + g);
    for (int it = 0; it < iters; ++it) {
        const unsigned line = probe_next(state, n_lines);
        red_add_f32x4(dst + (size_t)line * 32 + sub * 4, 1.f, 2.f, 3.f, 4.f);
    }
}

__global__ void msda_probe_gather(const float *__restrict__ src, float *out, unsigned n_lines, int iters)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    unsigned state = probe_hash(w * 4u + g);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; it += 16) {
        float4 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const unsigned line = probe_next(state, n_lines);
            v[k] = __ldg(reinterpret_cast<const float4 *>(src + (size_t)line * 32 + sub * 4));
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) acc.x += v[k].x, acc.y += v[k].y, acc.z += v[k].z, acc.w += v[k].w;
    }
    if (acc.x == 123.456f) out[w] = acc.x + acc.y + acc.z + acc.w;  // keeps the loads alive, never true in practice
}

}  // namespace msda
