// msda_kernels_planes.cuh -- backward whose COARSE-level grad_value is accumulated in shared memory as int32 fixed
// point with native integer shared-memory atomics (sm_100a).
//
// Why: the row backward (msda_bwd_v5) is bound by the SM's L1 -> XBAR request path: per (query, head) row it carries 64
// load requests (1 cycle each) and 64 128-byte `red.global.add.v4.f32` (4 cycles each) = 320 cycles.  Half of the reds go
// to the two coarsest levels, whose fp32 gradient planes for ONE (image, head) are tiny (800x1333: 1 323 px = 169 KB).
// sm_100a has no native fp32 shared-memory atomic (atomicAdd(float) and the 64-bit integer add are ATOMS.CAST.SPIN
// loops), but the 32-bit INTEGER add is native (ATOMS.ADD) and a B200 SM absorbs 101 G lines/s of them over the whole
// GPU -- twice the 50 G lines/s of the global reds, and on the shared-memory path, not on the XBAR path
// (scripts/micro/smem_accumulate.cu, profiles/r02_micro_smem_accumulate.txt).  So:
//
//   * a work item is (image, head, chunk of the queries), one CTA per item (items are ordinary CTAs of a grid many times
//     the machine, handed out by the hardware block scheduler, as in msda_kernels_staged.cuh).  Default launch shape: one
//     768-thread CTA per SM with all the shared memory, and a chunk count chosen by the host so that an image's items
//     fill whole waves (800x1333: 37 chunks x 8 heads = 2 x 148 items of 601 queries) -- otherwise the tail of every image
//     runs beside the head of the next and two images' value / grad_value maps compete for the L2 (4.2 instead of 2.7 GB
//     of DRAM traffic).  Alternative ("planes_threads" 256): four 256-thread CTAs per SM, each with the smallest levels'
//     planes in a right-sized slice of the shared memory -- faster timed alone, hungrier inside a long step;
//   * the CTA keeps the gradient planes of the levels that fit (chosen on the device, smallest first) in shared memory
//     as int32 fixed point.  The scale 2^k is chosen PER ITEM from a rigorous bound: a row r adds at most
//     s_r * |g_rc| to any one (pixel, channel) of the planes, where s_r is the sum of |attention weight| over the row's
//     points on the staged levels, so no sampling pattern can make a sum exceed W = max_c sum_r s_r |g_rc|; k is the
//     largest integer with W * 2^k < 2^30, which leaves a factor 2 for the roundings.  The bound is computed by a
//     vectorised pre-pass over the item's attention weights and grad_output rows (the fused variant, RD != 0, knows the
//     weights are a softmax and reads grad_output only).  The quantum is ~W * 2^-30: the rounding noise of a contribution is below the last bit of an fp32 add
//     into a sum of typical size;
//   * the rows are then walked exactly like the row kernel (warp per row, resolve once, lane group per tap, value taps
//     from L2, grad_loc / grad_attn by shuffle reductions); taps on staged levels are 16 ATOMS.ADD per lane and
//     iteration (4 taps x 4 channels; lane group g issues its channels rotated by g so that the four groups of a
//     warp instruction hit disjoint banks), taps on the other levels leave as REDG like before; the all-taps-valid fast
//     path is voted per ITERATION; the row's inputs and outputs are streaming (evict-first) accesses;
//   * the planes are flushed with one vector red per non-zero 16-byte chunk: a coarse pixel crosses the XBAR once per
//     item instead of once per tap (1 323 lines per ~32 K taps).
//   * a non-finite bound (NaN / Inf in grad_output or in an attention weight) switches the item to reds for every
//     level, so non-finite gradients propagate exactly as in the row kernel.
//
// Integer addition is associative, so the staged levels are also summed order-independently.
// Semantics are resolve_point() / bwd_row_body(): ms_deform_im2col_cuda.cuh:33-159, 272-296.
#pragma once

#include "msda_common.cuh"
#include "msda_kernels_staged.cuh"
#include "msda_kernels_v5.cuh"
#include "msda_kernels_fused.cuh"

namespace msda {

__device__ __forceinline__ void reds_add_s32(unsigned saddr, int v)
{
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// One tap line into the plane: channel slot e of this lane goes to word (e + g) & 3 of its 16-byte chunk; `rot` holds the
// four per-lane byte offsets, `gr` the grad_output channels in the same rotated order.
template <int OFF>
__device__ __forceinline__ void plane_add(unsigned line, const unsigned (&rot)[4], const float (&gr)[4], float s)
{
#pragma unroll
    for (int e = 0; e < 4; ++e) reds_add_s32(line + rot[e] + OFF, __float2int_rn(s * gr[e]));
}

// One iteration = the G sample points (one per lane group, all on level l) of a row.  ALL: every tap of the G points is
// inside the map (a per-ITERATION warp vote: with uniform locations 94 / 88 / 77 / 60 % of the iterations of levels
// 0..3 of an 800x1333 pyramid qualify, while only 38 % of whole rows do).  The general path is branch-free: loads are
// predicated with zero-fill, reds are predicated, and a dead tap's plane add becomes "+0 to the first line".
template <typename T, typename CH, int D, bool ALL>
__device__ __forceinline__ void planes_iteration(const T *vimg, float *gimg, int MD, int W, int sb, bool to_plane,
                                                 unsigned plane_lane, const unsigned (&rot)[4], float scale, int pm,
                                                 float a, float lh, float lw, const float (&go)[4],
                                                 const float (&gr)[4], float (&part3)[3])
{
    constexpr int E = 4;
    const int pix = pm >> 4;
    const int o0 = pix * MD, o1 = o0 + MD;
    const int o2 = o0 + W * MD, o3 = o2 + MD;
    float v0[E], v1[E], v2[E], v3[E];
    load_taps<T, CH, ALL>(vimg, o0, o1, o2, o3, pm, v0, v1, v2, v3);
    const float hh = 1.f - lh, hw = 1.f - lw;
    const float ah = a * hh, al = a * lh;
    if (to_plane) {  // warp-uniform
        const float sh = ah * scale, sl = al * scale;
        const unsigned line0 = plane_lane + (unsigned)(sb + pix * D) * 4u;
        const unsigned line2 = line0 + (unsigned)(W * D) * 4u;
        if (ALL) {
            plane_add<0>(line0, rot, gr, sh * hw);
            plane_add<D * 4>(line0, rot, gr, sh * lw);
            plane_add<0>(line2, rot, gr, sl * hw);
            plane_add<D * 4>(line2, rot, gr, sl * lw);
        } else {
            plane_add<0>((pm & 1) ? line0 : plane_lane, rot, gr, (pm & 1) ? sh * hw : 0.f);
            plane_add<0>((pm & 2) ? line0 + D * 4 : plane_lane, rot, gr, (pm & 2) ? sh * lw : 0.f);
            plane_add<0>((pm & 4) ? line2 : plane_lane, rot, gr, (pm & 4) ? sl * hw : 0.f);
            plane_add<0>((pm & 8) ? line2 + D * 4 : plane_lane, rot, gr, (pm & 8) ? sl * lw : 0.f);
        }
    } else {
        red_chunk<E, ALL>(gimg + o0, go, ah * hw, pm & 1);
        red_chunk<E, ALL>(gimg + o1, go, ah * lw, pm & 2);
        red_chunk<E, ALL>(gimg + o2, go, al * hw, pm & 4);
        red_chunk<E, ALL>(gimg + o3, go, al * lw, pm & 8);
    }
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        d0 = fmaf(go[e], v0[e], d0);
        d1 = fmaf(go[e], v1[e], d1);
        d2 = fmaf(go[e], v2[e], d2);
        d3 = fmaf(go[e], v3[e], d3);
    }
    const float top = d1 - d0, bot = d3 - d2;
    part3[0] = hh * fmaf(lw, top, d0) + lh * fmaf(lw, bot, d2);
    part3[1] = a * fmaf(lh, bot - top, top);
    part3[2] = a * fmaf(lw, (d3 - d1) - (d2 - d0), d2 - d0);
}

template <typename T, typename CH, int D, int L, int P>
__device__ __forceinline__ void planes_row_body(const Resolved &mine, const T *vimg, float *gimg, int MD,
                                                const V3Plan &plan, bool use_planes, unsigned plane_lane,
                                                const unsigned (&rot)[4], float scale, int g, const float (&go)[4],
                                                const float (&gr)[4],
                                                float (&part)[3 * (L * P / (32 / (D / 4)))])
{
    constexpr int E = 4;
    constexpr int G = 32 / (D / E);
    constexpr int PPG = L * P / G;
#pragma unroll
    for (int it = 0; it < PPG; ++it) {
        const int pt = it * G + g;
        const int l = (it * G) / P;  // compile-time: G divides P, so every lane group of an iteration is on one level
        const int pm = __shfl_sync(0xffffffffu, mine.pm, pt);
        const float a = __shfl_sync(0xffffffffu, mine.a, pt);
        const float lh = __shfl_sync(0xffffffffu, mine.lh, pt);
        const float lw = __shfl_sync(0xffffffffu, mine.lw, pt);
        const int W = plan.W[l];
        const int sb = plan.sbase[l];
        const bool to_plane = use_planes && sb != kNotStaged;
        float p3[3];
        if (__all_sync(0xffffffffu, (pm & 15) == 15))
            planes_iteration<T, CH, D, true>(vimg, gimg, MD, W, sb, to_plane, plane_lane, rot, scale, pm, a, lh, lw, go,
                                             gr, p3);
        else
            planes_iteration<T, CH, D, false>(vimg, gimg, MD, W, sb, to_plane, plane_lane, rot, scale, pm, a, lh, lw, go,
                                              gr, p3);
        part[3 * it + 0] = p3[0], part[3 * it + 1] = p3[1], part[3 * it + 2] = p3[2];
    }
}

// RD = 0: the plain op (loc = sampling locations, attn = attention weights, outputs grad_sampling_loc / grad_attn_weight).
// RD = 2 | 4: the fused module path of msda_kernels_fused.cuh -- loc = raw sampling offsets, attn = attention LOGITS,
// ref / vratio = reference points (boxes) and optional valid ratios, outputs grad_offsets / grad_logits (softmax backward
// = one more warp reduction per row).  Softmax weights sum to 1, so the bound pre-pass needs grad_output only.
template <typename T, typename CH, int D, int L, int P, int THREADS, int RD = 0>
__global__ void __launch_bounds__(THREADS, THREADS <= 256 ? 1024 / THREADS : 1)  // small CTAs: several per SM, 64 registers
msda_bwd_planes(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
                const float *__restrict__ loc, const float *__restrict__ attn, const T *__restrict__ grad_out,
                float *__restrict__ gv_acc, float *__restrict__ grad_loc, float *__restrict__ grad_attn, int N, int S,
                int M, int Lq, int budget_words, int chunks, const float *__restrict__ ref = nullptr,
                const float *__restrict__ vratio = nullptr)
{
    constexpr int E = CH::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    constexpr int PPG = LP / G;
    constexpr int WARPS = THREADS / 32;
    static_assert(E == 4, "a lane owns four channels = one 16-byte chunk of the fp32 gradient line");
    static_assert(L <= kV3MaxLevels && LP <= 32 && 32 % LP == 0 && LP % G == 0 && P % G == 0 && PPG <= LPT,
                  "unsupported");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *plane = reinterpret_cast<int *>(smem_raw);
    __shared__ V3Plan plan;
    __shared__ float sBound[WARPS][D];
    __shared__ float sScale[2];
    __shared__ int sUse;
    plan_levels<L, D>(shapes, lsi, budget_words, plan);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPT, sub = lane % LPT;
    const int MD = M * D;
    const int rp = lane % LP, rl = rp / P;
    const int rH = plan.H[rl], rW = plan.W[rl], rStart = plan.start[rl];
    const float rInvH = 1.f / (float)rH, rInvW = 1.f / (float)rW;  // fused path: offsets are in pixels of the level
    const unsigned plane_lane = (unsigned)__cvta_generic_to_shared(plane) + (unsigned)sub * 16u;
    unsigned rot[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) rot[e] = (unsigned)((e + g) & 3) * 4u;

    const unsigned n_items = (unsigned)N * (unsigned)chunks * (unsigned)M;
    for (unsigned item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int m = (int)(item % (unsigned)M);
        const int c = (int)((item / (unsigned)M) % (unsigned)chunks);
        const int b = (int)(item / ((unsigned)M * (unsigned)chunks));
        const int per = (Lq + chunks - 1) / chunks;
        const int q0 = c * per, q1 = min(q0 + per, Lq);
        if (q0 >= q1) continue;
        const int64_t img = ((int64_t)b * S * M + m) * D + sub * E;
        const T *vimg = opaque_ptr(value + img);  // one IMAD.WIDE per tap address (this kernel is issue-bound)
        float *gimg = opaque_ptr(gv_acc + img);

        __syncthreads();  // the previous item's flush is complete
        // ---- zero the planes, bound the sums -----------------------------------------------------------------------
        for (int i = threadIdx.x; i < plan.staged_elems / 4; i += THREADS)
            reinterpret_cast<int4 *>(plane)[i] = make_int4(0, 0, 0, 0);
        {
            // eight rows per warp and step: lane j looks at row j / 4; its float4 of attention weights is the four points
            // of level j % 4 (LP = 16, P = 4), its share of the grad_output row the channels [j % 4 * D/4, +D/4).  All the
            // loads of a step are independent, and a 1024-row item is four steps per warp (a first version walked one row
            // per warp and step: 32 dependent round trips to L2 / DRAM per item, ~15 % of the kernel).
            constexpr int CPL = D / 4, EV = Chunk<T>::E, NV = CPL / EV;
            static_assert(LP == 16 && P == 4 && CPL % EV == 0, "pre-pass lane map");
            const int pr = lane >> 2, pk = lane & 3;
            const bool k_staged = plan.sbase[pk] != kNotStaged;
            float bound[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) bound[i] = 0.f;
#pragma unroll 4
            for (int qb = q0 + warp * 8; qb < q1; qb += WARPS * 8) {
                const int q = qb + pr;
                const bool live = q < q1;
                const int64_t row = ((int64_t)b * Lq + (live ? q : q0)) * M + m;
                float4 a4 = make_float4(0.0625f, 0.0625f, 0.0625f, 0.0625f);  // fused: a softmax, the LP weights sum to 1
                if (RD == 0) a4 = __ldg(reinterpret_cast<const float4 *>(attn + row * LP) + pk);
                float gq[CPL];
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float t[EV];
                    Chunk<T>::load(grad_out + row * D + pk * CPL + v * EV, t);
#pragma unroll
                    for (int e = 0; e < EV; ++e) gq[v * EV + e] = t[e];
                }
                float sa = (live && (k_staged || RD != 0)) ? fabsf(a4.x) + fabsf(a4.y) + fabsf(a4.z) + fabsf(a4.w) : 0.f;
                sa += __shfl_xor_sync(0xffffffffu, sa, 1);
                sa += __shfl_xor_sync(0xffffffffu, sa, 2);
#pragma unroll
                for (int i = 0; i < CPL; ++i) bound[i] = fmaf(sa, live ? fabsf(gq[i]) : 0.f, bound[i]);
            }
#pragma unroll
            for (int off = 4; off < 32; off <<= 1)
#pragma unroll
                for (int i = 0; i < CPL; ++i) bound[i] += __shfl_xor_sync(0xffffffffu, bound[i], off);
            if (pr == 0) {
#pragma unroll
                for (int i = 0; i < CPL; ++i) sBound[warp][pk * CPL + i] = bound[i];
            }
        }
        __syncthreads();
        if (warp == 0) {
            // max over channels of the per-channel bound, on bit patterns (all values are >= 0 or NaN; Inf and NaN compare
            // above every finite value, so a non-finite input is seen, not dropped)
            unsigned bits = 0u;
            for (int ch = lane; ch < D; ch += 32) {
                float t = 0.f;
                for (int w = 0; w < WARPS; ++w) t += sBound[w][ch];
                bits = max(bits, __float_as_uint(fabsf(t)));
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, off));
            if (lane == 0) {
                const bool finite = bits <= 0x7f7fffffu;
                int ex = 0;
                frexpf(__uint_as_float(bits), &ex);  // bound < 2^ex
                int k = bits == 0u ? 0 : 30 - ex;
                k = max(-98, min(100, k));
                sScale[0] = ldexpf(1.f, k);
                sScale[1] = ldexpf(1.f, -k);
                sUse = finite && plan.staged_elems > 0;
            }
        }
        __syncthreads();
        const float scale = sScale[0];
        const bool use_planes = sUse != 0;

        // ---- rows of the item ------------------------------------------------------------------------------------------
        for (int q = q0 + warp; q < q1; q += WARPS) {
            const int64_t row = ((int64_t)b * Lq + q) * M + m;
            if (q + WARPS < q1 && lane < 3) {  // the next row's inputs travel to L1 while this row gathers (no registers)
                const int64_t rn = row + (int64_t)WARPS * M;
                const void *pf = lane == 0   ? (const void *)(loc + rn * LP * 2)
                                 : lane == 1 ? (const void *)(attn + rn * LP)
                                             : (const void *)(grad_out + rn * D);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
            }
            // The row's inputs are dead after this use, its outputs are never read again here: streaming (evict-first)
            // loads and stores keep them from pushing value / grad_value lines out of L2 -- with 148 items of ~0.5 MB of
            // row data in flight that was 2.3 GB of extra DRAM traffic per launch.
            float2 xy;
            float a_raw;
            const int64_t bq = (int64_t)b * Lq + q;
            if (RD == 0) {
                asm volatile("ld.global.cs.nc.v2.f32 {%0, %1}, [%2];"
                             : "=f"(xy.x), "=f"(xy.y)
                             : "l"(reinterpret_cast<const float2 *>(loc) + row * LP + rp));
                asm volatile("ld.global.cs.nc.f32 %0, [%1];" : "=f"(a_raw) : "l"(attn + row * LP + rp));
            } else {
                fused_point<L, P, RD == 0 ? 2 : RD>(loc, attn, ref, vratio, b, row, bq, rp, rl, rInvH, rInvW, a_raw, xy.x,
                                                    xy.y);
            }
            const Resolved mine = resolve_point_v(xy.x, xy.y, rH, rW, rStart, a_raw);
            float go[E], gr[E];
            CH::load_stream(grad_out + row * D + sub * E, go);
            {  // gr[e] = go[(e + g) & 3]
                const bool r1 = (g & 1) != 0, r2 = (g & 2) != 0;
                const float t0 = r1 ? go[1] : go[0], t1 = r1 ? go[2] : go[1], t2 = r1 ? go[3] : go[2],
                            t3 = r1 ? go[0] : go[3];
                gr[0] = r2 ? t2 : t0, gr[1] = r2 ? t3 : t1, gr[2] = r2 ? t0 : t2, gr[3] = r2 ? t1 : t3;
            }
            float part[3 * PPG];
            planes_row_body<T, CH, D, L, P>(mine, vimg, gimg, MD, plan, use_planes, plane_lane, rot, scale, g, go, gr,
                                            part);
            int it;
            float r3[3];
            const bool holder = reduce_points<PPG, LPT>(part, sub, it, r3);
            const int pt = it * G + g;
            const int l = pt / P;
            if (RD == 0) {
                if (holder) {
                    __stcs(reinterpret_cast<float2 *>(grad_loc) + row * LP + pt,
                           make_float2((float)plan.W[l] * r3[1], (float)plan.H[l] * r3[2]));
                    __stcs(grad_attn + row * LP + pt, r3[0]);
                }
            } else {
                // as msda_bwd_fused: the TRUE softmax weight of the point (also for skipped points), softmax backward
                const float a_pt = __shfl_sync(0xffffffffu, a_raw, pt);
                float dot = holder ? a_pt * r3[0] : 0.f;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
                if (holder) {
                    __stcs(grad_attn + row * LP + pt, a_pt * (r3[0] - dot));
                    float gx = r3[1], gy = r3[2];  // 2-d: d loc / d off = 1 / (W, H) cancels the (W, H) of d(w_im, h_im) / d loc
                    if (RD == 4) {
                        const float4 rf = load_ref<L, RD == 0 ? 2 : RD>(ref, vratio, bq, b, l);
                        gx = (float)plan.W[l] * gx * (rf.z * 0.5f / (float)P);
                        gy = (float)plan.H[l] * gy * (rf.w * 0.5f / (float)P);
                    }
                    __stcs(reinterpret_cast<float2 *>(grad_loc) + row * LP + pt, make_float2(gx, gy));
                }
            }
        }

        // ---- flush the planes: one red per non-zero 16-byte chunk ---------------------------------------------------
        __syncthreads();
        if (use_planes) {
            const float inv = sScale[1];
            const int n_lines = plan.staged_elems / D;
            float *gbase = gv_acc + ((int64_t)b * S * M + m) * D + sub * E;
            for (int i = warp * G + g; i < n_lines; i += WARPS * G) {
                int pix = 0;
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    const int so = plan.soff[l];
                    if (so >= 0 && i * D >= so && i * D < so + plan.H[l] * plan.W[l] * D)
                        pix = plan.start[l] + (i * D - so) / D;
                }
                const int4 w = *reinterpret_cast<const int4 *>(plane + i * D + sub * E);
                if ((w.x | w.y | w.z | w.w) != 0)
                    red_add_f32x4(gbase + (int64_t)pix * MD, (float)w.x * inv, (float)w.y * inv, (float)w.z * inv,
                                  (float)w.w * inv);
            }
        }
    }
}

}  // namespace msda
