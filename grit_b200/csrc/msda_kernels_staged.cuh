// msda_kernels_staged.cuh -- shared-memory-staged forward for large query counts (sm_100a).
//
// Why: every bilinear tap of the row kernel (msda_kernels_v5.cuh) is a 128-byte line that misses L1 and comes from
// L2.  The coarse pyramid levels of ONE (image, head) pair are tiny, though -- at 800x1333 levels 2+3 are 1323 pixels =
// 169 KB in fp32 -- and receive half of all taps, and shared memory serves random 128-byte lines 1.8x faster than L2
// (profiles/r02_micro_gather_paths.txt: 284 vs 162 G lines/s).  So:
//
//   * a work item is (image, head, chunk of the queries); one 1024-thread CTA per item stages that head's coarse-level
//     value planes in shared memory (which levels fit is decided on the device from spatial_shapes -- no host sync --
//     greedy from the smallest plane up, within the dynamic shared-memory budget passed at launch) and then walks the
//     chunk's rows exactly like the row kernel (warp per row, resolve once, lane group per tap): taps on staged levels
//     are LDS.128, the others L2 round trips;
//   * items are ordinary CTAs of a grid several times larger than the machine (images outermost, so the CTAs in flight
//     work on a few images whose maps stay L2-resident): the hardware block scheduler hands them out as SMs free up.
//     The first version of this kernel (round 1) was persistent with a static slice of the rows per SM; B200's SMs do
//     not run at one speed (two dies, per-GPU SM -> L2 distance map), so static slices left a tail that made the kernel
//     slower than the row kernel on some boxes (1.53 vs 1.47 ms) and faster on others (1.34 ms); items of ~1024 rows keep
//     the scheduler's granularity fine (5 chunks per (image, head) lost a fifth to wave quantisation, 1.54 ms) while
//     staging costs 169 KB of L2 reads per 8 MB of taps.
//   * it is the default only where at least three of four levels fit on chip (a 384x640-class pyramid: 0.60 vs 0.62 ms
//     for the row kernel); with two levels staged (800x1333) the row kernel, at 1.31 ms after its addressing was cut to
//     one IMAD.WIDE per tap, beats the staged kernel's 1.34 ms (profiles/r02_staged_ab.txt).
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kV3MaxLevels = 8;

struct V3Plan {
    int H[kV3MaxLevels], W[kV3MaxLevels], start[kV3MaxLevels];
    int sbase[kV3MaxLevels];  // element offset of pixel `start[l]`... see plan_levels(); INT_MIN/2 when not staged
    int soff[kV3MaxLevels];   // element offset of the level's plane inside the staging buffer, -1 when not staged
    int staged_elems;         // total staged elements (D per pixel)
};

constexpr int kNotStaged = -(1 << 30);

// Thread 0 decides which levels live in shared memory: smallest planes first while they fit.
template <int L, int D>
__device__ __forceinline__ void plan_levels(const int64_t *shapes, const int64_t *lsi, int budget_elems, V3Plan &plan)
{
    if (threadIdx.x == 0) {
        int order[L];
        for (int l = 0; l < L; ++l) {
            plan.H[l] = (int)shapes[2 * l];
            plan.W[l] = (int)shapes[2 * l + 1];
            plan.start[l] = (int)lsi[l];
            plan.soff[l] = -1;
            plan.sbase[l] = kNotStaged;
            order[l] = l;
        }
        for (int i = 1; i < L; ++i)  // insertion sort by plane size
            for (int j = i; j > 0 && plan.H[order[j]] * plan.W[order[j]] < plan.H[order[j - 1]] * plan.W[order[j - 1]]; --j) {
                const int t = order[j];
                order[j] = order[j - 1];
                order[j - 1] = t;
            }
        int used = 0;
        for (int i = 0; i < L; ++i) {
            const int l = order[i];
            const long long need = (long long)plan.H[l] * plan.W[l] * D;
            if (need > 0 && used + need <= budget_elems) {
                plan.soff[l] = used;
                plan.sbase[l] = used - plan.start[l] * D;  // smem element index of image pixel p is sbase + p*D
                used += (int)need;
            }
        }
        plan.staged_elems = used;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int D, int L, int P, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
msda_fwd_v3(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
            const float *__restrict__ loc, const float *__restrict__ attn, T *__restrict__ out, int N, int S, int M,
            int Lq, int budget_elems, int chunks)
{
    constexpr int E = Chunk<T>::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    constexpr int PPG = LP / G;
    constexpr int WARPS = THREADS / 32;
    static_assert(L <= kV3MaxLevels && LP <= 32 && 32 % LP == 0 && LP % G == 0, "unsupported");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *stage = reinterpret_cast<T *>(smem_raw);
    __shared__ V3Plan plan;
    plan_levels<L, D>(shapes, lsi, budget_elems, plan);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPT, sub = lane % LPT;
    const int MD = M * D;
    const int rp = lane % LP, rl = rp / P;
    const int rH = plan.H[rl], rW = plan.W[rl], rStart = plan.start[rl];
    // item = (image, chunk, head), head fastest: item = (b * chunks + c) * M + m.  The grid normally has one CTA per
    // item (dynamic scheduling by the hardware); a smaller grid walks the items round-robin (A/B knob "staged_persistent").
    const unsigned n_items = (unsigned)N * (unsigned)chunks * (unsigned)M;
    for (unsigned item = blockIdx.x; item < n_items; item += gridDim.x) {
        {
            const int m = (int)(item % (unsigned)M);
            const int c = (int)((item / (unsigned)M) % (unsigned)chunks);
            const int b = (int)(item / ((unsigned)M * (unsigned)chunks));
            const int per = (Lq + chunks - 1) / chunks;
            const int q0 = c * per, q1 = min(q0 + per, Lq);
            if (q0 >= q1) continue;
            const T *vimg = value + ((int64_t)b * S * M + m) * D;
            const T *vlane = opaque_ptr(vimg + sub * E);  // per-lane base of the L2 taps: one IMAD.WIDE per address

            // ---- stage this head's coarse planes ---------------------------------------------------
            __syncthreads();  // previous segment's readers are done
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const int so = plan.soff[l];
                if (so < 0) continue;
                const int chunks = plan.H[l] * plan.W[l] * LPT;
                const T *src = vimg + (int64_t)plan.start[l] * MD;
                for (int i = threadIdx.x; i < chunks; i += THREADS) {
                    const int pix = i / LPT, ch = i % LPT;
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (int64_t)pix * MD + ch * E));
                    *reinterpret_cast<uint4 *>(stage + so + pix * D + ch * E) = v;
                }
            }
            __syncthreads();

            // ---- rows of this segment ---------------------------------------------------------------
            // software pipeline: the next row's location / weight are in flight while this row gathers
            float2 xy_next = make_float2(0.f, 0.f);
            float a_next = 0.f;
            if (q0 + warp < q1) {
                const int64_t r0 = ((int64_t)b * Lq + q0 + warp) * M + m;
                xy_next = __ldg(reinterpret_cast<const float2 *>(loc) + r0 * LP + rp);
                a_next = __ldg(attn + r0 * LP + rp);
            }
            for (int q = q0 + warp; q < q1; q += WARPS) {
                const int64_t row = ((int64_t)b * Lq + q) * M + m;
                const float2 xy = xy_next;
                const float a_raw = a_next;
                if (q + WARPS < q1) {
                    const int64_t rn = row + (int64_t)WARPS * M;
                    xy_next = __ldg(reinterpret_cast<const float2 *>(loc) + rn * LP + rp);
                    a_next = __ldg(attn + rn * LP + rp);
                }
                const Resolved mine = resolve_point_v(xy.x, xy.y, rH, rW, rStart, a_raw);
                float acc[E];
#pragma unroll
                for (int e = 0; e < E; ++e) acc[e] = 0.f;
                // warp-uniform fast path: every tap of every point of the row is inside its map
                const bool all_valid = __all_sync(0xffffffffu, (mine.pm & 15) == 15);
#pragma unroll
                for (int it = 0; it < PPG; ++it) {
                    const int pt = it * G + g;
                    const int l = pt / P;
                    const int pm = __shfl_sync(0xffffffffu, mine.pm, pt);
                    const float a = __shfl_sync(0xffffffffu, mine.a, pt);
                    const float lh = __shfl_sync(0xffffffffu, mine.lh, pt);
                    const float lw = __shfl_sync(0xffffffffu, mine.lw, pt);
                    const int W = plan.W[l];   // (keeping these two in registers across rows was measured slower:
                    const int sb = plan.sbase[l];  //  the kernel sits at its 64-register budget)
                    const int pix = pm >> 4;
                    float v0[E], v1[E], v2[E], v3[E];
                    if (sb != kNotStaged) {  // taps from the staged plane: LDS.128, 32-bit offsets
                        const T *s0 = stage + (sb + pix * D + sub * E);
                        const int so1 = D, so2 = W * D, so3 = W * D + D;
                        if (all_valid) {
                            Chunk<T>::load_shared(s0, v0);
                            Chunk<T>::load_shared(s0 + so1, v1);
                            Chunk<T>::load_shared(s0 + so2, v2);
                            Chunk<T>::load_shared(s0 + so3, v3);
                        } else {
#pragma unroll
                            for (int e = 0; e < E; ++e) v0[e] = v1[e] = v2[e] = v3[e] = 0.f;
                            if (pm & 1) Chunk<T>::load_shared(s0, v0);
                            if (pm & 2) Chunk<T>::load_shared(s0 + so1, v1);
                            if (pm & 4) Chunk<T>::load_shared(s0 + so2, v2);
                            if (pm & 8) Chunk<T>::load_shared(s0 + so3, v3);
                        }
                    } else {  // taps from L2: one IMAD.WIDE per address
                        const int o0 = pix * MD, o1 = o0 + MD, o2 = o0 + W * MD, o3 = o2 + MD;
                        if (all_valid) {
                            Chunk<T>::load(vlane + o0, v0);
                            Chunk<T>::load(vlane + o1, v1);
                            Chunk<T>::load(vlane + o2, v2);
                            Chunk<T>::load(vlane + o3, v3);
                        } else {
#pragma unroll
                            for (int e = 0; e < E; ++e) v0[e] = v1[e] = v2[e] = v3[e] = 0.f;
                            if (pm & 1) Chunk<T>::load(vlane + o0, v0);
                            if (pm & 2) Chunk<T>::load(vlane + o1, v1);
                            if (pm & 4) Chunk<T>::load(vlane + o2, v2);
                            if (pm & 8) Chunk<T>::load(vlane + o3, v3);
                        }
                    }
                    const float ah = a - a * lh, al = a * lh, hw = 1.f - lw;
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
    }
}

}  // namespace msda
