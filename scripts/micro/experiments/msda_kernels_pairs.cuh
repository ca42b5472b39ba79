// msda_kernels_pairs.cuh -- bf16 forward over a PAIR-PACKED copy of value (sm_100a, D = 32).
//
// Why: the gather is bound by the number of 128-byte REQUESTS an SM can send to L2, not by bytes.
// A bf16 tap of a 32-channel head is 64 bytes, so in the reference layout (N, S, M, D) -- neighbouring pixels of one head
// 512 bytes apart -- a bf16 forward issues exactly as many requests as an fp32 one and runs no faster.  The two
// horizontal taps of a bilinear sample, (r, c) and (r, c+1), are neighbours in the flattened pixel index, so in a
// HEAD-MAJOR copy they are 128 contiguous bytes; storing that copy twice, the second shifted by one pixel, makes every
// such pair one ALIGNED 128-byte line (pair starts at an even pixel -> copy 0, odd -> copy 1): two requests per sample
// point instead of four.
//
//   msda_pack_pairs    value (N, S, M, 32) bf16  ->  packed (N, M, 2, Sp, 32) bf16, Sp = S + 2 rounded up to even;
//                      copy 0 holds pixel p at slot p, copy 1 at slot p + 1; unused slots are zero.  One pass, 3 V bytes
//                      of traffic, charged to the forward that uses it.
//   msda_fwd_pairs     warp per (image, query, head) row like the row kernel; 8 lanes per sample point: lane half h = 0/1
//                      takes the left / right tap, 4 lanes x 16 bytes cover its 32 channels; per point two LDG.E.128 per
//                      lane (top pair, bottom pair).  Lane groups and halves are combined with xor-shuffles at the end.
// Semantics are resolve_point_v() + per-tap masks, identical to every other kernel.
#pragma once

#include "msda_common.cuh"

namespace msda {

__host__ __device__ __forceinline__ int64_t pairs_padded_pixels(int64_t S) { return (S + 3) & ~(int64_t)1; }

// grid: (ceil(S / 8), N); block 256 = 8 pixels x 8 heads x 4 chunks of 16 bytes (M <= 8 per pass, loops over more heads)
__global__ void __launch_bounds__(256)
msda_pack_pairs(const __nv_bfloat16 *__restrict__ value, __nv_bfloat16 *__restrict__ packed, int S, int M, int64_t Sp)
{
    const int n = blockIdx.y;
    const int chunk = threadIdx.x & 3;           // 16-byte chunk of the 64-byte (pixel, head) tap
    const int mm = (threadIdx.x >> 2) & 7;       // head within a group of 8
    const int px = threadIdx.x >> 5;             // pixel within the block's 8
    const int p = blockIdx.x * 8 + px;
    if (p >= S) return;
    for (int m = mm; m < M; m += 8) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(value + (((int64_t)n * S + p) * M + m) * 32) + chunk);
        __nv_bfloat16 *plane = packed + (((int64_t)n * M + m) * 2) * Sp * 32;
        reinterpret_cast<uint4 *>(plane + (int64_t)p * 32)[chunk] = v;                    // copy 0, slot p
        reinterpret_cast<uint4 *>(plane + (Sp + p + 1) * 32)[chunk] = v;                  // copy 1, slot p + 1
    }
}

// zero the padding slots: copy 0 slots S..Sp-1, copy 1 slot 0 and slots S+1..Sp-1  (grid: N*M blocks of 32 threads)
__global__ void msda_pack_pairs_pad(__nv_bfloat16 *__restrict__ packed, int S, int64_t Sp)
{
    __nv_bfloat16 *plane = packed + (int64_t)blockIdx.x * 2 * Sp * 32;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const int chunk = threadIdx.x & 3, k = threadIdx.x >> 2;  // k = 0..7 candidate slots
    if (k == 0) reinterpret_cast<uint4 *>(plane + Sp * 32)[chunk] = z;  // copy 1, slot 0
    for (int64_t s = S + k; s < Sp; s += 8) reinterpret_cast<uint4 *>(plane + s * 32)[chunk] = z;
    for (int64_t s = S + 1 + k; s < Sp; s += 8) reinterpret_cast<uint4 *>(plane + (Sp + s) * 32)[chunk] = z;
}

template <int L, int P, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
msda_fwd_pairs(const __nv_bfloat16 *__restrict__ packed, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ lsi, const float *__restrict__ loc, const float *__restrict__ attn,
               __nv_bfloat16 *__restrict__ out, int S, int M, unsigned rows_per_image, int64_t Sp)
{
    constexpr int D = 32;
    constexpr int LP = L * P;
    constexpr int G = 4;           // sample points in flight per warp (8 lanes each)
    constexpr int PPG = LP / G;
    static_assert(LP % G == 0 && LP <= 32, "unsupported");

    __shared__ int sH[L], sW[L], sStart[L];
    stage_levels<L>(shapes, lsi, sH, sW, sStart);

    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, h = (lane >> 2) & 1, sub = lane & 3;
    const unsigned r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= rows_per_image) return;
    const unsigned m = ((M & (M - 1)) == 0) ? (r & (unsigned)(M - 1)) : (r % (unsigned)M);
    const int64_t row = (int64_t)blockIdx.y * rows_per_image + r;
    // this lane's view of the two copies of its (image, head) plane: + h * 32 (left / right tap) + sub * 8 channels
    const __nv_bfloat16 *plane = opaque_ptr(packed + (((int64_t)blockIdx.y * M + m) * 2) * Sp * D + h * D + sub * 8);
    const int copy_stride = (int)(Sp * D);

    const int rp = lane % LP, rl = rp / P;
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + row * LP + rp);
    const Resolved mine = resolve_point(xy.x, xy.y, sH[rl], sW[rl], sStart[rl], attn + row * LP + rp);
    const bool all_valid = __all_sync(0xffffffffu, (mine.pm & 15) == 15);

    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int it = 0; it < PPG; ++it) {
        const int pt = it * G + g;
        const int pm = __shfl_sync(0xffffffffu, mine.pm, pt);
        const float a = __shfl_sync(0xffffffffu, mine.a, pt);
        const float lh = __shfl_sync(0xffffffffu, mine.lh, pt);
        const float lw = __shfl_sync(0xffffffffu, mine.lw, pt);
        const int p_top = pm >> 4, p_bot = p_top + sW[pt / P];   // first pixel of the top / bottom pair (may be -1)
        // pair starting at pixel p lives in copy (p & 1) at slot p + (p & 1): an even slot, i.e. a 128-byte aligned line
        const int o_top = (p_top & 1) * copy_stride + (p_top + (p_top & 1)) * D;
        const int o_bot = (p_bot & 1) * copy_stride + (p_bot + (p_bot & 1)) * D;
        float vt[8], vb[8];
        if (all_valid) {
            Chunk<__nv_bfloat16>::load(plane + o_top, vt);
            Chunk<__nv_bfloat16>::load(plane + o_bot, vb);
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) vt[e] = vb[e] = 0.f;
            if (pm & (h ? 2 : 1)) Chunk<__nv_bfloat16>::load(plane + o_top, vt);
            if (pm & (h ? 8 : 4)) Chunk<__nv_bfloat16>::load(plane + o_bot, vb);
        }
        const float wx = h ? lw : 1.f - lw;
        const float wt = (a - a * lh) * wx, wb = (a * lh) * wx;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wt, vt[e], fmaf(wb, vb[e], acc[e]));
    }
    // combine the two halves (xor 4) and the four lane groups (xor 8, 16)
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
    }
    if (lane < 4) Chunk<__nv_bfloat16>::store(out + row * D + sub * 8, acc);
}

}  // namespace msda
