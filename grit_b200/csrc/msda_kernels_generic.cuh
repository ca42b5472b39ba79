// msda_kernels_generic.cuh -- generic kernels: any D, L, P; T in {float, double, bf16}; C = compute/loc type (float or
// double).  Warp per row, lanes stride over channels, scalar taps.  This is the path gradcheck (fp64) and odd channel
// counts (30, 71, ...) take.  Semantics follow ms_deform_im2col_cuda.cuh:33-159 and :272-296; see
// oracle/msda_oracle_impl.h for the restatement the tests compare against.
#pragma once

#include "msda_common.cuh"

namespace msda {

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
