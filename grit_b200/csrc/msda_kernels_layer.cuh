// msda_kernels_layer.cuh -- the epilogue of the decoder layer's attention blocks, sm_100a.
//
// DeformableTransformerDecoderLayer.forward (models/detection/det_module.py:330-339) follows the cross-attention's
// output_proj with        tgt = tgt + dropout1(tgt2);  tgt = norm1(tgt)
// (and does the same after self-attention, :331-333, and after the FFN, :316-318): three elementwise / normalisation
// launches in PyTorch (dropout, add, LayerNorm) and five in backward.  At GRIT's decoder shape (150 queries) each of
// those launches costs as much as the whole sampling kernel, so they are one kernel here:
//
//   forward : h = x + keep * scale * z ;  y = (h - mean(h)) * rstd(h) * gamma + beta        one warp per row
//             (saves h, mean, rstd for backward, like torch's LayerNorm saves its input and statistics)
//   backward: xhat = (h - mean) * rstd ;  g = dy * gamma
//             dh = rstd * (g - mean_C(g) - xhat * mean_C(g * xhat)) ;  dx = dh ;  dz = keep * scale * dh
//             dgamma = sum_rows dy * xhat, dbeta = sum_rows dy : per-CTA partial sums, then a fixed-order second
//             stage -- deterministic, no floating-point atomics.
// `keep` is the dropout mask as bytes (what nn.Dropout would draw for the same RNG state: the Python wrapper obtains it
// from torch's own generator); keep == nullptr means dropout is off (eval mode, or p = 0).
// Statistics are computed in fp32 in two passes over registers (mean first, then the centred second moment).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// VPL = float4 chunks per lane; C = VPL * 128.  Lane owns chunks lane, lane + 32, ... (coalesced 512-byte warp loads).
template <int VPL>
__global__ void __launch_bounds__(256)
msda_add_dropout_ln_fwd(const float *__restrict__ x, const float *__restrict__ z, const unsigned char *__restrict__ keep,
                        float scale, const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
                        float *__restrict__ y, float *__restrict__ h_out, float *__restrict__ mean_out,
                        float *__restrict__ rstd_out, int64_t rows)
{
    constexpr int C = VPL * 128;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4 *xr = reinterpret_cast<const float4 *>(x + row * C);
    const float4 *zr = reinterpret_cast<const float4 *>(z + row * C);
    float4 h[VPL];
    float sum = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c4 = lane + 32 * v;
        const float4 a = __ldg(xr + c4), b = __ldg(zr + c4);
        float k0 = scale, k1 = scale, k2 = scale, k3 = scale;
        if (keep) {
            const uchar4 m = __ldg(reinterpret_cast<const uchar4 *>(keep + row * C) + c4);
            k0 = m.x ? scale : 0.f, k1 = m.y ? scale : 0.f, k2 = m.z ? scale : 0.f, k3 = m.w ? scale : 0.f;
        }
        h[v] = make_float4(fmaf(k0, b.x, a.x), fmaf(k1, b.y, a.y), fmaf(k2, b.z, a.z), fmaf(k3, b.w, a.w));
        sum += (h[v].x + h[v].y) + (h[v].z + h[v].w);
    }
    const float mean = warp_sum(sum) * (1.f / C);
    float sq = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const float d0 = h[v].x - mean, d1 = h[v].y - mean, d2 = h[v].z - mean, d3 = h[v].w - mean;
        sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / C) + eps);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c4 = lane + 32 * v;
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(beta) + c4);
        float4 o;
        o.x = fmaf((h[v].x - mean) * rstd, g.x, b.x), o.y = fmaf((h[v].y - mean) * rstd, g.y, b.y);
        o.z = fmaf((h[v].z - mean) * rstd, g.z, b.z), o.w = fmaf((h[v].w - mean) * rstd, g.w, b.w);
        reinterpret_cast<float4 *>(y + row * C)[c4] = o;
        if (h_out) reinterpret_cast<float4 *>(h_out + row * C)[c4] = h[v];
    }
    if (lane == 0 && mean_out) mean_out[row] = mean, rstd_out[row] = rstd;
}

// partial: (gridDim.x, 2, C) -- per-CTA sums of dy*xhat and dy over the CTA's rows, reduced by msda_ln_param_grads.
template <int VPL>
__global__ void __launch_bounds__(256)
msda_add_dropout_ln_bwd(const float *__restrict__ dy, const float *__restrict__ h, const float *__restrict__ mean,
                        const float *__restrict__ rstd, const unsigned char *__restrict__ keep, float scale,
                        const float *__restrict__ gamma, float *__restrict__ dx, float *__restrict__ dz,
                        float *__restrict__ partial, int64_t rows)
{
    constexpr int C = VPL * 128;
    constexpr int WARPS = 8;
    __shared__ float4 red[WARPS][2][VPL][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 gam[VPL], sg[VPL], sb[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        gam[v] = __ldg(reinterpret_cast<const float4 *>(gamma) + lane + 32 * v);
        sg[v] = sb[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t row = (int64_t)blockIdx.x * WARPS + warp; row < rows; row += (int64_t)gridDim.x * WARPS) {
        const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
        float4 xh[VPL], g[VPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c4 = lane + 32 * v;
            const float4 d = __ldg(reinterpret_cast<const float4 *>(dy + row * C) + c4);
            const float4 hv = __ldg(reinterpret_cast<const float4 *>(h + row * C) + c4);
            xh[v] = make_float4((hv.x - mu) * rs, (hv.y - mu) * rs, (hv.z - mu) * rs, (hv.w - mu) * rs);
            g[v] = make_float4(d.x * gam[v].x, d.y * gam[v].y, d.z * gam[v].z, d.w * gam[v].w);
            s1 += (g[v].x + g[v].y) + (g[v].z + g[v].w);
            s2 += (g[v].x * xh[v].x + g[v].y * xh[v].y) + (g[v].z * xh[v].z + g[v].w * xh[v].w);
            sg[v].x = fmaf(d.x, xh[v].x, sg[v].x), sg[v].y = fmaf(d.y, xh[v].y, sg[v].y);
            sg[v].z = fmaf(d.z, xh[v].z, sg[v].z), sg[v].w = fmaf(d.w, xh[v].w, sg[v].w);
            sb[v].x += d.x, sb[v].y += d.y, sb[v].z += d.z, sb[v].w += d.w;
        }
        const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c4 = lane + 32 * v;
            float4 dh;
            dh.x = rs * (g[v].x - m1 - xh[v].x * m2), dh.y = rs * (g[v].y - m1 - xh[v].y * m2);
            dh.z = rs * (g[v].z - m1 - xh[v].z * m2), dh.w = rs * (g[v].w - m1 - xh[v].w * m2);
            reinterpret_cast<float4 *>(dx + row * C)[c4] = dh;
            float k0 = scale, k1 = scale, k2 = scale, k3 = scale;
            if (keep) {
                const uchar4 m = __ldg(reinterpret_cast<const uchar4 *>(keep + row * C) + c4);
                k0 = m.x ? scale : 0.f, k1 = m.y ? scale : 0.f, k2 = m.z ? scale : 0.f, k3 = m.w ? scale : 0.f;
            }
            reinterpret_cast<float4 *>(dz + row * C)[c4] = make_float4(k0 * dh.x, k1 * dh.y, k2 * dh.z, k3 * dh.w);
        }
    }
    // fixed-order reduction over the CTA's warps
#pragma unroll
    for (int v = 0; v < VPL; ++v) red[warp][0][v][lane] = sg[v], red[warp][1][v][lane] = sb[v];
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * VPL * 32; i += blockDim.x) {
        const int which = i / (VPL * 32), v = (i / 32) % VPL, l = i % 32;
        float4 acc = red[0][which][v][l];
        for (int w = 1; w < WARPS; ++w) {
            const float4 t = red[w][which][v][l];
            acc.x += t.x, acc.y += t.y, acc.z += t.z, acc.w += t.w;
        }
        reinterpret_cast<float4 *>(partial + ((int64_t)blockIdx.x * 2 + which) * C)[l + 32 * v] = acc;
    }
}

// dgamma[c] = sum_b partial[b][0][c], dbeta[c] = sum_b partial[b][1][c], in block order (deterministic).
__global__ void msda_ln_param_grads(const float *__restrict__ partial, int blocks, int C, float *__restrict__ dgamma,
                                    float *__restrict__ dbeta)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, b = 0.f;
    for (int i = 0; i < blocks; ++i) {
        a += partial[((int64_t)i * 2) * C + c];
        b += partial[((int64_t)i * 2 + 1) * C + c];
    }
    dgamma[c] = a;
    dbeta[c] = b;
}

// ---- GroupNorm epilogue -> packed memory ----------------------------------------------------------------
// Detector.input_proj (models/detection/detector.py:39-44, 64) is Conv2d(1x1) + GroupNorm(32, C) per level, and
// prepare_od_inputs (det_module.py:146-155) then flattens / transposes / concatenates the normalised NCHW maps into the
// (N, S, C) memory the op reads: GroupNorm writes N*C*H*W, the re-layout reads and writes it again.  Here the
// GroupNorm epilogue writes the memory layout directly (optionally as bf16): msda_gn_stats computes mean / rstd per
// (level, image, group) from the conv output, msda_pack_levels_gn normalises, applies the affine and transposes through
// a shared-memory tile in one pass.
struct GnPackArgs {
    const float *level[8];   // conv outputs, NCHW fp32
    const float *gamma[8];   // GroupNorm weight / bias of each level's input_proj
    const float *beta[8];
    int hw[8];
    int start[8];
    int tile_start[9];
    int num_levels;
};

// stats: (L, N, G, 2) = mean, rstd.  One CTA per (group, image, level); a group's channels are contiguous in NCHW.
__global__ void __launch_bounds__(256)
msda_gn_stats(GnPackArgs args, int N, int C, int G, float eps, float *__restrict__ stats)
{
    const int g = blockIdx.x, n = blockIdx.y, l = blockIdx.z;
    const int cpg = C / G, hw = args.hw[l];
    const int64_t m = (int64_t)cpg * hw;
    const float *x = args.level[l] + ((int64_t)n * C + (int64_t)g * cpg) * hw;
    const float shift = m > 0 ? __ldg(x) : 0.f;  // shifted sums: stable when |mean| >> std
    float s = 0.f, ss = 0.f;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
        const float d = __ldg(x + i) - shift;
        s += d;
        ss = fmaf(d, d, ss);
    }
    __shared__ float red[2][8];
    s = warp_sum(s), ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s, red[1][threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w) a += red[0][w], b += red[1][w];
        const float inv = m > 0 ? 1.f / (float)m : 0.f;
        const float mu = a * inv;
        const float var = fmaxf(b * inv - mu * mu, 0.f);
        float *o = stats + (((int64_t)l * N + n) * G + g) * 2;
        o[0] = shift + mu;
        o[1] = rsqrtf(var + eps);
    }
}

template <typename TOUT>
__device__ __forceinline__ TOUT gn_cast(float v);
template <>
__device__ __forceinline__ float gn_cast<float>(float v)
{
    return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 gn_cast<__nv_bfloat16>(float v)
{
    return __float2bfloat16_rn(v);
}

template <typename TOUT>
__global__ void __launch_bounds__(256)
msda_pack_levels_gn(GnPackArgs args, const float *__restrict__ stats, TOUT *__restrict__ memory, int N, int C, int G, int S)
{
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int c0 = blockIdx.y * 32;
    int l = 0;
    while (l + 1 < args.num_levels && (int)blockIdx.x >= args.tile_start[l + 1]) ++l;
    const int p0 = ((int)blockIdx.x - args.tile_start[l]) * 32;
    const int hw = args.hw[l], cpg = C / G;
    const float *lvl = args.level[l] + (int64_t)n * C * hw;
    TOUT *mem = memory + ((int64_t)n * S + args.start[l]) * C;
    const float *st = stats + ((int64_t)l * N + n) * G * 2;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {  // read (c, p) with p fastest, normalise
        const int c = c0 + ty + j, p = p0 + tx;
        if (c < C && p < hw) {
            const float mu = __ldg(st + (c / cpg) * 2), rs = __ldg(st + (c / cpg) * 2 + 1);
            tile[ty + j][tx] = fmaf((__ldg(lvl + (int64_t)c * hw + p) - mu) * rs, __ldg(args.gamma[l] + c),
                                    __ldg(args.beta[l] + c));
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {  // write (p, c) with c fastest
        const int p = p0 + ty + j, c = c0 + tx;
        if (c < C && p < hw) mem[(int64_t)p * C + c] = gn_cast<TOUT>(tile[tx][ty + j]);
    }
}

}  // namespace msda
