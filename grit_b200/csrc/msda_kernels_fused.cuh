// msda_kernels_fused.cuh -- the pre-op arithmetic of MSDeformAttn.forward fused into the gather.
//
// The reference module (models/ops/modules/ms_deform_attn.py:96-111) runs, between its Linears and the core op,
// a chain of elementwise PyTorch kernels: masked_fill on value, softmax over the L*P logits, offsets / normaliser,
// + reference points -- each a full read+write pass over tensors as large as `value` itself, and their autograd
// counterparts in backward.  Here the kernels take the Linears' raw outputs:
//     offsets (N, Lq, M, L, P, 2), logits (N, Lq, M, L*P), reference_points (N, Lq, L, 2|4)
// and the resolver lane of each sample point does softmax (two shuffle reductions over the row's L*P lanes) and
//     loc = ref.xy + off / (W_l, H_l)                    2-d reference points   (reference :105-108)
//     loc = ref.xy + off / P * ref.wh * 0.5              4-d reference boxes    (reference :109-111)
// (with `vratio` != nullptr the reference points arrive un-expanded, (N, Lq, 2|4), and are scaled by the level's valid
// ratio here -- DeformableTransformerDecoderLayer.forward, models/detection/det_module.py:323-328 -- instead of being
// materialised as (N, Lq, L, 2|4) by the caller)
// in registers; sampling_locations / attention_weights are never materialised.  The backward emits grad_offsets and
// grad_logits directly (softmax backward = one extra warp reduction per row); grad of the reference points, when
// needed, is a cheap reduction of grad_offsets done by the caller.  The padding mask is applied by msda_mask_rows
// (zero the masked pixels' rows in place: traffic proportional to the padded fraction, not to the tensor).
#pragma once

#include "msda_kernels_v5.cuh"

namespace msda {

// Maximum over the row's L*P logits.  Every lane of the warp works on the same row (lanes >= LP mirror lanes < LP), so
// this is a full-warp maximum: one sm_100a `redux.sync.max.f32` (CREDUX.MAX.F32) instead of log2(LP) shuffle + max steps.
template <int LP>
__device__ __forceinline__ float segment_max(float v)
{
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}
template <int LP>
__device__ __forceinline__ float segment_sum(float v)
{
#pragma unroll
    for (int off = LP / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Reference point (box) of query bq at level rl: either stored per level, or stored once and scaled by the valid ratio.
template <int L, int RD>
__device__ __forceinline__ float4 load_ref(const float *__restrict__ ref, const float *__restrict__ vratio, int64_t bq,
                                           int n, int rl)
{
    float4 rf = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vratio == nullptr) {
        if (RD == 2) {
            const float2 r = __ldg(reinterpret_cast<const float2 *>(ref) + bq * L + rl);
            rf.x = r.x, rf.y = r.y;
        } else {
            rf = __ldg(reinterpret_cast<const float4 *>(ref) + bq * L + rl);
        }
    } else {
        const float2 vr = __ldg(reinterpret_cast<const float2 *>(vratio) + (int64_t)n * L + rl);
        if (RD == 2) {
            const float2 r = __ldg(reinterpret_cast<const float2 *>(ref) + bq);
            rf.x = r.x * vr.x, rf.y = r.y * vr.y;
        } else {
            const float4 r = __ldg(reinterpret_cast<const float4 *>(ref) + bq);
            rf = make_float4(r.x * vr.x, r.y * vr.y, r.z * vr.x, r.w * vr.y);
        }
    }
    return rf;
}

// 2^x for x <= 0 (softmax numerators): one MUFU.EX2 instead of exp2f's range-scaling sequence; results below the
// smallest normal flush to zero, which a softmax weight of < 1e-38 may.
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Softmax weight and sampling location of point `rp` of row `row`, from the Linears' raw outputs.  inv_w / inv_h are the
// reciprocals of the level's size, computed once per CTA (stage_levels_inv): with three IEEE divisions per lane and row
// (two normalisers and the softmax denominator, ~15 instructions each) the fused kernels executed a fifth more
// instructions than the plain ones and were that much slower (profiles/r02_fused_ab.json: 1.52 vs 1.27 ms forward).
template <int L, int P, int RD>
__device__ __forceinline__ void fused_point(const float *__restrict__ offs, const float *__restrict__ logits,
                                            const float *__restrict__ ref, const float *__restrict__ vratio, int n,
                                            int64_t row, int64_t bq, int rp, int rl,
                                            float inv_h, float inv_w, float &a_soft, float &x, float &y)
{
    constexpr int LP = L * P;
    const float lg = __ldg(logits + row * LP + rp);
    const float2 of = __ldg(reinterpret_cast<const float2 *>(offs) + row * LP + rp);
    const float4 rf = load_ref<L, RD>(ref, vratio, bq, n, rl);
    const float mx = segment_max<LP>(lg);
    const float ex = ex2_approx((lg - mx) * 1.4426950408889634f);
    a_soft = __fdividef(ex, segment_sum<LP>(ex));  // denominator in [1, LP]: MUFU.RCP + FMUL, 2 ulp
    if (RD == 2) {
        x = fmaf(of.x, inv_w, rf.x);
        y = fmaf(of.y, inv_h, rf.y);
    } else {
        x = fmaf(of.x * (0.5f / (float)P), rf.z, rf.x);
        y = fmaf(of.y * (0.5f / (float)P), rf.w, rf.y);
    }
}

template <int L>
__device__ __forceinline__ void stage_levels_inv(const int64_t *shapes, const int64_t *lsi, int (&sH)[L], int (&sW)[L],
                                                 int (&sStart)[L], float (&sInvH)[L], float (&sInvW)[L])
{
    if (threadIdx.x < L) {
        const int h = (int)shapes[2 * threadIdx.x], w = (int)shapes[2 * threadIdx.x + 1];
        sH[threadIdx.x] = h, sW[threadIdx.x] = w;
        sInvH[threadIdx.x] = 1.f / (float)h, sInvW[threadIdx.x] = 1.f / (float)w;
        sStart[threadIdx.x] = (int)lsi[threadIdx.x];
    }
    __syncthreads();
}

template <typename T, int D, int L, int P, int WARPS, int RD, typename CH = Chunk<T>>
__global__ void __launch_bounds__(WARPS * 32, ((sizeof(T) == 2 && CH::E == 8) ? 1280 : 1536) / (WARPS * 32))  // as msda_fwd_v5
msda_fwd_fused(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
               const float *__restrict__ offs, const float *__restrict__ logits, const float *__restrict__ ref,
               const float *__restrict__ vratio, T *__restrict__ out, int S, int M, int Lq, unsigned rows_per_image)
{
    constexpr int E = CH::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    static_assert(D % E == 0 && 32 % LPT == 0 && LP % G == 0 && LP <= 32 && 32 % LP == 0, "unsupported");

    __shared__ int sH[L], sW[L], sStart[L];
    __shared__ float sInvH[L], sInvW[L];
    stage_levels_inv<L>(shapes, lsi, sH, sW, sStart, sInvH, sInvW);

    const int lane = threadIdx.x & 31;
    const int g = lane / LPT, sub = lane % LPT;
    const unsigned r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= rows_per_image) return;
    const bool pow2 = (M & (M - 1)) == 0;
    const unsigned m = pow2 ? (r & (unsigned)(M - 1)) : (r % (unsigned)M);
    const unsigned q = pow2 ? (r >> (31 - __clz(M))) : (r / (unsigned)M);
    const int MD = M * D;
    const int64_t row = (int64_t)blockIdx.y * rows_per_image + r;
    const int64_t bq = (int64_t)blockIdx.y * Lq + q;
    const T *vimg = opaque_ptr(value + ((int64_t)blockIdx.y * S * M + m) * D + sub * E);

    const int rp = lane % LP;
    const int rl = rp / P;
    float a_soft, x, y;
    fused_point<L, P, RD>(offs, logits, ref, vratio, (int)blockIdx.y, row, bq, rp, rl, sInvH[rl], sInvW[rl], a_soft, x, y);
    const Resolved mine = resolve_point_v(x, y, sH[rl], sW[rl], sStart[rl], a_soft);

    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    if (__all_sync(0xffffffffu, (mine.pm & 15) == 15))
        fwd_row_body<T, D, L, P, true, CH>(mine, vimg, MD, sW, g, acc);
    else
        fwd_row_body<T, D, L, P, false, CH>(mine, vimg, MD, sW, g, acc);
#pragma unroll
    for (int off = LPT; off < 32; off <<= 1) {
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
    }
    if (g == 0) CH::store(out + row * D + sub * E, acc);
}

template <typename T, typename CH, typename ACC, int D, int L, int P, int WARPS, int RD>
__global__ void __launch_bounds__(WARPS * 32, 1024 / (WARPS * 32))
msda_bwd_fused(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
               const float *__restrict__ offs, const float *__restrict__ logits, const float *__restrict__ ref,
               const float *__restrict__ vratio, const T *__restrict__ grad_out, typename ACC::elem *__restrict__ gv_acc,
               const float *__restrict__ det_scale, float *__restrict__ grad_offs, float *__restrict__ grad_logits,
               int S, int M, int Lq, unsigned rows_per_image)
{
    constexpr int E = CH::E;
    constexpr int LPT = D / E;
    constexpr int G = 32 / LPT;
    constexpr int LP = L * P;
    constexpr int PPG = LP / G;
    static_assert(D % E == 0 && 32 % LPT == 0 && LP % G == 0 && LP <= 32 && 32 % LP == 0, "unsupported");
    static_assert(PPG <= LPT && (PPG & (PPG - 1)) == 0, "halving reduction needs PPG to be a power of two <= LPT");

    __shared__ int sH[L], sW[L], sStart[L];
    __shared__ float sInvH[L], sInvW[L];
    stage_levels_inv<L>(shapes, lsi, sH, sW, sStart, sInvH, sInvW);

    const int lane = threadIdx.x & 31;
    const int g = lane / LPT, sub = lane % LPT;
    const unsigned r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= rows_per_image) return;
    const bool pow2 = (M & (M - 1)) == 0;
    const unsigned m = pow2 ? (r & (unsigned)(M - 1)) : (r % (unsigned)M);
    const unsigned q = pow2 ? (r >> (31 - __clz(M))) : (r / (unsigned)M);
    const int MD = M * D;
    const int64_t row = (int64_t)blockIdx.y * rows_per_image + r;
    const int64_t bq = (int64_t)blockIdx.y * Lq + q;
    const int64_t img = ((int64_t)blockIdx.y * S * M + m) * D + sub * E;
    const T *vimg = value + img;
    typename ACC::elem *gimg = gv_acc + img;
    ACC accp;
    if constexpr (sizeof(typename ACC::elem) == 8) accp.scale = __ldg(det_scale);
    accp.template prepare<T, E>(grad_out + row * D, sub, LPT);

    const int rp = lane % LP;
    const int rl = rp / P;
    float a_soft, x, y;
    fused_point<L, P, RD>(offs, logits, ref, vratio, (int)blockIdx.y, row, bq, rp, rl, sInvH[rl], sInvW[rl], a_soft, x, y);
    const Resolved mine = resolve_point_v(x, y, sH[rl], sW[rl], sStart[rl], a_soft);

    float go[E];
    CH::load(grad_out + row * D + sub * E, go);

    float part[3 * PPG];
    if (__all_sync(0xffffffffu, (mine.pm & 15) == 15))
        bwd_row_body<T, CH, ACC, D, L, P, true>(accp, mine, vimg, gimg, MD, sW, g, go, part, 0u);
    else
        bwd_row_body<T, CH, ACC, D, L, P, false>(accp, mine, vimg, gimg, MD, sW, g, go, part, 0u);
    group_reduce3<PPG, LPT>(part, sub);

    // after the reduction lane (g, sub) with sub % SPAN == 0 holds point pt: part[0] = d out/d attn,
    // part[1], part[2] = a * d val / d(w_im, h_im)
    constexpr int SPAN = LPT / PPG;
    const bool holder = (sub % SPAN) == 0;
    const int pt = (sub / SPAN) * G + g;
    const float a_pt = __shfl_sync(0xffffffffu, a_soft, pt);  // the TRUE softmax weight, also for skipped points
    float dot = holder ? a_pt * part[0] : 0.f;                // softmax backward: sum_j a_j * dL/da_j over the row
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    if (holder) {
        const int l = pt / P;
        grad_logits[row * LP + pt] = a_pt * (part[0] - dot);
        float gx = part[1], gy = part[2];  // 2-d: d loc/d off = 1/(W,H) cancels the (W,H) of d(w_im,h_im)/d loc
        if (RD == 4) {
            const float4 rf = load_ref<L, RD>(ref, vratio, bq, (int)blockIdx.y, l);
            gx = (float)sW[l] * gx * (rf.z * 0.5f / (float)P);
            gy = (float)sH[l] * gy * (rf.w * 0.5f / (float)P);
        }
        reinterpret_cast<float2 *>(grad_offs)[row * LP + pt] = make_float2(gx, gy);
    }
}

// Zero the channel rows of masked pixels in place: rows (n_rows, row_elems) of T, mask (n_rows,) of bytes.
template <typename T>
__global__ void msda_mask_rows(T *__restrict__ data, const unsigned char *__restrict__ mask, int64_t n_rows,
                               int row_bytes)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t rowi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rowi < n_rows; rowi += warps) {
        if (!mask[rowi]) continue;
        uint4 *p = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(data) + rowi * row_bytes);
        for (int i = lane; i < row_bytes / 16; i += 32) p[i] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ---- level packing --------------------------------------------------------------------------------
// prepare_od_inputs (models/detection/det_module.py:146-155) turns the per-level NCHW outputs of input_proj into the
// (N, S, C) memory the op reads: `src.flatten(2).transpose(1, 2)` per level, then `torch.cat` -- a strided-read copy.
// msda_pack_levels does it in one launch with a 32x32 shared-memory tile transpose (coalesced 128-byte reads along
// H*W and writes along C); PACK=false is the adjoint (grad of memory -> per-level NCHW grads).
struct PackArgs {
    void *level[8];          // NCHW level tensors (sources when packing, destinations when unpacking)
    int hw[8];               // H_l * W_l
    int start[8];            // level_start_index
    int tile_start[9];       // prefix sum of tiles per level; tile_start[L] = total
    int num_levels;
};

template <typename T, bool PACK>
__global__ void __launch_bounds__(256)
msda_pack_levels(PackArgs args, T *__restrict__ memory, int C, int S)
{
    __shared__ T tile[32][33];
    const int n = blockIdx.z;
    const int c0 = blockIdx.y * 32;
    int l = 0;
    while (l + 1 < args.num_levels && (int)blockIdx.x >= args.tile_start[l + 1]) ++l;
    const int p0 = ((int)blockIdx.x - args.tile_start[l]) * 32;  // first pixel of this tile inside level l
    const int hw = args.hw[l];
    T *lvl = reinterpret_cast<T *>(args.level[l]) + (int64_t)n * C * hw;  // (C, hw) plane of image n
    T *mem = memory + ((int64_t)n * S + args.start[l]) * C;              // (hw, C) rows of image n, level l
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;              // 32 x 8 threads
    if (PACK) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {  // read (c, p) with p fastest
            const int c = c0 + ty + j, p = p0 + tx;
            if (c < C && p < hw) tile[ty + j][tx] = lvl[(int64_t)c * hw + p];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {  // write (p, c) with c fastest
            const int p = p0 + ty + j, c = c0 + tx;
            if (c < C && p < hw) mem[(int64_t)p * C + c] = tile[tx][ty + j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const int p = p0 + ty + j, c = c0 + tx;
            if (c < C && p < hw) tile[tx][ty + j] = mem[(int64_t)p * C + c];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const int c = c0 + ty + j, p = p0 + tx;
            if (c < C && p < hw) lvl[(int64_t)c * hw + p] = tile[ty + j][tx];
        }
    }
}

}  // namespace msda
