// Micro-benchmark: scatter-add of 128-byte fp32 lines to random global addresses, two ways.
//  (a) red.global.add.v4.f32 from registers (what msda_bwd_v5 does)
//  (b) stage in shared memory, cp.reduce.async.bulk.global.shared::cta.add.f32 (TMA bulk reduction)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_vs_tma red_vs_tma.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// each warp performs `iters` iterations; per iteration 4 lines (one per 8-lane group), like one point of the bwd
__global__ void k_red(float* dst, unsigned n_lines, int iters)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (int it = 0; it < iters; ++it) {
        const unsigned line = hash32(w * 9781u + it * 4u + g) % n_lines;
        float* p = dst + (size_t)line * 32 + sub * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    }
}

__global__ void k_tma(float* dst, unsigned n_lines, int iters)
{
    extern __shared__ __align__(128) float smem[];  // per warp: 2 buffers x 4 lines x 32 floats
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7, warp = threadIdx.x >> 5;
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + warp;
    float* wbuf = smem + warp * 2 * 128;
    for (int it = 0; it < iters; ++it) {
        float* buf = wbuf + (it & 1) * 128;
        // the buffer used two iterations ago must have been read by the TMA engine
        if (lane < 4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        *reinterpret_cast<float4*>(buf + g * 32 + sub * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane < 4) {
            const unsigned line = hash32(w * 9781u + it * 4u + lane) % n_lines;
            float* p = dst + (size_t)line * 32;
            const unsigned s = (unsigned)__cvta_generic_to_shared(buf + lane * 32);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(p), "r"(s) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// (d) both egress paths at once: even warps use register reds, odd warps the TMA bulk reduction
__global__ void k_mixed(float* dst, unsigned n_lines, int iters, int tma_every)
{
    extern __shared__ __align__(128) float smem[];
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7, warp = threadIdx.x >> 5;
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + warp;
    float* wbuf = smem + warp * 2 * 128;
    for (int it = 0; it < iters; ++it) {
        if (tma_every > 0 && (it % tma_every) == 0) {
            float* buf = wbuf + ((it / tma_every) & 1) * 128;
            if (lane < 4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            *reinterpret_cast<float4*>(buf + g * 32 + sub * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane < 4) {
                const unsigned line = hash32(w * 9781u + it * 4u + lane) % n_lines;
                float* p = dst + (size_t)line * 32;
                const unsigned s = (unsigned)__cvta_generic_to_shared(buf + lane * 32);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(p), "r"(s) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            const unsigned line = hash32(w * 9781u + it * 4u + g) % n_lines;
            float* p = dst + (size_t)line * 32 + sub * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
        }
    }
    if (lane < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// (c) pure gather: LDG.E.128 of random 128-byte lines (4 per warp instruction, 16 instructions in flight per lane),
//     the access pattern of the forward without any of its arithmetic
__global__ void k_gather(const float* __restrict__ src, float* out, unsigned n_lines, int iters)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; it += 16) {
        float4 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const unsigned line = hash32(w * 9781u + (it + k) * 4u + g) % n_lines;
            v[k] = __ldg(reinterpret_cast<const float4*>(src + (size_t)line * 32 + sub * 4));
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) acc.x += v[k].x, acc.y += v[k].y, acc.z += v[k].z, acc.w += v[k].w;
    }
    if (acc.x == 123.456f) out[w] = acc.x + acc.y + acc.z + acc.w;
}

int main()
{
    const unsigned n_lines = 22223u * 8u;  // one image of grad_value lines (22.8 MB): L2-resident like the real backward
    float* dst;
    cudaMalloc(&dst, (size_t)n_lines * 128);
    cudaMemset(dst, 0, (size_t)n_lines * 128);
    const int warps_per_block = 4, blocks = 148 * 8 * 4, iters = 1024;  // 16 lines... total lines = blocks*4warps*iters*4
    const double lines = (double)blocks * warps_per_block * iters * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_red<<<blocks, warps_per_block * 32>>>(dst, n_lines, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("red.v4.f32      : %.3f ms  %.1f G lines/s  %.2f TB/s payload  (%s)\n", ms, lines / ms / 1e6, lines * 128 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_tma<<<blocks, warps_per_block * 32, warps_per_block * 2 * 128 * sizeof(float)>>>(dst, n_lines, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("cp.reduce.bulk  : %.3f ms  %.1f G lines/s  %.2f TB/s payload  (%s)\n", ms, lines / ms / 1e6, lines * 128 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    // (e) where is the reduction ceiling: SM egress or L2?  Same stream from 37 / 74 / 148 SMs (one 1024-thread CTA each)
    for (int sms = 37; sms <= 148; sms *= 2) {
        const int it2 = 4096;
        const double lines2 = (double)sms * 32 * it2 * 4;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_red<<<sms, 1024>>>(dst, n_lines, it2);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("red.v4.f32 from %3d SMs: %.3f ms  %.1f G lines/s  = %.2f lines/us/SM\n", sms, ms, lines2 / ms / 1e6, lines2 / ms / 1e3 / sms);
    }
    for (int every = 2; every <= 4; ++every)
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_mixed<<<blocks, warps_per_block * 32, warps_per_block * 2 * 128 * sizeof(float)>>>(dst, n_lines, iters, every);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("mixed 1/%d via TMA: %.3f ms  %.1f G lines/s  %.2f TB/s payload  (%s)\n", every, ms, lines / ms / 1e6, lines * 128 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_gather<<<blocks, warps_per_block * 32>>>(dst, dst, n_lines, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("ldg.128 gather  : %.3f ms  %.1f G lines/s  %.2f TB/s payload  (%s)\n", ms, lines / ms / 1e6, lines * 128 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
