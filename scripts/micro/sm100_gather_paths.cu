// sm100_gather_paths.cu -- how fast can one SM obtain random 128-byte tap lines through each path Blackwell offers?
//
// The forward kernel gathers 64 random 128-byte lines per (query, head) row; this microbenchmark measures the line rate
// of every candidate source with the kernels' access shape (8 lanes x 16 bytes per line, 4 lines per warp instruction)
// and no arithmetic:
//   ldg      LDG.E.128 from an L2-resident region                      (what msda_fwd_v5 does)
//   lds      LDS.128 from the CTA's own shared memory                  (what the staged forward does for coarse levels)
//   dsmem8 / dsmem16   ld.shared::cluster.v4 from a random CTA of an 8- / 16-CTA cluster (distributed shared memory:
//            a 16-CTA cluster holds 3.6 MB, enough for one (image, head) pyramid at 800x1333)
//   gather4  TMA cp.async.bulk.tensor.2d.tile::gather4 (UTMALDG) of 4 random rows into shared memory per instruction,
//            consumed with LDS.128                                     (the sm_100-only gather path)
// Output: G lines/s over the whole GPU and bytes/clk/SM-equivalent, one line per path.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/sm100_gather_paths scripts/micro/sm100_gather_paths.cu
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                                    \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) {                                                                 \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16, x *= 0x7feb352dU, x ^= x >> 15, x *= 0x846ca68bU, x ^= x >> 16;
    return x;
}

// index generation must not be what is measured: one LCG step + one multiply-high per line
__device__ __forceinline__ unsigned next_line(unsigned &state, unsigned n)
{
    state = state * 1664525u + 1013904223u;
    return __umulhi(state, n);
}

constexpr int kWarps = 8;

__global__ void __launch_bounds__(kWarps * 32) k_ldg(const float4 *__restrict__ src, float *out, unsigned n_lines, int iters)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
    const unsigned w = blockIdx.x * kWarps + (threadIdx.x >> 5);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned state = hash32(w * 4u + g);
    for (int it = 0; it < iters; it += 8) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned line = next_line(state, n_lines);
            v[k] = __ldg(src + (size_t)line * 8 + sub);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.x += v[k].x, acc.y += v[k].y, acc.z += v[k].z, acc.w += v[k].w;
    }
    if (acc.x == 123.456f) out[w] = acc.x + acc.y + acc.z + acc.w;
}

// CLUSTER == 1: local shared memory; > 1: a random CTA of the cluster
extern __shared__ __align__(128) unsigned char smem_raw[];

template <bool REMOTE>
__global__ void __launch_bounds__(kWarps * 32) k_smem(float *out, unsigned lines_per_cta, int iters)
{
    float4 *sm = reinterpret_cast<float4 *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned nranks = REMOTE ? cluster.num_blocks() : 1u;
    for (unsigned i = threadIdx.x; i < lines_per_cta * 8; i += blockDim.x) sm[i] = make_float4((float)i, 1.f, 2.f, 3.f);
    if (REMOTE)
        cluster.sync();
    else
        __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
    const unsigned w = blockIdx.x * kWarps + (threadIdx.x >> 5);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned state = hash32(w * 4u + g);
    for (int it = 0; it < iters; it += 8) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned line = next_line(state, lines_per_cta);
            const float4 *base = REMOTE ? cluster.map_shared_rank(sm, (state >> 4) & (nranks - 1)) : sm;
            v[k] = base[(size_t)line * 8 + sub];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.x += v[k].x, acc.y += v[k].y, acc.z += v[k].z, acc.w += v[k].w;
    }
    if (REMOTE) cluster.sync();  // nobody leaves while its shared memory may still be read
    if (acc.x == 123.456f) out[w] = acc.x + acc.y + acc.z + acc.w;
}

// ---- TMA gather4 ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kDepth = 4;  // gather4 instructions in flight per warp (each brings 4 lines = 512 bytes)

__global__ void __launch_bounds__(kWarps * 32) k_gather4(const __grid_constant__ CUtensorMap tmap, float *out,
                                                          unsigned n_lines, int iters, int *fail_flag)
{
    // per warp: kDepth buffers of 512 bytes + kDepth mbarriers
    float4 *buf = reinterpret_cast<float4 *>(smem_raw);                                         // [kWarps][kDepth][32]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + kWarps * kDepth * 512);           // [kWarps][kDepth]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned w = blockIdx.x * kWarps + warp;
    float4 *mybuf = buf + warp * kDepth * 32;
    uint64_t *mybar = bars + warp * kDepth;
    if (lane == 0)
        for (int d = 0; d < kDepth; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mybar + d)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    auto issue = [&](int it, int d) {
        if (lane == 0) {
            unsigned st = hash32(w * 9781u + it);
            const unsigned h0 = next_line(st, n_lines), h1 = next_line(st, n_lines);
            const unsigned h2 = next_line(st, n_lines), h3 = next_line(st, n_lines);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(smem_u32(mybar + d)) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
                "%4, %5, %6}], [%7];" ::"r"(smem_u32(mybuf + d * 32)),
                "l"(&tmap), "r"(0), "r"((int)h0), "r"((int)h1), "r"((int)h2), "r"((int)h3), "r"(smem_u32(mybar + d))
                : "memory");
        }
    };
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d = 0; d < kDepth && d < iters; ++d) issue(d, d);
    for (int it = 0; it < iters; ++it) {
        const int d = it % kDepth;
        const unsigned parity = (it / kDepth) & 1;
        unsigned done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(mybar + d)), "r"(parity)
                : "memory");
        if (!done) {  // never hang the box: report and leave
            if (lane == 0) atomicExch(fail_flag, 1);
            return;
        }
        const float4 v = mybuf[d * 32 + lane];  // 4 lines x 8 lanes x 16 bytes
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        __syncwarp();
        if (it + kDepth < iters) issue(it + kDepth, d);
    }
    if (acc.x == 123.456f) out[w] = acc.x + acc.y + acc.z + acc.w;
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    int dev = 0, sms = 0, clock_khz = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
    const unsigned n_lines = 22223u * 8u;  // one image of value at 800x1333, D=32 fp32: 22.8 MB, L2-resident
    float *region = nullptr, *out = nullptr;
    int *fail_flag = nullptr;
    CK(cudaMalloc(&region, (size_t)n_lines * 128));
    CK(cudaMemset(region, 0, (size_t)n_lines * 128));
    CK(cudaMalloc(&out, 1 << 22));
    CK(cudaMalloc(&fail_flag, 4));
    CK(cudaMemset(fail_flag, 0, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 2048;

    auto report = [&](const char *name, double lines, float ms, int active_sms) {
        const double glps = lines / ms / 1e6;
        const double bytes_per_clk_sm = lines * 128.0 / (ms * 1e-3) / ((double)clock_khz * 1e3) / active_sms;
        printf("%-10s %8.2f G lines/s  %7.2f TB/s  %6.1f B/clk/SM (at %.0f MHz nominal, %d SMs)  %.3f ms\n", name, glps,
               glps * 128 / 1e3, bytes_per_clk_sm, clock_khz / 1e3, active_sms, ms);
    };

    {  // ldg
        const int blocks = sms * 8;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            k_ldg<<<blocks, kWarps * 32>>>((const float4 *)region, out, n_lines, iters);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        report("ldg", (double)blocks * kWarps * iters * 4, ms, sms);
    }
    const int smem_bytes = 64 * 1024;  // 512 lines per CTA; three CTAs (24 warps) per SM
    const unsigned lines_per_cta = smem_bytes / 128;
    {  // lds
        CK(cudaFuncSetAttribute(k_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        const int blocks = sms * 3;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            k_smem<false><<<blocks, kWarps * 32, smem_bytes>>>(out, lines_per_cta, iters * 8);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        report("lds", (double)blocks * kWarps * iters * 8 * 4, ms, sms);
    }
    for (int cluster : {8, 16}) {  // dsmem
        CK(cudaFuncSetAttribute(k_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        if (cluster > 8) CK(cudaFuncSetAttribute(k_smem<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        cfg.blockDim = dim3(kWarps * 32), cfg.dynamicSmemBytes = smem_bytes;
        int max_clusters = 0;
        cfg.gridDim = dim3(cluster);
        cudaError_t qe = cudaOccupancyMaxActiveClusters(&max_clusters, k_smem<true>, &cfg);
        if (qe != cudaSuccess || max_clusters < 1) {
            printf("dsmem%-5d not launchable (%s, max active clusters %d)\n", cluster, cudaGetErrorString(qe), max_clusters);
            cudaGetLastError();
            continue;
        }
        const int blocks = max_clusters * cluster;
        cfg.gridDim = dim3(blocks);
        float ms = 0;
        bool ok = true;
        for (int rep = 0; rep < 2 && ok; ++rep) {
            CK(cudaEventRecord(e0));
            cudaError_t le = cudaLaunchKernelEx(&cfg, k_smem<true>, out, lines_per_cta, iters);
            CK(cudaEventRecord(e1));
            if (le != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
                printf("dsmem%-5d launch failed: %s\n", cluster, cudaGetErrorString(le));
                cudaGetLastError();
                ok = false;
            }
        }
        if (ok) {
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char name[32];
            snprintf(name, sizeof(name), "dsmem%d", cluster);
            report(name, (double)blocks * kWarps * iters * 4, ms, blocks);
        }
    }
    {  // TMA gather4
        EncodeTiled encode = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres);
        if (ge != cudaSuccess || !encode) {
            printf("gather4    cuTensorMapEncodeTiled not available (%s)\n", cudaGetErrorString(ge));
        } else {
            CUtensorMap tmap;
            const cuuint64_t gdim[2] = {32, n_lines};   // 32 floats per row, one row per (pixel, head) line
            const cuuint64_t gstride[1] = {128};        // bytes between rows
            const cuuint32_t box[2] = {32, 1};          // gather4 fetches four such one-row boxes
            const cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, region, gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                printf("gather4    cuTensorMapEncodeTiled failed with CUresult %d\n", (int)r);
            } else {
                const int smem = kWarps * kDepth * 512 + kWarps * kDepth * 8;
                const int blocks = sms * 8;
                float ms = 0;
                bool ok = true;
                for (int rep = 0; rep < 2 && ok; ++rep) {
                    CK(cudaEventRecord(e0));
                    k_gather4<<<blocks, kWarps * 32, smem>>>(tmap, out, n_lines, iters, fail_flag);
                    CK(cudaEventRecord(e1));
                    cudaError_t se = cudaDeviceSynchronize();
                    int failed = 0;
                    cudaMemcpy(&failed, fail_flag, 4, cudaMemcpyDeviceToHost);
                    if (se != cudaSuccess || failed) {
                        printf("gather4    kernel failed (%s, timeout flag %d)\n", cudaGetErrorString(se), failed);
                        ok = false;
                    }
                }
                if (ok) {
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    report("gather4", (double)blocks * kWarps * iters * 4, ms, sms);
                }
            }
        }
    }
    return 0;
}
