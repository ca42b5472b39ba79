// smem_accumulate.cu -- can shared memory absorb the backward's coarse-level gradient lines faster than REDG?
//
// VERDICT r1 item 2, candidate B: keep the fp32 gradient plane of the coarse levels (1 323 px x 128 B = 169 KB per
// (image, head) at 800x1333) in shared memory as integer fixed point and add every tap line to it with NATIVE integer
// shared-memory atomics (ATOMS.ADD), flushing one REDG per pixel at the end.  Before building the kernel this measures
// the only number that decides it: lines (32 channels x 4 B) per second an SM can add into a random line of such a plane
//   atoms32   atomicAdd(int *)                 32 lanes x 4 B, one line per warp instruction (ATOMS.ADD)
//   atoms64   atomicAdd(unsigned long long *)  32 lanes x 8 B = two lines per warp instruction (two packed int32)
//   atomsf32  atomicAdd(float *) on shared memory (what ptxas makes of it on sm_100a)
//   rmw       LDS.128 + 4 FADD + STS.128, 8 lanes per line, NOT atomic (the LSU floor; needs exclusive ownership)
//   redg      red.global.add.v4.f32 into an L2-resident region (what msda_bwd_v5 does today)
// Output: G lines/s over the whole GPU, one line per path.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/smem_accumulate scripts/micro/smem_accumulate.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16, x *= 0x7feb352dU, x ^= x >> 15, x *= 0x846ca68bU, x ^= x >> 16;
    return x;
}
__device__ __forceinline__ unsigned next_line(unsigned &state, unsigned n)
{
    state = state * 1664525u + 1013904223u;
    return __umulhi(state, n);
}

extern __shared__ __align__(16) unsigned char smem_raw[];

__global__ void __launch_bounds__(1024) k_atoms32(int *out, unsigned n_lines, int iters)
{
    int *plane = reinterpret_cast<int *>(smem_raw);
    for (unsigned i = threadIdx.x; i < n_lines * 32; i += blockDim.x) plane[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned state = hash32(blockIdx.x * 32u + (threadIdx.x >> 5));  // warp-uniform line choice
    for (int it = 0; it < iters; ++it) {
        const unsigned line = next_line(state, n_lines);
        atomicAdd(plane + line * 32 + lane, it + lane);
    }
    __syncthreads();
    if (plane[threadIdx.x] == 0x7fffffff) out[blockIdx.x] = 1;
}

__global__ void __launch_bounds__(1024) k_atoms64(int *out, unsigned n_lines, int iters)
{
    unsigned long long *plane = reinterpret_cast<unsigned long long *>(smem_raw);
    for (unsigned i = threadIdx.x; i < n_lines * 16; i += blockDim.x) plane[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, half = lane >> 4, sub = lane & 15;
    unsigned state = hash32((blockIdx.x * 32u + (threadIdx.x >> 5)) * 2u + half);  // two lines per warp instruction
    for (int it = 0; it < iters; ++it) {
        const unsigned line = next_line(state, n_lines);
        atomicAdd(plane + line * 16 + sub, (unsigned long long)(it + lane) * 0x100000001ull);
    }
    __syncthreads();
    if (plane[threadIdx.x] == 0x7fffffffull) out[blockIdx.x] = 1;
}

__global__ void __launch_bounds__(1024) k_atomsf32(int *out, unsigned n_lines, int iters)
{
    float *plane = reinterpret_cast<float *>(smem_raw);
    for (unsigned i = threadIdx.x; i < n_lines * 32; i += blockDim.x) plane[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned state = hash32(blockIdx.x * 32u + (threadIdx.x >> 5));
    for (int it = 0; it < iters; ++it) {
        const unsigned line = next_line(state, n_lines);
        atomicAdd(plane + line * 32 + lane, 1.0f + lane);
    }
    __syncthreads();
    if (plane[threadIdx.x] == -1.f) out[blockIdx.x] = 1;
}

__global__ void __launch_bounds__(1024) k_rmw(int *out, unsigned n_lines, int iters)
{
    float4 *plane = reinterpret_cast<float4 *>(smem_raw);
    for (unsigned i = threadIdx.x; i < n_lines * 8; i += blockDim.x) plane[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
    unsigned state = hash32((blockIdx.x * 32u + (threadIdx.x >> 5)) * 4u + g);  // four lines per warp instruction
    for (int it = 0; it < iters; ++it) {
        const unsigned line = next_line(state, n_lines);
        float4 v = plane[line * 8 + sub];
        v.x += 1.f, v.y += 2.f, v.z += 3.f, v.w += 4.f;
        plane[line * 8 + sub] = v;  // racy on purpose: this is the floor, not a usable accumulator
    }
    __syncthreads();
    if (plane[threadIdx.x].x == -1.f) out[blockIdx.x] = 1;
}

__global__ void __launch_bounds__(1024) k_redg(float *dst, unsigned n_lines, int iters)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
    unsigned state = hash32((blockIdx.x * 32u + (threadIdx.x >> 5)) * 4u + g);
    for (int it = 0; it < iters; ++it) {
        const unsigned line = next_line(state, n_lines);
        float *p = dst + (size_t)line * 32 + sub * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f)
                     : "memory");
    }
}

template <typename F>
static float time_ms(F &&launch, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const unsigned n_lines = 1323;  // levels 2+3 of the 800x1333 pyramid
    const size_t smem = (size_t)n_lines * 128;
    const int iters = 4096, reps = 5;
    int *out;
    float *gdst;
    const unsigned g_lines = 16u * 22223u * 8u / 16u;  // ~23 MB: one image's grad_value, L2-resident
    CK(cudaMalloc(&out, sms * sizeof(int)));
    CK(cudaMalloc(&gdst, (size_t)g_lines * 128));
    CK(cudaMemset(gdst, 0, (size_t)g_lines * 128));
    CK(cudaFuncSetAttribute(k_atoms32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_atoms64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_atomsf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    printf("device %s, %d SMs, plane %u lines (%zu KB) per CTA, 1 CTA of 1024 threads per SM\n", prop.name, sms, n_lines,
           smem / 1024);
    const double warps = (double)sms * 32.0;
    struct {
        const char *name;
        double lines_per_warp_iter;
        float ms;
    } r[5];
    r[0] = {"atoms32  (ATOMS.ADD, 32 lanes x 4 B)", 1.0, time_ms([&] { k_atoms32<<<sms, 1024, smem>>>(out, n_lines, iters); }, reps)};
    r[1] = {"atoms64  (ATOMS.ADD.64, 2 lines/instr)", 2.0, time_ms([&] { k_atoms64<<<sms, 1024, smem>>>(out, n_lines, iters); }, reps)};
    r[2] = {"atomsf32 (atomicAdd(float) on smem)", 1.0, time_ms([&] { k_atomsf32<<<sms, 1024, smem>>>(out, n_lines, iters / 4); }, reps) * 4.f};
    r[3] = {"rmw      (LDS.128+FADD+STS.128, racy)", 4.0, time_ms([&] { k_rmw<<<sms, 1024, smem>>>(out, n_lines, iters); }, reps)};
    r[4] = {"redg     (red.global.add.v4.f32, L2)", 4.0, time_ms([&] { k_redg<<<sms, 1024>>>(gdst, g_lines, iters); }, reps)};
    CK(cudaDeviceSynchronize());
    for (auto &x : r) {
        const double lines = warps * iters * x.lines_per_warp_iter;
        printf("%-42s %8.3f ms  %8.1f G lines/s  %6.2f TB/s\n", x.name, x.ms, lines / x.ms * 1e-6, lines * 128 / x.ms * 1e-9);
    }
    return 0;
}
