// Pure C/C++ client of the C ABI (no Python, no torch): what a non-Python host of the reference would link.
// Allocates with cudaMalloc, runs msda_forward / msda_backward from include/msda.h and checks them against the CPU
// oracle (oracle/libmsda_oracle.so, test infrastructure).  Built and run by tests/test_c_abi_client.py on the GPU box:
//   nvcc -O2 -I include -o c_abi_parity tests/c_abi/c_abi_parity.cu -L grit_b200 -lmsda_b200 -L oracle -lmsda_oracle
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "msda.h"

extern "C" {
void msda_oracle_forward_f64(const double *, const int64_t *, const int64_t *, const double *, const double *, double *,
                             int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t);
void msda_oracle_backward_f64(const double *, const int64_t *, const int64_t *, const double *, const double *,
                              const double *, double *, double *, double *, int64_t, int64_t, int64_t, int64_t, int64_t,
                              int64_t, int64_t);
}

static double urand(unsigned &s)
{
    s = s * 1664525u + 1013904223u;
    return (s >> 8) * (1.0 / 16777216.0);
}

template <typename V>
static V *to_device(const std::vector<V> &h)
{
    V *d = nullptr;
    cudaMalloc(&d, h.size() * sizeof(V));
    cudaMemcpy(d, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice);
    return d;
}

static double max_norm_err(const std::vector<float> &got, const std::vector<double> &ref)
{
    double e = 0, m = 1e-300;
    for (size_t i = 0; i < ref.size(); ++i) {
        e = fmax(e, fabs((double)got[i] - ref[i]));
        m = fmax(m, fabs(ref[i]));
    }
    return e / m;
}

int main()
{
    const int64_t N = 2, M = 8, D = 32, L = 4, P = 4, Lq = 333;
    const int64_t shapes_h[8] = {20, 30, 10, 15, 5, 8, 3, 4};
    int64_t lsi_h[4], S = 0;
    for (int l = 0; l < L; ++l) lsi_h[l] = S, S += shapes_h[2 * l] * shapes_h[2 * l + 1];
    msda_dims dims = {N, S, M, D, L, Lq, P};
    unsigned seed = 12345u;
    std::vector<float> value(N * S * M * D), loc(N * Lq * M * L * P * 2), attn(N * Lq * M * L * P), gout(N * Lq * M * D);
    for (auto &v : value) v = (float)(urand(seed) * 2 - 1);
    for (auto &v : loc) v = (float)(urand(seed) * 1.2 - 0.1);
    for (auto &v : gout) v = (float)(urand(seed) * 2 - 1);
    for (size_t r = 0; r < attn.size(); r += L * P) {
        double sum = 0;
        for (int k = 0; k < L * P; ++k) sum += (attn[r + k] = (float)(urand(seed) + 0.01));
        for (int k = 0; k < L * P; ++k) attn[r + k] = (float)(attn[r + k] / sum);
    }
    // oracle in fp64 on the same fp32 numbers
    std::vector<double> v64(value.begin(), value.end()), l64(loc.begin(), loc.end()), a64(attn.begin(), attn.end()),
        g64(gout.begin(), gout.end()), out_ref(N * Lq * M * D), gv_ref(value.size(), 0.0), gl_ref(loc.size()),
        ga_ref(attn.size());
    msda_oracle_forward_f64(v64.data(), shapes_h, lsi_h, l64.data(), a64.data(), out_ref.data(), N, S, M, D, L, Lq, P);
    msda_oracle_backward_f64(v64.data(), shapes_h, lsi_h, l64.data(), a64.data(), g64.data(), gv_ref.data(),
                             gl_ref.data(), ga_ref.data(), N, S, M, D, L, Lq, P);

    float *d_value = to_device(value), *d_loc = to_device(loc), *d_attn = to_device(attn), *d_gout = to_device(gout);
    std::vector<int64_t> sh(shapes_h, shapes_h + 8), ls(lsi_h, lsi_h + 4);
    int64_t *d_shapes = to_device(sh), *d_lsi = to_device(ls);
    float *d_out, *d_gv, *d_gl, *d_ga;
    cudaMalloc(&d_out, out_ref.size() * 4), cudaMalloc(&d_gv, value.size() * 4);
    cudaMalloc(&d_gl, loc.size() * 4), cudaMalloc(&d_ga, attn.size() * 4);
    cudaStream_t st;
    cudaStreamCreate(&st);

    int rc = msda_forward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_out, &dims, MSDA_F32, 0, st);
    if (rc) return printf("msda_forward failed: %s\n", msda_last_error()), 1;
    printf("forward kernel : %s\n", msda_last_kernel());
    rc = msda_backward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_gout, d_gv, d_gl, d_ga, &dims, MSDA_F32,
                       MSDA_FLAG_ZERO_GRAD_VALUE, nullptr, 0, st);
    if (rc) return printf("msda_backward failed: %s\n", msda_last_error()), 1;
    printf("backward kernel: %s\n", msda_last_kernel());
    if (cudaStreamSynchronize(st) != cudaSuccess) return printf("stream sync failed\n"), 1;

    std::vector<float> out(out_ref.size()), gv(value.size()), gl(loc.size()), ga(attn.size());
    cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(gv.data(), d_gv, gv.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(gl.data(), d_gl, gl.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ga.data(), d_ga, ga.size() * 4, cudaMemcpyDeviceToHost);
    const double e_out = max_norm_err(out, out_ref), e_gv = max_norm_err(gv, gv_ref), e_ga = max_norm_err(ga, ga_ref);
    const double e_gl = max_norm_err(gl, gl_ref);  // (no sample of this seed sits on an integer pixel coordinate)
    printf("max-normalised error: out %.2e grad_value %.2e grad_loc %.2e grad_attn %.2e\n", e_out, e_gv, e_gl, e_ga);

    // error paths: bad dtype, null pointer, missing workspace
    int bad = 0;
    bad += msda_forward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_out, &dims, 9, 0, st) != MSDA_ERR_INVALID_ARGUMENT;
    bad += msda_forward(nullptr, d_shapes, d_lsi, d_loc, d_attn, d_out, &dims, MSDA_F32, 0, st) != MSDA_ERR_INVALID_ARGUMENT;
    bad += msda_backward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_gout, d_gv, d_gl, d_ga, &dims, MSDA_BF16, 0, nullptr, 0,
                         st) != MSDA_ERR_WORKSPACE;
    printf("error-path checks failed: %d (last message: %s)\n", bad, msda_last_error());
    const bool ok = e_out < 1e-5 && e_gv < 1e-4 && e_gl < 1e-4 && e_ga < 1e-4 && bad == 0;
    printf(ok ? "C ABI PARITY OK\n" : "C ABI PARITY FAILED\n");
    return ok ? 0 : 1;
}
