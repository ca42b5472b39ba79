/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for multi-scale deformable attention.
 *
 * This header is included twice by msda_oracle.c, once with REAL=double and once with
 * REAL=float, to stamp out the two precisions.  Nothing under grit_b200/ may include,
 * link or call this code: it exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py have something independent to check
 * (and time) the CUDA path against.
 *
 * Semantics restated from the reference (paths relative to /root/reference):
 *   - sample position and validity window:
 *       models/ops/src/cuda/ms_deform_im2col_cuda.cuh:272-296
 *       (h_im = y*H - 0.5, w_im = x*W - 0.5; a point counts only if
 *        h_im > -1 && w_im > -1 && h_im < H && w_im < W)
 *   - bilinear taps with per-tap zero padding:
 *       models/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-84
 *   - analytic gradients (grad_value scatter, grad_sampling_loc, grad_attn_weight):
 *       models/ops/src/cuda/ms_deform_im2col_cuda.cuh:87-159
 *   - which equals the grid_sample composition of
 *       models/ops/functions/ms_deform_attn_func.py:41-61
 *       (grid = 2*loc-1, align_corners=False, padding_mode='zeros').
 *
 * Layouts (all contiguous, row-major):
 *   value   (N, S, M, D)        sampling_loc (N, Lq, M, L, P, 2)  -- (x, y) in [0,1]
 *   attn    (N, Lq, M, L, P)    out / grad_out (N, Lq, M, D)
 *   shapes  (L, 2) int64 [H, W] level_start (L,) int64
 */

#ifndef REAL
#error "include from msda_oracle.c with REAL and SUFFIX defined"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* One sample point resolved to its four taps.  off[i] < 0 means "tap outside the map". */
typedef struct {
    int live;        /* 0: point is outside the validity window, contributes nothing */
    int64_t off[4];  /* pixel index inside the level (row*W + col), or -1           */
    REAL wt[4];      /* bilinear weights  (tl, tr, bl, br)                           */
    REAL lh, lw;     /* fractional parts, needed by the location gradient            */
} FN(tap_set);

static inline void FN(resolve_taps)(REAL x, REAL y, int64_t H, int64_t W, FN(tap_set) * t)
{
    const REAL h_im = y * (REAL)H - (REAL)0.5;
    const REAL w_im = x * (REAL)W - (REAL)0.5;
    t->live = (h_im > (REAL)-1 && w_im > (REAL)-1 && h_im < (REAL)H && w_im < (REAL)W);
    if (!t->live) return; /* NaN coordinates fail every comparison and land here too */

    const int64_t r0 = (int64_t)FLOOR(h_im), c0 = (int64_t)FLOOR(w_im);
    const int64_t r1 = r0 + 1, c1 = c0 + 1;
    const REAL lh = h_im - (REAL)r0, lw = w_im - (REAL)c0;
    const REAL hh = (REAL)1 - lh, hw = (REAL)1 - lw;
    const int top = r0 >= 0, bot = r1 <= H - 1, lef = c0 >= 0, rig = c1 <= W - 1;

    t->off[0] = (top && lef) ? r0 * W + c0 : -1;
    t->off[1] = (top && rig) ? r0 * W + c1 : -1;
    t->off[2] = (bot && lef) ? r1 * W + c0 : -1;
    t->off[3] = (bot && rig) ? r1 * W + c1 : -1;
    t->wt[0] = hh * hw;
    t->wt[1] = hh * lw;
    t->wt[2] = lh * hw;
    t->wt[3] = lh * lw;
    t->lh = lh;
    t->lw = lw;
}

/* out[b,q,m,:] = sum_{l,p} attn * bilinear(value_l, loc)  */
void FN(msda_oracle_forward)(const REAL *value, const int64_t *shapes, const int64_t *level_start,
                             const REAL *loc, const REAL *attn, REAL *out, int64_t N, int64_t S,
                             int64_t M, int64_t D, int64_t L, int64_t Lq, int64_t P)
{
    const int64_t rows = N * Lq * M;
#pragma omp parallel for schedule(static)
    for (int64_t row = 0; row < rows; ++row) {
        const int64_t m = row % M;
        const int64_t b = row / (M * Lq);
        REAL *o = out + row * D;
        for (int64_t c = 0; c < D; ++c) o[c] = 0;
        const REAL *lp = loc + row * L * P * 2;
        const REAL *ap = attn + row * L * P;
        for (int64_t l = 0; l < L; ++l) {
            const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];
            const REAL *plane = value + ((b * S + level_start[l]) * M + m) * D;
            for (int64_t p = 0; p < P; ++p) {
                FN(tap_set) t;
                FN(resolve_taps)(lp[(l * P + p) * 2], lp[(l * P + p) * 2 + 1], H, W, &t);
                if (!t.live) continue;
                const REAL a = ap[l * P + p];
                for (int i = 0; i < 4; ++i) {
                    if (t.off[i] < 0) continue;
                    const REAL *v = plane + t.off[i] * M * D;
                    const REAL s = a * t.wt[i];
                    for (int64_t c = 0; c < D; ++c) o[c] += s * v[c];
                }
            }
        }
    }
}

/*
 * grad_value must be zero-filled by the caller; grad_loc and grad_attn are fully written.
 * Parallel over (image, head): every grad_value element belongs to exactly one such pair,
 * so the scatter needs no atomics and the summation order is fixed (query-major).
 */
void FN(msda_oracle_backward)(const REAL *value, const int64_t *shapes, const int64_t *level_start,
                              const REAL *loc, const REAL *attn, const REAL *grad_out,
                              REAL *grad_value, REAL *grad_loc, REAL *grad_attn, int64_t N,
                              int64_t S, int64_t M, int64_t D, int64_t L, int64_t Lq, int64_t P)
{
    const int64_t pairs = N * M;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t pair = 0; pair < pairs; ++pair) {
        const int64_t b = pair / M, m = pair % M;
        for (int64_t q = 0; q < Lq; ++q) {
            const int64_t row = (b * Lq + q) * M + m;
            const REAL *g = grad_out + row * D;
            const REAL *lp = loc + row * L * P * 2;
            const REAL *ap = attn + row * L * P;
            REAL *glp = grad_loc + row * L * P * 2;
            REAL *gap = grad_attn + row * L * P;
            for (int64_t l = 0; l < L; ++l) {
                const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];
                const int64_t plane_off = ((b * S + level_start[l]) * M + m) * D;
                for (int64_t p = 0; p < P; ++p) {
                    const int64_t k = l * P + p;
                    glp[2 * k] = glp[2 * k + 1] = gap[k] = 0;
                    FN(tap_set) t;
                    FN(resolve_taps)(lp[2 * k], lp[2 * k + 1], H, W, &t);
                    if (!t.live) continue;
                    const REAL a = ap[k];
                    const REAL hh = (REAL)1 - t.lh, hw = (REAL)1 - t.lw;
                    /* d(val)/d(w_im) and d(val)/d(h_im) tap coefficients */
                    const REAL dx[4] = {-hh, hh, -t.lh, t.lh};
                    const REAL dy[4] = {-hw, -t.lw, hw, t.lw};
                    REAL s_attn = 0, s_x = 0, s_y = 0;
                    for (int i = 0; i < 4; ++i) {
                        if (t.off[i] < 0) continue;
                        const REAL *v = value + plane_off + t.off[i] * M * D;
                        REAL *gv = grad_value + plane_off + t.off[i] * M * D;
                        const REAL s = a * t.wt[i];
                        REAL dot = 0;
                        for (int64_t c = 0; c < D; ++c) {
                            dot += g[c] * v[c];
                            gv[c] += s * g[c];
                        }
                        s_attn += t.wt[i] * dot;
                        s_x += dx[i] * dot;
                        s_y += dy[i] * dot;
                    }
                    gap[k] = s_attn;
                    glp[2 * k] = (REAL)W * a * s_x;
                    glp[2 * k + 1] = (REAL)H * a * s_y;
                }
            }
        }
    }
}

#undef FN
#undef CAT
#undef CAT_
