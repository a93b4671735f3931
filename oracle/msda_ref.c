/* ORACLE -- test infrastructure, not product code.
 *
 * Plain-C restatement of the multi-scale deformable attention core of raphael-baena/DTLR
 * (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this).
 *
 * Follows, by reading (no code copied):
 *   forward   models/dino/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299  (ms_deformable_im2col_gpu_kernel)
 *   bilinear  models/dino/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-84    (ms_deform_attn_im2col_bilinear)
 *   backward  models/dino/ops/src/cuda/ms_deform_im2col_cuda.cuh:87-159, 301-403 (col2im bilinear + reduce_v1)
 * and is pinned against the reference's own CPU statement of the same op,
 *   models/dino/ops/functions/ms_deform_attn_func.py:41-61 (ms_deform_attn_core_pytorch),
 * on the fixture of models/dino/ops/test.py:21-28 (tests/golden/msda_kat.npz, tests/test_oracle_msda.py).
 *
 * Semantics: value (B,S,M,D); shapes (L,2) = (H_l,W_l) int64; lsi (L) int64; loc (B,Lq,M,L,P,2) = (x,y) in [0,1];
 * w (B,Lq,M,L,P); out (B,Lq,M*D).  Pixel coords: x = loc_x*W - 0.5, y = loc_y*H - 0.5 (align_corners=False);
 * a point contributes iff -1 < y < H and -1 < x < W; corners outside the map read as zero.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_MSDA(T, SUF)                                                                                   \
    void msda_ref_fwd_##SUF(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc,         \
                            const T* w, T* out, int B, int S, int M, int D, int L, int Lq, int P)             \
    {                                                                                                         \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                              \
        for (int b = 0; b < B; ++b)                                                                           \
            for (int q = 0; q < Lq; ++q)                                                                      \
                for (int m = 0; m < M; ++m) {                                                                 \
                    T* o = out + (((int64_t)b * Lq + q) * M + m) * D;                                         \
                    for (int c = 0; c < D; ++c) o[c] = 0;                                                     \
                    const int64_t pw = (((int64_t)b * Lq + q) * M + m) * L * P;                               \
                    for (int l = 0; l < L; ++l) {                                                             \
                        const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                         \
                        const T* vl = value + ((int64_t)b * S + lsi[l]) * M * D;                              \
                        for (int p = 0; p < P; ++p) {                                                         \
                            const T lx = loc[(pw + l * P + p) * 2], ly = loc[(pw + l * P + p) * 2 + 1];       \
                            const T aw = w[pw + l * P + p];                                                   \
                            const T y = ly * H - (T)0.5, x = lx * W - (T)0.5;                                 \
                            if (!(y > -1 && x > -1 && y < H && x < W)) continue;                              \
                            const int y0 = (int)floor((double)y), x0 = (int)floor((double)x);                 \
                            const T fy = y - y0, fx = x - x0, gy = 1 - fy, gx = 1 - fx;                       \
                            const T cw[4] = {gy * gx, gy * fx, fy * gx, fy * fx};                             \
                            const int cy[4] = {y0, y0, y0 + 1, y0 + 1}, cx[4] = {x0, x0 + 1, x0, x0 + 1};     \
                            for (int c = 0; c < D; ++c) {                                                     \
                                T v[4];                                                                       \
                                for (int k = 0; k < 4; ++k)                                                   \
                                    v[k] = (cy[k] >= 0 && cy[k] <= H - 1 && cx[k] >= 0 && cx[k] <= W - 1)     \
                                               ? vl[((int64_t)(cy[k] * W + cx[k]) * M + m) * D + c]           \
                                               : (T)0;                                                        \
                                o[c] += aw * (cw[0] * v[0] + cw[1] * v[1] + cw[2] * v[2] + cw[3] * v[3]);     \
                            }                                                                                 \
                        }                                                                                     \
                    }                                                                                         \
                }                                                                                             \
    }                                                                                                         \
                                                                                                              \
    /* grad_value must be zero-filled by the caller; serial over (b,q,m) so that the scatter is race free */ \
    void msda_ref_bwd_##SUF(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc,         \
                            const T* w, const T* gout, T* gvalue, T* gloc, T* gw, int B, int S, int M,        \
                            int D, int L, int Lq, int P)                                                      \
    {                                                                                                         \
        memset(gvalue, 0, sizeof(T) * (size_t)B * S * M * D);                                                 \
        for (int b = 0; b < B; ++b)                                                                           \
            for (int q = 0; q < Lq; ++q)                                                                      \
                for (int m = 0; m < M; ++m) {                                                                 \
                    const T* go = gout + (((int64_t)b * Lq + q) * M + m) * D;                                 \
                    const int64_t pw = (((int64_t)b * Lq + q) * M + m) * L * P;                               \
                    for (int l = 0; l < L; ++l) {                                                             \
                        const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                         \
                        const int64_t base = ((int64_t)b * S + lsi[l]) * M * D;                               \
                        for (int p = 0; p < P; ++p) {                                                         \
                            const int64_t ip = pw + l * P + p;                                                \
                            gloc[ip * 2] = 0; gloc[ip * 2 + 1] = 0; gw[ip] = 0;                               \
                            const T y = loc[ip * 2 + 1] * H - (T)0.5, x = loc[ip * 2] * W - (T)0.5;           \
                            const T aw = w[ip];                                                               \
                            if (!(y > -1 && x > -1 && y < H && x < W)) continue;                              \
                            const int y0 = (int)floor((double)y), x0 = (int)floor((double)x);                 \
                            const T fy = y - y0, fx = x - x0, gy = 1 - fy, gx = 1 - fx;                       \
                            const T cw[4] = {gy * gx, gy * fx, fy * gx, fy * fx};                             \
                            const int cy[4] = {y0, y0, y0 + 1, y0 + 1}, cx[4] = {x0, x0 + 1, x0, x0 + 1};     \
                            /* d(bilinear)/dy and /dx coefficients per corner */                              \
                            const T dy[4] = {-gx, -fx, gx, fx}, dx[4] = {-gy, gy, -fy, fy};                   \
                            T acc_w = 0, acc_x = 0, acc_y = 0;                                                \
                            for (int c = 0; c < D; ++c) {                                                     \
                                const T g = go[c], ga = g * aw;                                               \
                                T val = 0, ddy = 0, ddx = 0;                                                  \
                                for (int k = 0; k < 4; ++k) {                                                 \
                                    if (!(cy[k] >= 0 && cy[k] <= H - 1 && cx[k] >= 0 && cx[k] <= W - 1))      \
                                        continue;                                                             \
                                    const int64_t idx = base + ((int64_t)(cy[k] * W + cx[k]) * M + m) * D + c;\
                                    const T v = value[idx];                                                   \
                                    val += cw[k] * v; ddy += dy[k] * v; ddx += dx[k] * v;                     \
                                    gvalue[idx] += cw[k] * ga;                                                \
                                }                                                                             \
                                acc_w += g * val;                                                             \
                                acc_x += W * ddx * ga;                                                        \
                                acc_y += H * ddy * ga;                                                        \
                            }                                                                                 \
                            gw[ip] = acc_w; gloc[ip * 2] = acc_x; gloc[ip * 2 + 1] = acc_y;                   \
                        }                                                                                     \
                    }                                                                                         \
                }                                                                                             \
    }

DEFINE_MSDA(float, f32)
DEFINE_MSDA(double, f64)
