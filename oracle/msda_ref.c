/*
 * ORACLE (test infrastructure, NOT the product): plain-C CPU restatement of the reference's
 * multi-scale deformable attention kernels.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call this.
 *
 * Follows, per (batch b, query q, head m, channel c):
 *   forward   <proj>/models/model_utils/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299
 *             bilinear sample :33-84
 *   backward  :87-159 (per-corner grads), :301-403 (reduction of grad_loc / grad_attn over
 *             channels; the reference sums the per-channel partials serially in channel order,
 *             :376-393, which is what the inner c loop here does)
 * The reference has no CPU implementation (ops/src/cpu/ms_deform_attn_cpu.cpp:26,39 -> AT_ERROR);
 * this file is pinned against the reference's own pure-PyTorch ms_deform_attn_core_pytorch
 * (ops/functions/ms_deform_attn_func.py:41-61) through tests/golden/msda_*.npz.
 *
 * Compiled with -ffp-contract=off so fp32 products are rounded exactly as written.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_MSDA(T, SUF)                                                                      \
  void oracle_msda_forward_##SUF(const T* value, const int64_t* shapes, const int64_t* lsi,      \
                                 const T* loc, const T* attn, T* out, int64_t N, int64_t S,      \
                                 int64_t M, int64_t D, int64_t L, int64_t Lq, int64_t P) {       \
    const int64_t nqm = N * Lq * M;                                                              \
    _Pragma("omp parallel for schedule(static)") for (int64_t qm = 0; qm < nqm; ++qm) {          \
      const int64_t m = qm % M, b = (qm / M) / Lq;                                               \
      const int64_t pix = M * D;                                                                 \
      for (int64_t c = 0; c < D; ++c) {                                                          \
        T col = 0;                                                                               \
        for (int64_t l = 0; l < L; ++l) {                                                        \
          const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                          \
          const T* vl = value + (b * S + lsi[l]) * pix + m * D + c;                              \
          for (int64_t p = 0; p < P; ++p) {                                                      \
            const int64_t pi = (qm * L + l) * P + p;                                             \
            const T lw_ = loc[2 * pi], lh_ = loc[2 * pi + 1], w = attn[pi];                      \
            const T h_im = lh_ * (T)H - (T)0.5, w_im = lw_ * (T)W - (T)0.5;                      \
            if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;                     \
            const int h_low = (int)floor(h_im), w_low = (int)floor(w_im);                        \
            const int h_high = h_low + 1, w_high = w_low + 1;                                    \
            const T lh = h_im - h_low, lw = w_im - w_low, hh = 1 - lh, hw = 1 - lw;              \
            T v1 = 0, v2 = 0, v3 = 0, v4 = 0;                                                    \
            if (h_low >= 0 && w_low >= 0) v1 = vl[((int64_t)h_low * W + w_low) * pix];           \
            if (h_low >= 0 && w_high <= W - 1) v2 = vl[((int64_t)h_low * W + w_high) * pix];     \
            if (h_high <= H - 1 && w_low >= 0) v3 = vl[((int64_t)h_high * W + w_low) * pix];     \
            if (h_high <= H - 1 && w_high <= W - 1) v4 = vl[((int64_t)h_high * W + w_high) * pix]; \
            const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                      \
            col += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * w;                                  \
          }                                                                                      \
        }                                                                                        \
        out[qm * D + c] = col;                                                                   \
      }                                                                                          \
    }                                                                                            \
  }                                                                                              \
                                                                                                 \
  /* grad_value accumulates with += in (b,q,m,c,l,p) order; callers get zeroed outputs. */       \
  void oracle_msda_backward_##SUF(const T* value, const int64_t* shapes, const int64_t* lsi,     \
                                  const T* loc, const T* attn, const T* gout, T* gvalue,         \
                                  T* gloc, T* gattn, int64_t N, int64_t S, int64_t M, int64_t D, \
                                  int64_t L, int64_t Lq, int64_t P) {                            \
    memset(gvalue, 0, sizeof(T) * (size_t)(N * S * M * D));                                      \
    memset(gloc, 0, sizeof(T) * (size_t)(N * Lq * M * L * P * 2));                               \
    memset(gattn, 0, sizeof(T) * (size_t)(N * Lq * M * L * P));                                  \
    const int64_t pix = M * D;                                                                   \
    /* parallel over (b, m): each owns a disjoint slice of grad_value -> no atomics */           \
    _Pragma("omp parallel for schedule(dynamic, 1)") for (int64_t bm = 0; bm < N * M; ++bm) {    \
      const int64_t b = bm / M, m = bm % M;                                                      \
      for (int64_t q = 0; q < Lq; ++q) {                                                         \
        const int64_t qm = (b * Lq + q) * M + m;                                                 \
        for (int64_t l = 0; l < L; ++l) {                                                        \
          const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                          \
          const int64_t lo = (b * S + lsi[l]) * pix + m * D;                                     \
          for (int64_t p = 0; p < P; ++p) {                                                      \
            const int64_t pi = (qm * L + l) * P + p;                                             \
            const T lw_ = loc[2 * pi], lh_ = loc[2 * pi + 1], a = attn[pi];                      \
            const T h_im = lh_ * (T)H - (T)0.5, w_im = lw_ * (T)W - (T)0.5;                      \
            if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;                     \
            const int h_low = (int)floor(h_im), w_low = (int)floor(w_im);                        \
            const int h_high = h_low + 1, w_high = w_low + 1;                                    \
            const T lh = h_im - h_low, lw = w_im - w_low, hh = 1 - lh, hw = 1 - lw;              \
            const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                      \
            const int ok1 = h_low >= 0 && w_low >= 0, ok2 = h_low >= 0 && w_high <= W - 1,       \
                      ok3 = h_high <= H - 1 && w_low >= 0,                                       \
                      ok4 = h_high <= H - 1 && w_high <= W - 1;                                  \
            const int64_t o1 = lo + ((int64_t)h_low * W + w_low) * pix, o2 = o1 + pix,           \
                          o3 = o1 + (int64_t)W * pix, o4 = o3 + pix;                             \
            T ga = 0, gw = 0, gh = 0;                                                            \
            for (int64_t c = 0; c < D; ++c) {                                                    \
              const T g = gout[qm * D + c], tg = g * a;                                          \
              T v1 = 0, v2 = 0, v3 = 0, v4 = 0, ghw = 0, gww = 0;                                \
              if (ok1) { v1 = value[o1 + c]; ghw -= hw * v1; gww -= hh * v1; gvalue[o1 + c] += w1 * tg; } \
              if (ok2) { v2 = value[o2 + c]; ghw -= lw * v2; gww += hh * v2; gvalue[o2 + c] += w2 * tg; } \
              if (ok3) { v3 = value[o3 + c]; ghw += hw * v3; gww -= lh * v3; gvalue[o3 + c] += w3 * tg; } \
              if (ok4) { v4 = value[o4 + c]; ghw += lw * v4; gww += lh * v4; gvalue[o4 + c] += w4 * tg; } \
              ga += g * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);                                 \
              gw += (T)W * gww * tg;                                                             \
              gh += (T)H * ghw * tg;                                                             \
            }                                                                                    \
            gattn[pi] = ga;                                                                      \
            gloc[2 * pi] = gw;                                                                   \
            gloc[2 * pi + 1] = gh;                                                               \
          }                                                                                      \
        }                                                                                        \
      }                                                                                          \
    }                                                                                            \
  }

DEFINE_MSDA(float, f32)
DEFINE_MSDA(double, f64)
