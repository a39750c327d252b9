/*
 * ORACLE (test infrastructure, NOT the product): plain-C restatement of the reference's vendored
 * spconv v1 CPU path.
 *   rulebook   TransFusion/mmdet3d/ops/spconv/include/spconv/geometry.h:24-85 (getValidOutPos),
 *              :144-194 (getIndicePairsConv), :247-297 (getIndicePairsSubM),
 *              SubM stride/padding override spconv_ops.h:76-79
 *   conv       spconv_ops.h:260-361 (indiceConv: out[o] += W[k]^T in[i] over the pairs of offset k),
 *              :363-456 (indiceConvBackward)
 * The dense gridsOut array of the reference is replaced by a hash map with the same meaning.
 * Output voxels come out in the CPU path's FIRST-TOUCH order; oracle/spconv.py relabels them to
 * the GPU path's sorted-by-flat-index order (spconv_ops.h:129-137) when asked.
 * Pinned against the reference extension itself (oracle/_ref/sparse_conv_ext.so) by
 * tests/test_oracle_spconv.py; the reference's own tests hold no value vectors for this path.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ND 3

static int valid_out_pos(const int* in, const int* ks, const int* st, const int* pad, const int* dil,
                         const int* oshape, int* out /* [kvol][ND+1] */) {
  int lowers[ND], uppers[ND], counter[ND], csize[ND];
  int npts = 1, cnt = 0;
  for (int i = 0; i < ND; ++i) {
    lowers[i] = (in[i] - (ks[i] - 1) * dil[i] - 1 + st[i] + pad[i]) / st[i]; /* C truncation */
    uppers[i] = (in[i] + pad[i]) / st[i];
  }
  for (int i = 0; i < ND; ++i) {
    csize[i] = (uppers[i] - lowers[i]) / dil[i] + 1;
    npts *= csize[i];
    counter[i] = 0;
  }
  for (int i = 0; i < npts; ++i) {
    int valid = 1, m = 1, offset = 0;
    for (int j = ND - 1; j >= 0; --j) {
      const int val = uppers[j] - counter[j] * dil[j];
      out[cnt * (ND + 1) + j] = val;
      if (val < 0 || val > oshape[j] - 1) valid = 0;
      offset += m * (in[j] - val * st[j] + pad[j]) / dil[j];
      m *= ks[j];
    }
    out[cnt * (ND + 1) + ND] = offset;
    if (valid) ++cnt;
    counter[ND - 1] += 1;
    for (int c = ND - 1; c >= 0; --c) {
      if (counter[c] == csize[c] && c > 0) {
        counter[c - 1] += 1;
        counter[c] = 0;
      }
    }
  }
  return cnt;
}

typedef struct { int64_t* key; int32_t* val; uint64_t mask; } map_t;
static map_t map_new(int64_t n) {
  map_t m; uint64_t s = 1024;
  while (s < (uint64_t)(2 * n + 2)) s <<= 1;
  m.key = (int64_t*)malloc(s * sizeof(int64_t)); m.val = (int32_t*)malloc(s * sizeof(int32_t));
  for (uint64_t i = 0; i < s; ++i) m.key[i] = -1;
  m.mask = s - 1; return m;
}
static int32_t* map_slot(map_t* m, int64_t key, int* fresh) {
  uint64_t h = (((uint64_t)key * 0x9E3779B97F4A7C15ull) >> 17) & m->mask;
  while (m->key[h] != -1 && m->key[h] != key) h = (h + 1) & m->mask;
  *fresh = m->key[h] == -1;
  m->key[h] = key;
  return &m->val[h];
}
static void map_free(map_t* m) { free(m->key); free(m->val); }

static int64_t flat_idx(const int* p, const int* shape, int64_t b) {
  int64_t idx = b;
  for (int i = 0; i < ND; ++i) idx = idx * shape[i] + p[i];
  return idx;
}

/* indices [n,4] (b,z,y,x); indice_pairs [K,2,n] pre-filled with -1; indice_num [K] zeroed;
 * out_indices [n*K,4]. Returns numActOut (== n for subm). */
int64_t oracle_get_indice_pairs(const int32_t* indices, int64_t n, const int32_t* out_shape,
                                const int32_t* ksize, const int32_t* stride_, const int32_t* padding_,
                                const int32_t* dilation, int subm, int32_t* out_indices,
                                int32_t* indice_pairs, int32_t* indice_num) {
  int st[ND], pad[ND], kvol = 1;
  for (int i = 0; i < ND; ++i) {
    st[i] = subm ? 1 : stride_[i];
    pad[i] = subm ? ksize[i] / 2 : padding_[i];
    kvol *= ksize[i];
  }
  int* vp = (int*)malloc(sizeof(int) * kvol * (ND + 1));
  map_t grid = map_new(subm ? n : n * kvol);
  int64_t num_act = 0;
  int fresh;
  if (subm) {
    for (int64_t j = 0; j < n; ++j)
      *map_slot(&grid, flat_idx(indices + j * 4 + 1, out_shape, indices[j * 4]), &fresh) = (int32_t)j;
    for (int64_t j = 0; j < n; ++j) {
      const int nv = valid_out_pos(indices + j * 4 + 1, ksize, st, pad, dilation, out_shape, vp);
      for (int i = 0; i < nv; ++i) {
        const int* p = vp + i * (ND + 1);
        const int off = p[ND];
        const int64_t key = flat_idx(p, out_shape, indices[j * 4]);
        /* lookup without inserting */
        uint64_t h = (((uint64_t)key * 0x9E3779B97F4A7C15ull) >> 17) & grid.mask;
        while (grid.key[h] != -1 && grid.key[h] != key) h = (h + 1) & grid.mask;
        if (grid.key[h] == key) {
          indice_pairs[((int64_t)off * 2) * n + indice_num[off]] = (int32_t)j;
          indice_pairs[((int64_t)off * 2 + 1) * n + indice_num[off]++] = grid.val[h];
        }
      }
    }
    num_act = n;
  } else {
    for (int64_t j = 0; j < n; ++j) {
      const int32_t b = indices[j * 4];
      const int nv = valid_out_pos(indices + j * 4 + 1, ksize, st, pad, dilation, out_shape, vp);
      for (int i = 0; i < nv; ++i) {
        const int* p = vp + i * (ND + 1);
        const int off = p[ND];
        int32_t* slot = map_slot(&grid, flat_idx(p, out_shape, b), &fresh);
        if (fresh) {
          out_indices[num_act * 4] = b;
          for (int k = 0; k < ND; ++k) out_indices[num_act * 4 + 1 + k] = p[k];
          *slot = (int32_t)num_act++;
        }
        indice_pairs[((int64_t)off * 2) * n + indice_num[off]] = (int32_t)j;
        indice_pairs[((int64_t)off * 2 + 1) * n + indice_num[off]++] = *slot;
      }
    }
  }
  map_free(&grid);
  free(vp);
  return num_act;
}

/* out [n_out, cout] zeroed inside. filters [K, cin, cout]. inverse swaps the pair roles. */
void oracle_indice_conv_f32(const float* feat, const float* filters, const int32_t* pairs,
                            const int32_t* num, int64_t n_pairs_stride, int64_t kvol, int64_t cin,
                            int64_t cout, int64_t n_out, int inverse, float* out) {
  memset(out, 0, sizeof(float) * (size_t)(n_out * cout));
  for (int64_t k = 0; k < kvol; ++k) {
    const int32_t* pin = pairs + (k * 2 + (inverse ? 1 : 0)) * n_pairs_stride;
    const int32_t* pout = pairs + (k * 2 + (inverse ? 0 : 1)) * n_pairs_stride;
    const float* w = filters + k * cin * cout;
    for (int64_t s = 0; s < num[k]; ++s) {
      const float* x = feat + (int64_t)pin[s] * cin;
      float* y = out + (int64_t)pout[s] * cout;
      for (int64_t ci = 0; ci < cin; ++ci) {
        const float xv = x[ci];
        const float* wr = w + ci * cout;
        for (int64_t co = 0; co < cout; ++co) y[co] += xv * wr[co];
      }
    }
  }
}

/* dgrad [n_in, cin], wgrad [K, cin, cout], both zeroed inside. */
void oracle_indice_conv_backward_f32(const float* feat, const float* filters, const float* gout,
                                     const int32_t* pairs, const int32_t* num, int64_t n_pairs_stride,
                                     int64_t kvol, int64_t cin, int64_t cout, int64_t n_in, int inverse,
                                     float* gin, float* gw) {
  memset(gin, 0, sizeof(float) * (size_t)(n_in * cin));
  memset(gw, 0, sizeof(float) * (size_t)(kvol * cin * cout));
  for (int64_t k = 0; k < kvol; ++k) {
    const int32_t* pin = pairs + (k * 2 + (inverse ? 1 : 0)) * n_pairs_stride;
    const int32_t* pout = pairs + (k * 2 + (inverse ? 0 : 1)) * n_pairs_stride;
    const float* w = filters + k * cin * cout;
    float* dw = gw + k * cin * cout;
    for (int64_t s = 0; s < num[k]; ++s) {
      const float* x = feat + (int64_t)pin[s] * cin;
      const float* g = gout + (int64_t)pout[s] * cout;
      float* dx = gin + (int64_t)pin[s] * cin;
      for (int64_t ci = 0; ci < cin; ++ci) {
        const float xv = x[ci];
        float acc = 0.f;
        for (int64_t co = 0; co < cout; ++co) {
          dw[ci * cout + co] += xv * g[co];
          acc += w[ci * cout + co] * g[co];
        }
        dx[ci] += acc;
      }
    }
  }
}
