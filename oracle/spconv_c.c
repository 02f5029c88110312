/*
 * Plain-C restatement of the sparse-convolution semantics the reference relies on — TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py; nothing under ponderv2_b200/ links or loads this).
 *
 * A second, independently written restatement next to oracle/spconv_oracle.py (numpy/torch): both follow SURVEY.md
 * Appendix B, anchored on the reference's call sites
 *   SubMConv3d ............ ponder/models/sparse_unet/spconv_unet_v1m1_base.py:41-66,111-119
 *   SparseConv3d(k2,s2) ... :135-142            SparseInverseConv3d ... :171-177
 * and tests/test_oracle_cpu.py requires them to agree bit-exactly (integer rulebooks) / to 1e-12 (fp64 convolution).
 * PARITY UNPINNED against spconv itself (third-party, absent from /root/reference, not installable offline).
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC (ponderv2_b200/build.py: build_oracle_c).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int64_t key; int32_t row; } KeyRow;

static int cmp_keyrow(const void* a, const void* b) {
  const KeyRow* x = (const KeyRow*)a;
  const KeyRow* y = (const KeyRow*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->row < y->row ? -1 : (x->row > y->row ? 1 : 0);
}

static int64_t linear_key(int64_t b, int64_t c0, int64_t c1, int64_t c2, const int32_t* shape) {
  return ((b * shape[0] + c0) * shape[1] + c1) * shape[2] + c2;
}

/* smallest row with this key, or -1 (table sorted by (key, row)) */
static int32_t lookup(const KeyRow* t, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (t[mid].key < key) lo = mid + 1; else hi = mid;
  }
  return (lo < n && t[lo].key == key) ? t[lo].row : -1;
}

/* coords [n,4] int32 (batch, c0, c1, c2); nbr [ksize^3, n]: nbr[k][j] = smallest row i with
 * coord[i] == coord[j] + (k_a - ksize/2), k = (k0*ksize + k1)*ksize + k2, else -1.  Returns 0 / -1 (allocation). */
int oc_subm_rulebook(const int32_t* coords, int64_t n, const int32_t* shape, int ksize, int32_t* nbr) {
  const int K = ksize * ksize * ksize, R = ksize / 2;
  KeyRow* t = (KeyRow*)malloc(sizeof(KeyRow) * (size_t)(n > 0 ? n : 1));
  if (!t) return -1;
  for (int64_t i = 0; i < n; ++i) {
    t[i].key = linear_key(coords[4 * i], coords[4 * i + 1], coords[4 * i + 2], coords[4 * i + 3], shape);
    t[i].row = (int32_t)i;
  }
  qsort(t, (size_t)n, sizeof(KeyRow), cmp_keyrow);
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {
    const int32_t* c = coords + 4 * j;
    for (int k = 0; k < K; ++k) {
      const int k0 = k / (ksize * ksize), k1 = (k / ksize) % ksize, k2 = k % ksize;
      const int64_t a0 = c[1] + k0 - R, a1 = c[2] + k1 - R, a2 = c[3] + k2 - R;
      int32_t r = -1;
      if (a0 >= 0 && a0 < shape[0] && a1 >= 0 && a1 < shape[1] && a2 >= 0 && a2 < shape[2])
        r = lookup(t, n, linear_key(c[0], a0, a1, a2, shape));
      nbr[(int64_t)k * n + j] = r;
    }
  }
  free(t);
  return 0;
}

/* Kernel 2, stride 2, pad 0.  Output voxels = distinct (batch, c >> 1), numbered in ASCENDING linear key over the output
 * shape ((s - 2) / 2 + 1 per axis): the canonical order of Appendix B.  out_coords [n,4] (first *n_out rows valid),
 * in2out [n], koff [n] = ((c0&1)*2 + (c1&1))*2 + (c2&1). */
int oc_down_rulebook(const int32_t* coords, int64_t n, const int32_t* shape, int32_t* out_coords, int32_t* in2out,
                     int32_t* koff, int64_t* n_out) {
  int32_t oshape[3];
  for (int a = 0; a < 3; ++a) oshape[a] = (shape[a] - 2) / 2 + 1;
  KeyRow* t = (KeyRow*)malloc(sizeof(KeyRow) * (size_t)(n > 0 ? n : 1));
  if (!t) return -1;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* c = coords + 4 * i;
    /* a 2x2x2 window that falls outside the unpadded input produces no output (in2out = -1) */
    const int inside = (c[1] >> 1) < oshape[0] && (c[2] >> 1) < oshape[1] && (c[3] >> 1) < oshape[2];
    t[i].key = inside ? linear_key(c[0], c[1] >> 1, c[2] >> 1, c[3] >> 1, oshape) : INT64_MAX;
    t[i].row = (int32_t)i;
    koff[i] = ((c[1] & 1) * 2 + (c[2] & 1)) * 2 + (c[3] & 1);
  }
  qsort(t, (size_t)n, sizeof(KeyRow), cmp_keyrow);
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (t[i].key == INT64_MAX) { in2out[t[i].row] = -1; continue; }
    if (i == 0 || t[i].key != t[i - 1].key) {
      const int32_t* c = coords + 4 * (int64_t)t[i].row;
      out_coords[4 * m] = c[0]; out_coords[4 * m + 1] = c[1] >> 1; out_coords[4 * m + 2] = c[2] >> 1;
      out_coords[4 * m + 3] = c[3] >> 1;
      ++m;
    }
    in2out[t[i].row] = (int32_t)(m - 1);
  }
  *n_out = m;
  free(t);
  return 0;
}

/* y[j, co] = bias[co] + sum_k sum_ci w[co, k, ci] * x[nbr[k][j], ci]   (double precision; rows with nbr == -1 skipped) */
int oc_sparse_conv_f64(const double* x, const double* w, const double* bias, const int32_t* nbr, double* y, int64_t n_out,
                       int cin, int cout, int kvol) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n_out; ++j) {
    double* yr = y + j * cout;
    for (int co = 0; co < cout; ++co) yr[co] = bias ? bias[co] : 0.0;
    for (int k = 0; k < kvol; ++k) {
      const int32_t i = nbr[(int64_t)k * n_out + j];
      if (i < 0) continue;
      const double* xr = x + (int64_t)i * cin;
      for (int co = 0; co < cout; ++co) {
        const double* wr = w + ((int64_t)co * kvol + k) * cin;
        double acc = 0.0;
        for (int ci = 0; ci < cin; ++ci) acc += wr[ci] * xr[ci];
        yr[co] += acc;
      }
    }
  }
  return 0;
}
