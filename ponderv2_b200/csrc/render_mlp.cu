// SDF decoder of the outdoor renderer configuration as two fused kernels (outdoor):
//
//   x_0 = pf (Wp p + bp);  z_l = x_l + Fc_l f + bc_l;  y_l = W_l z_l + b_l;  x_{l+1} = softplus_100(y_l);  sdf = y_{L-1}[0]
//
// (`SDFDecoder.forward`, ponder/models/ponder/render_utils/decoders.py:6-36 with hidden_size 16, in_dim 32, n_blocks 5:
// configs/nuscenes/pretrain-ponder-spunet-v1m1-0-base.py:36-41) together with the two vector-Jacobian products the
// renderer needs from it, u = d sdf / d f and v = d sdf / d p (direct path through fc_p), so that
//   d sdf / d points = (d f / d points)^T u + v
// is completed by the trilinear sampler's own backward (fields/sdf_field.py:226-238 does the same through
// autograd.grad(create_graph=True) over ~40 small torch kernels per pass).  The reference differentiates THROUGH that
// gradient (eikonal / alpha terms), so the backward kernel is the hand-derived reverse of (sdf, u, v): it recomputes
// the forward per point and emits, per layer, the four adjoint vectors whose point-contractions are the parameter
// gradients (A_l = dL/dy_l, C_l = ybar_l, Z_l = z_l, D_l = adjoint of zbar_l), plus dL/df.  The contractions over the
// points ([P,16]^T [P,16|32]) run on the tensor-core weight-gradient kernel (host side: render/mlp.py).
//
// One thread per sample point, weights (4.9 k floats for L = 6) in shared memory and read as broadcast float4; per point
// ~9 k FMA forward, ~25 k backward: compute-light, bound by the 128-byte feature row in and the per-layer vectors out.
#include "pv2_common.cuh"

namespace {

constexpr int kH = 16, kF = 32, kMaxL = 8, kMlpThreads = 128;

struct MlpLayout {
  int wp, bp, fc, bc, w, b, total;   // float offsets into the packed parameter vector
};
__host__ __device__ inline MlpLayout mlp_layout(int L, int O) {
  MlpLayout m;
  m.wp = 0; m.bp = kH * 3; m.fc = m.bp + kH; m.bc = m.fc + L * kH * kF; m.w = m.bc + L * kH;
  m.b = m.w + (L - 1) * kH * kH + O * kH; m.total = m.b + (L - 1) * kH + O;
  return m;
}

__device__ __forceinline__ float softplus100(float y) {
  const float t = 100.f * y;
  return t > 20.f ? y : log1pf(expf(t)) * 0.01f;      // torch.nn.Softplus(beta=100, threshold=20)
}
__device__ __forceinline__ float sigmoid100(float y) { return 1.f / (1.f + expf(-100.f * y)); }

// out[i] = sum_j M[i][j] v[j], M row-major [kH][n] in shared memory (n = 16 or 32), broadcast float4 reads
template <int N>
__device__ __forceinline__ void matvec(const float* __restrict__ M, const float (&v)[N], float (&out)[kH]) {
#pragma unroll
  for (int i = 0; i < kH; ++i) {
    float acc = 0.f;
    const float4* row = reinterpret_cast<const float4*>(M + i * N);
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 w = row[q];
      acc = fmaf(w.x, v[4 * q], acc); acc = fmaf(w.y, v[4 * q + 1], acc);
      acc = fmaf(w.z, v[4 * q + 2], acc); acc = fmaf(w.w, v[4 * q + 3], acc);
    }
    out[i] = acc;
  }
}
// out[j] (+)= sum_i M[i][j] v[i]  (transposed product), M [kH][N]
template <int N, bool kAcc>
__device__ __forceinline__ void matvec_t(const float* __restrict__ M, const float (&v)[kH], float (&out)[N]) {
  if (!kAcc) {
#pragma unroll
    for (int j = 0; j < N; ++j) out[j] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < kH; ++i) {
    const float4* row = reinterpret_cast<const float4*>(M + i * N);
    const float vi = v[i];
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 w = row[q];
      out[4 * q] = fmaf(w.x, vi, out[4 * q]); out[4 * q + 1] = fmaf(w.y, vi, out[4 * q + 1]);
      out[4 * q + 2] = fmaf(w.z, vi, out[4 * q + 2]); out[4 * q + 3] = fmaf(w.w, vi, out[4 * q + 3]);
    }
  }
}

__device__ __forceinline__ void load_row32(const float* __restrict__ p, float (&v)[kF]) {
#pragma unroll
  for (int q = 0; q < kF / 4; ++q) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
template <int N>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&v)[N]) {
#pragma unroll
  for (int q = 0; q < N / 4; ++q)
    reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// forward: sdf [P]; u [P,32], v [P,3] optional (nullptr: the sampler's no-grad coarse pass only needs sdf)
__global__ void __launch_bounds__(kMlpThreads) sdf_mlp_fwd_kernel(const float* __restrict__ f, const float* __restrict__ pts,
                                                                  const float* __restrict__ prm, int L, int O, float pf,
                                                                  int64_t P, float* __restrict__ sdf, float* __restrict__ u,
                                                                  float* __restrict__ v) {
  extern __shared__ __align__(16) float sp[];
  const MlpLayout m = mlp_layout(L, O);
  for (int i = threadIdx.x; i < m.total; i += kMlpThreads) sp[i] = prm[i];
  __syncthreads();
  for (int64_t pt = (int64_t)blockIdx.x * kMlpThreads + threadIdx.x; pt < P; pt += (int64_t)gridDim.x * kMlpThreads) {
    float fv[kF];
    load_row32(f + pt * kF, fv);
    const float p0 = pts[pt * 3], p1 = pts[pt * 3 + 1], p2 = pts[pt * 3 + 2];
    float x[kH], s[kMaxL - 1][kH];
#pragma unroll
    for (int i = 0; i < kH; ++i)
      x[i] = pf * (sp[m.wp + i * 3] * p0 + sp[m.wp + i * 3 + 1] * p1 + sp[m.wp + i * 3 + 2] * p2 + sp[m.bp + i]);
    float y0 = 0.f;
    for (int l = 0; l < L; ++l) {
      float z[kH], t[kH];
      matvec<kF>(sp + m.fc + l * kH * kF, fv, t);
#pragma unroll
      for (int i = 0; i < kH; ++i) z[i] = x[i] + t[i] + sp[m.bc + l * kH + i];
      if (l < L - 1) {
        matvec<kH>(sp + m.w + l * kH * kH, z, t);
#pragma unroll
        for (int i = 0; i < kH; ++i) {
          const float y = t[i] + sp[m.b + l * kH + i];
          s[l][i] = sigmoid100(y);
          x[i] = softplus100(y);
        }
      } else {
        const float* wl = sp + m.w + (L - 1) * kH * kH;     // row 0 of the last layer = the sdf output
        y0 = sp[m.b + (L - 1) * kH];
#pragma unroll
        for (int j = 0; j < kH; ++j) y0 = fmaf(wl[j], z[j], y0);
      }
    }
    sdf[pt] = y0;
    if (u == nullptr) continue;
    // reverse sweep for u = d sdf / d f and v = d sdf / d p: zbar_{L-1} = W_{L-1}[0], ybar_l = zbar_{l+1} * s_l,
    // zbar_l = W_l^T ybar_l, u = sum_l Fc_l^T zbar_l, v = pf Wp^T zbar_0
    float zbar[kH], uv[kF];
#pragma unroll
    for (int j = 0; j < kH; ++j) zbar[j] = sp[m.w + (L - 1) * kH * kH + j];
    matvec_t<kF, false>(sp + m.fc + (L - 1) * kH * kF, zbar, uv);
    for (int l = L - 2; l >= 0; --l) {
      float yb[kH];
#pragma unroll
      for (int i = 0; i < kH; ++i) yb[i] = zbar[i] * s[l][i];
      matvec_t<kH, false>(sp + m.w + l * kH * kH, yb, zbar);
      matvec_t<kF, true>(sp + m.fc + l * kH * kF, zbar, uv);
    }
    store_row<kF>(u + pt * kF, uv);
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int j = 0; j < kH; ++j) {
      v0 = fmaf(sp[m.wp + j * 3], zbar[j], v0); v1 = fmaf(sp[m.wp + j * 3 + 1], zbar[j], v1);
      v2 = fmaf(sp[m.wp + j * 3 + 2], zbar[j], v2);
    }
    v[pt * 3] = pf * v0; v[pt * 3 + 1] = pf * v1; v[pt * 3 + 2] = pf * v2;
  }
}

// backward of (sdf, u, v) wrt f and the parameters, per point (derivation checked against autograd's double backward in
// fp64, tests/test_host_cpu.py::test_sdf_mlp_adjoint_math).  Outputs per point:
//   fbar [P,32]                      dL/df
//   A [L-1][P][16] = dL/dy_l         C [L-1][P][16] = ybar_l          (l < L-1; the last layer's are rank one: g_sdf e_0, e_0)
//   Z [L][P][16]   = z_l             D [L][P][16]   = adjoint of zbar_l
__global__ void __launch_bounds__(kMlpThreads) sdf_mlp_bwd_kernel(const float* __restrict__ f, const float* __restrict__ pts,
                                                                  const float* __restrict__ prm, int L, int O, float pf,
                                                                  int64_t P, const float* __restrict__ g_sdf,
                                                                  const float* __restrict__ g_u, const float* __restrict__ g_v,
                                                                  float* __restrict__ fbar, float* __restrict__ A,
                                                                  float* __restrict__ Cc, float* __restrict__ Z,
                                                                  float* __restrict__ D) {
  extern __shared__ __align__(16) float sp[];
  const MlpLayout m = mlp_layout(L, O);
  for (int i = threadIdx.x; i < m.total; i += kMlpThreads) sp[i] = prm[i];
  __syncthreads();
  for (int64_t pt = (int64_t)blockIdx.x * kMlpThreads + threadIdx.x; pt < P; pt += (int64_t)gridDim.x * kMlpThreads) {
    float fv[kF], gu[kF];
    load_row32(f + pt * kF, fv);
    const float p0 = pts[pt * 3], p1 = pts[pt * 3 + 1], p2 = pts[pt * 3 + 2];
    const float gs = g_sdf != nullptr ? g_sdf[pt] : 0.f;
    if (g_u != nullptr) load_row32(g_u + pt * kF, gu);
    else {
#pragma unroll
      for (int c = 0; c < kF; ++c) gu[c] = 0.f;
    }
    const float gv0 = g_v ? g_v[pt * 3] : 0.f, gv1 = g_v ? g_v[pt * 3 + 1] : 0.f, gv2 = g_v ? g_v[pt * 3 + 2] : 0.f;
    // ---- forward recompute: z_l (stored to Z), s_l
    float x[kH], s[kMaxL - 1][kH];
#pragma unroll
    for (int i = 0; i < kH; ++i)
      x[i] = pf * (sp[m.wp + i * 3] * p0 + sp[m.wp + i * 3 + 1] * p1 + sp[m.wp + i * 3 + 2] * p2 + sp[m.bp + i]);
    for (int l = 0; l < L; ++l) {
      float z[kH], t[kH];
      matvec<kF>(sp + m.fc + l * kH * kF, fv, t);
#pragma unroll
      for (int i = 0; i < kH; ++i) z[i] = x[i] + t[i] + sp[m.bc + l * kH + i];
      store_row<kH>(Z + ((int64_t)l * P + pt) * kH, z);
      if (l < L - 1) {
        matvec<kH>(sp + m.w + l * kH * kH, z, t);
#pragma unroll
        for (int i = 0; i < kH; ++i) {
          const float y = t[i] + sp[m.b + l * kH + i];
          s[l][i] = sigmoid100(y);
          x[i] = softplus100(y);
        }
      }
    }
    // ---- the u-chain (reverse sweep of the forward kernel), keeping zbar_{l+1} for l < L-1 and storing ybar_l
    float zb[kMaxL][kH];
#pragma unroll
    for (int j = 0; j < kH; ++j) zb[L - 1][j] = sp[m.w + (L - 1) * kH * kH + j];
    for (int l = L - 2; l >= 0; --l) {
      float yb[kH];
#pragma unroll
      for (int i = 0; i < kH; ++i) yb[i] = zb[l + 1][i] * s[l][i];
      store_row<kH>(Cc + ((int64_t)l * P + pt) * kH, yb);
      matvec_t<kH, false>(sp + m.w + l * kH * kH, yb, zb[l]);
    }
    // ---- adjoint of the u-chain, forward in l:  D_0 = Fc_0 gu + pf Wp gv;  a_l = W_l D_l;  sbar_l = a_l * zbar_{l+1};
    //      D_{l+1} = Fc_{l+1} gu + a_l * s_l
    float sbar[kMaxL - 1][kH];
    {
      float d[kH], a[kH];
      matvec<kF>(sp + m.fc, gu, d);
#pragma unroll
      for (int j = 0; j < kH; ++j)
        d[j] += pf * (sp[m.wp + j * 3] * gv0 + sp[m.wp + j * 3 + 1] * gv1 + sp[m.wp + j * 3 + 2] * gv2);
      store_row<kH>(D + pt * kH, d);
      for (int l = 0; l < L - 1; ++l) {
        matvec<kH>(sp + m.w + l * kH * kH, d, a);
        float t[kH];
        matvec<kF>(sp + m.fc + (l + 1) * kH * kF, gu, t);
#pragma unroll
        for (int i = 0; i < kH; ++i) {
          sbar[l][i] = a[i] * zb[l + 1][i];
          d[i] = t[i] + a[i] * s[l][i];
        }
        store_row<kH>(D + ((int64_t)(l + 1) * P + pt) * kH, d);
      }
    }
    // ---- ordinary reverse sweep: zhat_{L-1} = g_sdf W_{L-1}[0];  yhat_l = zhat_{l+1} * s_l + sbar_l 100 s_l (1 - s_l);
    //      zhat_l = W_l^T yhat_l;  fbar = sum_l Fc_l^T zhat_l
    float zh[kH], fb[kF];
#pragma unroll
    for (int j = 0; j < kH; ++j) zh[j] = gs * sp[m.w + (L - 1) * kH * kH + j];
    matvec_t<kF, false>(sp + m.fc + (L - 1) * kH * kF, zh, fb);
    for (int l = L - 2; l >= 0; --l) {
      float yh[kH];
#pragma unroll
      for (int i = 0; i < kH; ++i) yh[i] = zh[i] * s[l][i] + sbar[l][i] * 100.f * s[l][i] * (1.f - s[l][i]);
      store_row<kH>(A + ((int64_t)l * P + pt) * kH, yh);
      matvec_t<kH, false>(sp + m.w + l * kH * kH, yh, zh);
      matvec_t<kF, true>(sp + m.fc + l * kH * kF, zh, fb);
    }
    store_row<kF>(fbar + pt * kF, fb);
  }
}

}  // namespace

extern "C" {

/* number of floats of the packed parameter vector [Wp(16x3) | bp(16) | Fc_l(16x32) x L | bc_l(16) x L | W_l(16x16) x (L-1),
 * W_last(O x 16) | b_l(16) x (L-1), b_last(O)] */
int64_t pv2_sdf_mlp_param_count(int L, int O) {
  if (L < 1 || L > kMaxL || O < 1) return 0;
  return mlp_layout(L, O).total;
}

int pv2_sdf_mlp_fwd(const float* f, const float* pts, const float* params, int L, int F, int H, int O, float points_factor,
                    int64_t P, float* sdf, float* u, float* v, void* stream_) {
  PV2_CHECK_ARG(P >= 0 && L >= 1 && O >= 1);
  if (F != kF || H != kH || L > kMaxL) return PV2_EUNSUPPORTED;
  if (P == 0) return 0;
  PV2_CHECK_ARG(f && pts && params && sdf && ((u == nullptr) == (v == nullptr)));
  PV2_CHECK_ARG((((uintptr_t)f | (uintptr_t)u) & 15) == 0);
  const size_t smem = (size_t)mlp_layout(L, O).total * sizeof(float);
  sdf_mlp_fwd_kernel<<<pv2_grid_for(P, kMlpThreads, 4), kMlpThreads, smem, (cudaStream_t)stream_>>>(f, pts, params, L, O,
                                                                                                     points_factor, P, sdf, u, v);
  PV2_DONE(1);
}

int pv2_sdf_mlp_bwd(const float* f, const float* pts, const float* params, int L, int F, int H, int O, float points_factor,
                    int64_t P, const float* g_sdf, const float* g_u, const float* g_v, float* fbar, float* A, float* C,
                    float* Z, float* D, void* stream_) {
  PV2_CHECK_ARG(P >= 0 && L >= 1 && O >= 1);
  if (F != kF || H != kH || L > kMaxL) return PV2_EUNSUPPORTED;
  if (P == 0) return 0;
  PV2_CHECK_ARG(f && pts && params && fbar && Z && D && (L == 1 || (A && C)));
  PV2_CHECK_ARG((((uintptr_t)f | (uintptr_t)g_u | (uintptr_t)fbar | (uintptr_t)A | (uintptr_t)C | (uintptr_t)Z | (uintptr_t)D) & 15) == 0);
  const size_t smem = (size_t)mlp_layout(L, O).total * sizeof(float);
  sdf_mlp_bwd_kernel<<<pv2_grid_for(P, kMlpThreads, 4), kMlpThreads, smem, (cudaStream_t)stream_>>>(
      f, pts, params, L, O, points_factor, P, g_sdf, g_u, g_v, fbar, A, C, Z, D);
  PV2_DONE(1);
}

}  // extern "C"
