// Trilinear volume sampler with analytic first and second derivatives (boundary B2).
//
// Same contract as the reference extension smooth_sampler._C (libs/smooth-sampler/smooth_sampler/csrc/
// smooth_sampler.cpp:36-97, kernels smooth_sampler_kernel.cu:39-153 fwd, :155-356 bwd, :358-619 bwd-bwd),
// whose coordinate handling is torch's grid_sample (unnormalise, border clip / reflection, zeros padding).
// input (N,C,D,H,W) contiguous, grid (N,P,3) in [-1,1] (x->W, y->H, z->D), output (N,C,P).
//
// One thread per sample point, channels walked in the inner loop: the 8 corner addresses and the 8 (or
// 8x12) weights are computed once per point.  The renderer's own hot path uses the channels-last fused
// kernels in render_*.cu; this file is the general-purpose, twice-differentiable operator.
#include "pv2_common.cuh"

namespace {

template <typename S>
struct AxisT {
  int i0;      // floor index; corner 1 is i0 + 1
  S w[2];      // interpolation weights of corner 0 / 1
  S d[2];      // d w / d grid-coordinate (includes the unnormalise / clip / reflect multiplier)
  S dd[2];     // d2 w / d grid-coordinate^2 (non-zero only with smoothstep)
};

template <typename S>
__device__ __forceinline__ S clip_coord(S x, int size, S* g) {
  if (x <= (S)0) { *g = (S)0; return (S)0; }
  S mx = (S)(size - 1);
  if (x >= mx) { *g = (S)0; return mx; }
  *g = (S)1;
  return x;
}

template <typename S>
__device__ __forceinline__ S reflect_coord(S x, int twice_low, int twice_high, S* g) {
  if (twice_low == twice_high) { *g = (S)0; return (S)0; }
  S mn = (S)twice_low / (S)2;
  S span = (S)(twice_high - twice_low) / (S)2;
  x = x - mn;
  S sign = (S)1;
  if (x < (S)0) { sign = (S)-1; x = -x; }
  S extra = fmod(x, span);
  int flips = (int)floor(x / span);
  if ((flips & 1) == 0) { *g = sign; return extra + mn; }
  *g = -sign;
  return span - extra + mn;
}

template <typename S>
__device__ __forceinline__ AxisT<S> make_axis(S g, int size, int pad, bool align, bool smooth) {
  AxisT<S> a;
  S mult, x;
  if (align) { x = ((g + (S)1) / (S)2) * (S)(size - 1); mult = (S)(size - 1) / (S)2; }
  else       { x = ((g + (S)1) * (S)size - (S)1) / (S)2; mult = (S)size / (S)2; }
  if (pad == 1) {
    S gc; x = clip_coord<S>(x, size, &gc); mult *= gc;
  } else if (pad == 2) {
    S gr, gc;
    x = align ? reflect_coord<S>(x, 0, 2 * (size - 1), &gr) : reflect_coord<S>(x, -1, 2 * size - 1, &gr);
    x = clip_coord<S>(x, size, &gc);
    mult *= gr * gc;
  }
  // torch's safe_downgrade_to_int_range: non-finite / huge coordinates land far outside the volume
  if (!(x <= (S)2147483646.0 && x >= (S)-2147483648.0)) x = (S)-100;
  S fl = floor(x);
  a.i0 = (int)fl;
  S t = x - fl;
  S d1 = mult, dd1 = (S)0;
  if (smooth) {
    d1 = mult * ((S)6 * t * ((S)1 - t));
    dd1 = mult * mult * ((S)6 - (S)12 * t);
    t = t * t * ((S)3 - (S)2 * t);
  }
  a.w[1] = t;   a.w[0] = (S)1 - t;
  a.d[1] = d1;  a.d[0] = -d1;
  a.dd[1] = dd1; a.dd[0] = -dd1;
  return a;
}

template <typename S>
struct PointCtx {
  AxisT<S> ax, ay, az;
  int64_t off[8];   // element offset of each corner inside one channel, -1 when out of bounds
};

template <typename S>
__device__ __forceinline__ PointCtx<S> make_point(const S* __restrict__ grid, int64_t pt, int D, int H, int W, int pad,
                                                  bool align, bool smooth) {
  PointCtx<S> p;
  S gx = grid[pt * 3 + 0], gy = grid[pt * 3 + 1], gz = grid[pt * 3 + 2];
  p.ax = make_axis<S>(gx, W, pad, align, smooth);
  p.ay = make_axis<S>(gy, H, pad, align, smooth);
  p.az = make_axis<S>(gz, D, pad, align, smooth);
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    int ix = p.ax.i0 + (s & 1), iy = p.ay.i0 + ((s >> 1) & 1), iz = p.az.i0 + ((s >> 2) & 1);
    bool in = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H) & (iz >= 0) & (iz < D);
    p.off[s] = in ? ((int64_t)iz * H + iy) * W + ix : (int64_t)-1;
  }
  return p;
}

template <typename S>
__global__ void __launch_bounds__(256) trilinear_fwd_kernel(const S* __restrict__ input, const S* __restrict__ grid,
                                                            S* __restrict__ output, int64_t N, int64_t C, int D, int H,
                                                            int W, int64_t P, int pad, bool align, bool smooth) {
  const int64_t total = N * P;
  const int64_t vol = (int64_t)D * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / P, pp = idx - n * P;
    PointCtx<S> p = make_point<S>(grid, idx, D, H, W, pad, align, smooth);
    S wgt[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) wgt[s] = p.ax.w[s & 1] * p.ay.w[(s >> 1) & 1] * p.az.w[(s >> 2) & 1];
    const S* in_c = input + n * C * vol;
    S* out_c = output + n * C * P + pp;
    for (int64_t c = 0; c < C; ++c, in_c += vol, out_c += P) {
      S acc = (S)0;
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (p.off[s] >= 0) acc += __ldg(in_c + p.off[s]) * wgt[s];
      *out_c = acc;
    }
  }
}

template <typename S>
__global__ void __launch_bounds__(256) trilinear_bwd_kernel(const S* __restrict__ grad_output, const S* __restrict__ input,
                                                            const S* __restrict__ grid, S* __restrict__ grad_input,
                                                            S* __restrict__ grad_grid, int64_t N, int64_t C, int D, int H,
                                                            int W, int64_t P, int pad, bool align, bool smooth) {
  const int64_t total = N * P;
  const int64_t vol = (int64_t)D * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / P, pp = idx - n * P;
    PointCtx<S> p = make_point<S>(grid, idx, D, H, W, pad, align, smooth);
    // Same evaluation order as torch's grid_sampler_3d backward (which the reference kernel copies): signed sums of
    // v * w_b * w_c * gOut per corner, scaled once at the end by the coordinate multiplier (and smoothstep').
    S wgt[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) wgt[s] = p.ax.w[s & 1] * p.ay.w[(s >> 1) & 1] * p.az.w[(s >> 2) & 1];
    S gx = (S)0, gy = (S)0, gz = (S)0;
    const S* in_c = input + n * C * vol;
    const S* go_c = grad_output + n * C * P + pp;
    S* gi_c = grad_input ? grad_input + n * C * vol : nullptr;
    for (int64_t c = 0; c < C; ++c, in_c += vol, go_c += P) {
      const S go = __ldg(go_c);
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (p.off[s] < 0) continue;
        const int px = s & 1, py = (s >> 1) & 1, pz = (s >> 2) & 1;
        const S v = __ldg(in_c + p.off[s]);
        const S tx = v * p.ay.w[py] * p.az.w[pz] * go;
        const S ty = v * p.ax.w[px] * p.az.w[pz] * go;
        const S tz = v * p.ax.w[px] * p.ay.w[py] * go;
        gx = px ? gx + tx : gx - tx;
        gy = py ? gy + ty : gy - ty;
        gz = pz ? gz + tz : gz - tz;
        if (gi_c) atomicAdd(gi_c + p.off[s], go * wgt[s]);
      }
      if (gi_c) gi_c += vol;
    }
    gx *= p.ax.d[1];
    gy *= p.ay.d[1];
    gz *= p.az.d[1];
    grad_grid[idx * 3 + 0] = gx;
    grad_grid[idx * 3 + 1] = gy;
    grad_grid[idx * 3 + 2] = gz;
  }
}

template <typename S>
__global__ void __launch_bounds__(256) trilinear_bwd_bwd_kernel(
    const S* __restrict__ g_out_input, const S* __restrict__ g_out_grid, const S* __restrict__ input,
    const S* __restrict__ grid, const S* __restrict__ grad_output, S* __restrict__ grad_input, S* __restrict__ grad_grid,
    S* __restrict__ grad_grad_out, int64_t N, int64_t C, int D, int H, int W, int64_t P, int pad, bool align, bool smooth) {
  const int64_t total = N * P;
  const int64_t vol = (int64_t)D * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / P, pp = idx - n * P;
    PointCtx<S> p = make_point<S>(grid, idx, D, H, W, pad, align, smooth);
    const S ux = g_out_grid[idx * 3 + 0], uy = g_out_grid[idx * 3 + 1], uz = g_out_grid[idx * 3 + 2];
    // per-corner: w, grad(w).u, and (Hessian(w) u) per axis
    S wgt[8], dwu[8], hx[8], hy[8], hz[8], dwx[8], dwy[8], dwz[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      int px = s & 1, py = (s >> 1) & 1, pz = (s >> 2) & 1;
      const S wx = p.ax.w[px], wy = p.ay.w[py], wz = p.az.w[pz];
      const S dx = p.ax.d[px], dy = p.ay.d[py], dz = p.az.d[pz];
      wgt[s] = wx * wy * wz;
      dwx[s] = dx * wy * wz;
      dwy[s] = wx * dy * wz;
      dwz[s] = wx * wy * dz;
      dwu[s] = dwx[s] * ux + dwy[s] * uy + dwz[s] * uz;
      hx[s] = p.ax.dd[px] * wy * wz * ux + dx * dy * wz * uy + dx * wy * dz * uz;
      hy[s] = dx * dy * wz * ux + wx * p.ay.dd[py] * wz * uy + wx * dy * dz * uz;
      hz[s] = dx * wy * dz * ux + wx * dy * dz * uy + wx * wy * p.az.dd[pz] * uz;
    }
    S ggx = (S)0, ggy = (S)0, ggz = (S)0;
    const S* in_c = input + n * C * vol;
    const S* go_c = grad_output + n * C * P + pp;
    const S* goi_c = g_out_input ? g_out_input + n * C * vol : nullptr;
    S* gi_c = grad_input + n * C * vol;
    S* ggo_c = grad_grad_out + n * C * P + pp;
    for (int64_t c = 0; c < C; ++c, in_c += vol, go_c += P, gi_c += vol, ggo_c += P) {
      const S go = __ldg(go_c);
      S ggo = (S)0;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (p.off[s] < 0) continue;
        const S v = __ldg(in_c + p.off[s]);
        ggo += v * dwu[s];
        if (goi_c) {
          const S goi = __ldg(goi_c + p.off[s]);
          ggo += goi * wgt[s];
          ggx += go * dwx[s] * goi;
          ggy += go * dwy[s] * goi;
          ggz += go * dwz[s] * goi;
        }
        ggx += v * go * hx[s];
        ggy += v * go * hy[s];
        ggz += v * go * hz[s];
        atomicAdd(gi_c + p.off[s], go * dwu[s]);
      }
      *ggo_c = ggo;
      if (goi_c) goi_c += vol;
    }
    grad_grid[idx * 3 + 0] = ggx;
    grad_grid[idx * 3 + 1] = ggy;
    grad_grid[idx * 3 + 2] = ggz;
  }
}

inline bool shape_ok(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W, int64_t P) {
  return N >= 0 && C >= 0 && D > 0 && H > 0 && W > 0 && P >= 0 && D < (1 << 30) && H < (1 << 30) && W < (1 << 30);
}

}  // namespace

extern "C" {

int pv2_trilinear_fwd(const void* input, const void* grid, void* output, int64_t N, int64_t C, int64_t D, int64_t H,
                      int64_t W, int64_t P, int pad, int align, int smooth, int dtype, void* stream_) {
  PV2_CHECK_ARG(shape_ok(N, C, D, H, W, P) && pad >= 0 && pad <= 2);
  if (N * P == 0 || C == 0) return 0;
  PV2_CHECK_ARG(input && grid && output);
  cudaStream_t stream = (cudaStream_t)stream_;
  int g = pv2_grid_for(N * P, 256);
  if (dtype == PV2_F32)
    trilinear_fwd_kernel<float><<<g, 256, 0, stream>>>((const float*)input, (const float*)grid, (float*)output, N, C, (int)D, (int)H, (int)W, P, pad, align != 0, smooth != 0);
  else if (dtype == PV2_F64)
    trilinear_fwd_kernel<double><<<g, 256, 0, stream>>>((const double*)input, (const double*)grid, (double*)output, N, C, (int)D, (int)H, (int)W, P, pad, align != 0, smooth != 0);
  else
    return PV2_EUNSUPPORTED;
  PV2_DONE(1);
}

int pv2_trilinear_bwd(const void* grad_output, const void* input, const void* grid, void* grad_input, void* grad_grid,
                      int64_t N, int64_t C, int64_t D, int64_t H, int64_t W, int64_t P, int pad, int align, int smooth,
                      int dtype, void* stream_) {
  PV2_CHECK_ARG(shape_ok(N, C, D, H, W, P) && pad >= 0 && pad <= 2);
  if (N * P == 0) return 0;
  PV2_CHECK_ARG(grad_output && input && grid && grad_grid);
  cudaStream_t stream = (cudaStream_t)stream_;
  int g = pv2_grid_for(N * P, 256);
  if (dtype == PV2_F32)
    trilinear_bwd_kernel<float><<<g, 256, 0, stream>>>((const float*)grad_output, (const float*)input, (const float*)grid, (float*)grad_input, (float*)grad_grid, N, C, (int)D, (int)H, (int)W, P, pad, align != 0, smooth != 0);
  else if (dtype == PV2_F64)
    trilinear_bwd_kernel<double><<<g, 256, 0, stream>>>((const double*)grad_output, (const double*)input, (const double*)grid, (double*)grad_input, (double*)grad_grid, N, C, (int)D, (int)H, (int)W, P, pad, align != 0, smooth != 0);
  else
    return PV2_EUNSUPPORTED;
  PV2_DONE(1);
}

int pv2_trilinear_bwd_bwd(const void* g_out_input, const void* g_out_grid, const void* input, const void* grid,
                          const void* grad_output, void* grad_input, void* grad_grid, void* grad_grad_out, int64_t N,
                          int64_t C, int64_t D, int64_t H, int64_t W, int64_t P, int pad, int align, int smooth, int dtype,
                          void* stream_) {
  PV2_CHECK_ARG(shape_ok(N, C, D, H, W, P) && pad >= 0 && pad <= 2);
  if (N * P == 0) return 0;
  PV2_CHECK_ARG(g_out_grid && input && grid && grad_output && grad_input && grad_grid && grad_grad_out);
  cudaStream_t stream = (cudaStream_t)stream_;
  int g = pv2_grid_for(N * P, 256);
  if (dtype == PV2_F32)
    trilinear_bwd_bwd_kernel<float><<<g, 256, 0, stream>>>((const float*)g_out_input, (const float*)g_out_grid, (const float*)input, (const float*)grid, (const float*)grad_output, (float*)grad_input, (float*)grad_grid, (float*)grad_grad_out, N, C, (int)D, (int)H, (int)W, P, pad, align != 0, smooth != 0);
  else if (dtype == PV2_F64)
    trilinear_bwd_bwd_kernel<double><<<g, 256, 0, stream>>>((const double*)g_out_input, (const double*)g_out_grid, (const double*)input, (const double*)grid, (const double*)grad_output, (double*)grad_input, (double*)grad_grid, (double*)grad_grad_out, N, C, (int)D, (int)H, (int)W, P, pad, align != 0, smooth != 0);
  else
    return PV2_EUNSUPPORTED;
  PV2_DONE(1);
}

}  // extern "C"
