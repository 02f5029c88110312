// Field kernels of the NeuS renderer on a channels-last volume: trilinear feature fetch, analytic d sdf / d p,
// the colour head, and their hand-derived backward passes (first AND second order — the reference obtains these
// through autograd + the twice-differentiable sampler, fields/sdf_field.py:148-257, smooth_sampler_kernel.cu:39-619).
//
// Layout: volume [Z][Y][X][C] fp32 (torch.channels_last_3d view of the reference's (C,Z,Y,X)); a point's corner is one
// contiguous C*4-byte row, so a (half-)warp reads it with one coalesced 128-bit load per lane.  Points are the
// reference's normalised coordinates p in [0,1]^3; grid = 2p-1, zeros padding, align_corners=True (sdf_field.py:156-167).
//
// All of these are gather/scatter kernels bound by L2/HBM traffic: per point 8*C*4 B of corner rows (<= 4 KB at C=128,
// ~half of it shared with the neighbouring sample of the same ray) plus the per-point outputs.
#include "pv2_common.cuh"
#include <stdlib.h>

namespace {

struct Corner8 {
  int64_t off[8];  // row offset (in floats) of each corner, -1 when outside the volume
  float w[8];      // trilinear weights
  float dx[8], dy[8], dz[8];  // d w / d p (includes the (size-1) factor of the coordinate transform)
};

// exact fp32 replica of grid = 2p-1 followed by grid_sample's align_corners unnormalisation ((g+1)/2)*(size-1)
__device__ __forceinline__ void axis_setup(float p, int size, int& i0, float& t, float& scale) {
  const float g = p * 2.f - 1.f;
  const float x = ((g + 1.f) / 2.f) * (float)(size - 1);
  const float fl = floorf(x);
  i0 = (int)fl;
  t = x - fl;
  scale = (float)(size - 1);
}

__device__ __forceinline__ Corner8 make_corners(const float* __restrict__ pts, int64_t pt, int Z, int Y, int X, int C) {
  Corner8 c;
  int ix, iy, iz;
  float tx, ty, tz, sx, sy, sz;
  axis_setup(pts[pt * 3 + 0], X, ix, tx, sx);
  axis_setup(pts[pt * 3 + 1], Y, iy, ty, sy);
  axis_setup(pts[pt * 3 + 2], Z, iz, tz, sz);
  const float wx[2] = {1.f - tx, tx}, wy[2] = {1.f - ty, ty}, wz[2] = {1.f - tz, tz};
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int px = s & 1, py = (s >> 1) & 1, pz = (s >> 2) & 1;
    const int x = ix + px, y = iy + py, z = iz + pz;
    const bool in = (x >= 0) & (x < X) & (y >= 0) & (y < Y) & (z >= 0) & (z < Z);
    c.off[s] = in ? (((int64_t)z * Y + y) * X + x) * C : (int64_t)-1;
    c.w[s] = wx[px] * wy[py] * wz[pz];
    c.dx[s] = (px ? sx : -sx) * wy[py] * wz[pz];
    c.dy[s] = wx[px] * (py ? sy : -sy) * wz[pz];
    c.dz[s] = wx[px] * wy[py] * (pz ? sz : -sz);
  }
  return c;
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
  lo = __uint_as_float(l);
}

__device__ __forceinline__ float4 f4_fma(float a, float4 v, float4 acc) {
  return make_float4(fmaf(a, v.x, acc.x), fmaf(a, v.y, acc.y), fmaf(a, v.z, acc.z), fmaf(a, v.w, acc.w));
}

// ---------------------------------------------------------------------------------------------------------------
// Feature fetch.  LANES = c_use / 4 lanes cooperate on one point (16 for the 64-channel coarse pass, 32 for 128).
// Channels [0, ca) go to out_a in split-precision form (the SDF decoder's tensor-core input), channels [ca, c_use)
// to out_b plain.
template <int LANES>
__global__ void __launch_bounds__(256) field_sample_fwd_kernel(const float* __restrict__ vol, const float* __restrict__ pts,
                                                               int64_t P, int Z, int Y, int X, int C, int ca,
                                                               float* __restrict__ out_a, int64_t a_row, int64_t a_lo,
                                                               float* __restrict__ out_b, int64_t b_row) {
  const int sub = threadIdx.x % LANES;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / LANES;
  for (int64_t pt = group; pt < P; pt += ngroups) {
    const Corner8 c = make_corners(pts, pt, Z, Y, X, C);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (c.off[s] < 0) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(vol + c.off[s]) + sub);
      acc = f4_fma(c.w[s], v, acc);
    }
    const int ch = sub * 4;
    if (ch < ca) {
      float* dst = out_a + pt * a_row + ch;
      if (a_lo == 0) {   // plain fp32 (the tensor-core layers split on chip)
        *reinterpret_cast<float4*>(dst) = acc;
      } else {
        float4 h, l;
        split_tf32(acc.x, h.x, l.x); split_tf32(acc.y, h.y, l.y); split_tf32(acc.z, h.z, l.z); split_tf32(acc.w, h.w, l.w);
        *reinterpret_cast<float4*>(dst) = h;
        *reinterpret_cast<float4*>(dst + a_lo) = l;
      }
    } else {
      *reinterpret_cast<float4*>(out_b + pt * b_row + (ch - ca)) = acc;
    }
  }
}

template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LANES);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// grad = d sdf / d p = sum_i (dw_i/dp) (u . V_i[0:cs]),  rgb = sigmoid(Mr . [grad, f_r, geo, dir] + cr)
// (fields/sdf_field.py:226-257 with the colour decoder's two bias-connected linears folded into Mr, cr on the host).
// 16 lanes per point: lane l owns sdf-half channels 4l..4l+3.   cs = 64 (C/2), geo = 64.
__global__ void __launch_bounds__(256) field_post_fwd_kernel(const float* __restrict__ vol, const float* __restrict__ pts,
                                                             const float* __restrict__ dirs, int samples_per_ray,
                                                             const float* __restrict__ u, const float* __restrict__ f_r,
                                                             const float* __restrict__ out_geo, int64_t geo_row,
                                                             const float* __restrict__ Mr, const float* __restrict__ cr,
                                                             int64_t P, int Z, int Y, int X, int C,
                                                             float* __restrict__ grad, float* __restrict__ rgb) {
  constexpr int LANES = 16;
  __shared__ float Ms[3 * 134 + 3];
  for (int i = threadIdx.x; i < 3 * 134 + 3; i += blockDim.x) Ms[i] = (i < 402) ? Mr[i] : (cr ? cr[i - 402] : 0.f);
  __syncthreads();
  const int sub = threadIdx.x % LANES;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / LANES;
  const int64_t iters = (P + ngroups - 1) / ngroups;
  for (int64_t itn = 0; itn < iters; ++itn) {
    const int64_t pt = group + itn * ngroups;
    const bool live = pt < P;
    const int64_t q = live ? pt : 0;
    const Corner8 c = make_corners(pts, q, Z, Y, X, C);
    const float4 uu = __ldg(reinterpret_cast<const float4*>(u + q * 64) + sub);
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      float t = 0.f;
      if (c.off[s] >= 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(vol + c.off[s]) + sub);
        t = uu.x * v.x + uu.y * v.y + uu.z * v.z + uu.w * v.w;
      }
      t = group_sum<LANES>(t);
      gx = fmaf(c.dx[s], t, gx); gy = fmaf(c.dy[s], t, gy); gz = fmaf(c.dz[s], t, gz);
    }
    // colour head: inputs [grad(3) | f_r(64) | geo(64) | dir(3)]
    const float4 fr = __ldg(reinterpret_cast<const float4*>(f_r + q * 64) + sub);
    const float* gp = out_geo + q * geo_row + sub * 4;
    const float ge[4] = {__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), __ldg(gp + 3)};
    const int64_t ray = q / samples_per_ray;
    const float d0 = __ldg(dirs + ray * 3), d1 = __ldg(dirs + ray * 3 + 1), d2 = __ldg(dirs + ray * 3 + 2);
    float z[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* m = Ms + k * 134;
      float a = m[3 + sub * 4] * fr.x + m[4 + sub * 4] * fr.y + m[5 + sub * 4] * fr.z + m[6 + sub * 4] * fr.w;
      a += m[67 + sub * 4] * ge[0] + m[68 + sub * 4] * ge[1] + m[69 + sub * 4] * ge[2] + m[70 + sub * 4] * ge[3];
      a = group_sum<LANES>(a);
      a += m[0] * gx + m[1] * gy + m[2] * gz + m[131] * d0 + m[132] * d1 + m[133] * d2 + Ms[402 + k];
      z[k] = 1.f / (1.f + expf(-a));
    }
    if (live && sub == 0) {
      grad[pt * 3 + 0] = gx; grad[pt * 3 + 1] = gy; grad[pt * 3 + 2] = gz;
      rgb[pt * 3 + 0] = z[0]; rgb[pt * 3 + 1] = z[1]; rgb[pt * 3 + 2] = z[2];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of field_post.  Per point:
//   zbar = g_rgb * rgb (1 - rgb);  inbar = Mr^T zbar;  gbar = g_grad + inbar[0:3]          (total gradient on d sdf/d p)
//   d f_r = inbar[3:67]  -> dF[:, cs:2cs];  d geo = inbar[67:131], d sdf = g_sdf  -> doutbar [P, 72] (68 used)
//   ubar  = sum_i (dw_i . gbar) V_i[0:cs]  -> [P,64]  (second-order term: d grad / d u)
//   dMr  += zbar (x) in,  dcr += zbar     (block partial sums, then atomics)
__global__ void __launch_bounds__(256, 2) field_post_bwd_kernel(
    const float* __restrict__ vol, const float* __restrict__ pts, const float* __restrict__ dirs, int samples_per_ray,
    const float* __restrict__ f_r, const float* __restrict__ out_geo, int64_t geo_row, const float* __restrict__ grad,
    const float* __restrict__ rgb, const float* __restrict__ Mr, const float* __restrict__ g_rgb,
    const float* __restrict__ g_grad, const float* __restrict__ g_sdf, int64_t P, int Z, int Y, int X, int C,
    float* __restrict__ gbar_out, float* __restrict__ dF, int64_t dF_row, float* __restrict__ doutbar,
    float* __restrict__ ubar, float* __restrict__ dMr, float* __restrict__ dcr, float* __restrict__ ubar_sum,
    float* __restrict__ dout_sum) {
  constexpr int LANES = 16;
  __shared__ float Ms[402];
  __shared__ float acc_s[405 + 64 + 68];   // dMr (402) | dcr (3) | column sums of ubar (64) | of doutbar (68)
  for (int i = threadIdx.x; i < 402; i += blockDim.x) Ms[i] = Mr[i];
  for (int i = threadIdx.x; i < 405 + 64 + 68; i += blockDim.x) acc_s[i] = 0.f;
  __syncthreads();
  const int sub = threadIdx.x % LANES;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / LANES;
  const int64_t iters = (P + ngroups - 1) / ngroups;
  // per-lane partial sums of dMr: rows k, columns {f_r 4 ch, geo 4 ch}; lane 0 also carries grad/dir/bias columns
  float am[3][8];
  float a0[3][7];
  float us[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f}, ds0 = 0.f;   // bias gradients: column sums
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) am[k][i] = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) a0[k][i] = 0.f;
  }
  for (int64_t itn = 0; itn < iters; ++itn) {
    const int64_t pt = group + itn * ngroups;
    if (pt >= P) continue;  // whole 16-lane group exits together; no shuffles below cross groups of different pt
    const Corner8 c = make_corners(pts, pt, Z, Y, X, C);
    float zb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float r = __ldg(rgb + pt * 3 + k);
      zb[k] = __ldg(g_rgb + pt * 3 + k) * r * (1.f - r);
    }
    float gb[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      gb[a] = __ldg(g_grad + pt * 3 + a) + Ms[a] * zb[0] + Ms[134 + a] * zb[1] + Ms[268 + a] * zb[2];
    // d f_r and d geo for this lane's 4 channels
    float dfr[4], dge[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cf = 3 + sub * 4 + i, cg = 67 + sub * 4 + i;
      dfr[i] = Ms[cf] * zb[0] + Ms[134 + cf] * zb[1] + Ms[268 + cf] * zb[2];
      dge[i] = Ms[cg] * zb[0] + Ms[134 + cg] * zb[1] + Ms[268 + cg] * zb[2];
    }
    *reinterpret_cast<float4*>(dF + pt * dF_row + 64 + sub * 4) = make_float4(dfr[0], dfr[1], dfr[2], dfr[3]);
    // doutbar row (72 floats, plain fp32): col 0 = d sdf, cols 1..64 = d geo, cols 65..71 = 0 (the row is padded to a
    // multiple of 8 channels so that the linears consuming it run on the bf16x3 tensor-core kernel)
    {
      float* row = doutbar + pt * 72;
#pragma unroll
      for (int i = 0; i < 4; ++i) { row[1 + sub * 4 + i] = dge[i]; ds[i] += dge[i]; }
      if (sub == 0) {
        row[0] = __ldg(g_sdf + pt);
        ds0 += row[0];
        row[65] = row[66] = row[67] = 0.f;
        *reinterpret_cast<float4*>(row + 68) = make_float4(0.f, 0.f, 0.f, 0.f);
        gbar_out[pt * 3 + 0] = gb[0]; gbar_out[pt * 3 + 1] = gb[1]; gbar_out[pt * 3 + 2] = gb[2];
      }
    }
    // ubar = sum_i (dw_i . gbar) V_i
    float4 ub = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (c.off[s] < 0) continue;
      const float coef = c.dx[s] * gb[0] + c.dy[s] * gb[1] + c.dz[s] * gb[2];
      const float4 v = __ldg(reinterpret_cast<const float4*>(vol + c.off[s]) + sub);
      ub = f4_fma(coef, v, ub);
    }
    *reinterpret_cast<float4*>(ubar + pt * 64 + sub * 4) = ub;
    us[0] += ub.x; us[1] += ub.y; us[2] += ub.z; us[3] += ub.w;
    // dMr partial sums
    const float4 fr = __ldg(reinterpret_cast<const float4*>(f_r + pt * 64) + sub);
    const float* gp = out_geo + pt * geo_row + sub * 4;
    const float inl[8] = {fr.x, fr.y, fr.z, fr.w, __ldg(gp), __ldg(gp + 1), __ldg(gp + 2), __ldg(gp + 3)};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) am[k][i] = fmaf(zb[k], inl[i], am[k][i]);
    if (sub == 0) {
      const int64_t ray = pt / samples_per_ray;
      const float in0[7] = {__ldg(grad + pt * 3), __ldg(grad + pt * 3 + 1), __ldg(grad + pt * 3 + 2),
                            __ldg(dirs + ray * 3), __ldg(dirs + ray * 3 + 1), __ldg(dirs + ray * 3 + 2), 1.f};
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 7; ++i) a0[k][i] = fmaf(zb[k], in0[i], a0[k][i]);
    }
  }
  // block reduction through shared memory, then one atomic per entry per block
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(&acc_s[k * 134 + 3 + sub * 4 + i], am[k][i]);
      atomicAdd(&acc_s[k * 134 + 67 + sub * 4 + i], am[k][4 + i]);
    }
    if (sub == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        atomicAdd(&acc_s[k * 134 + i], a0[k][i]);
        atomicAdd(&acc_s[k * 134 + 131 + i], a0[k][3 + i]);
      }
      atomicAdd(&acc_s[402 + k], a0[k][6]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    atomicAdd(&acc_s[405 + sub * 4 + i], us[i]);
    atomicAdd(&acc_s[405 + 64 + 1 + sub * 4 + i], ds[i]);
  }
  if (sub == 0) atomicAdd(&acc_s[405 + 64], ds0);
  __syncthreads();
  for (int i = threadIdx.x; i < 405 + 64 + 68; i += blockDim.x) {
    const float v = acc_s[i];
    if (v == 0.f) continue;
    if (i < 402) atomicAdd(&dMr[i], v);
    else if (i < 405) atomicAdd(&dcr[i - 402], v);
    else if (i < 405 + 64) { if (ubar_sum != nullptr) atomicAdd(&ubar_sum[i - 405], v); }
    else if (dout_sum != nullptr) atomicAdd(&dout_sum[i - 405 - 64], v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Volume gradient: dV[corner_i, c] += w_i dF[p, c]  (all C channels)  +  [c < cs] (dw_i . gbar) u[p, c]
// One warp per point, lane = 4 channels, 128-bit vector reductions (red.global.add.v4.f32).
__global__ void __launch_bounds__(256) field_sample_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ dF,
                                                               int64_t dF_row, const float* __restrict__ u,
                                                               const float* __restrict__ gbar, int64_t P, int Z, int Y,
                                                               int X, int C, int cs, float* __restrict__ dvol) {
  const int lanes = C / 4;  // 32 for C = 128
  const int sub = threadIdx.x % lanes;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / lanes;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / lanes;
  for (int64_t pt = group; pt < P; pt += ngroups) {
    const Corner8 c = make_corners(pts, pt, Z, Y, X, C);
    const float4 g = __ldg(reinterpret_cast<const float4*>(dF + pt * dF_row) + sub);
    float4 uu = make_float4(0.f, 0.f, 0.f, 0.f);
    float gb0 = 0.f, gb1 = 0.f, gb2 = 0.f;
    const bool sdf_half = (sub * 4 < cs) && (u != nullptr);
    if (sdf_half) {
      uu = __ldg(reinterpret_cast<const float4*>(u + pt * cs) + sub);
      gb0 = __ldg(gbar + pt * 3); gb1 = __ldg(gbar + pt * 3 + 1); gb2 = __ldg(gbar + pt * 3 + 2);
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (c.off[s] < 0) continue;
      const float coef = c.dx[s] * gb0 + c.dy[s] * gb1 + c.dz[s] * gb2;
      const float w = c.w[s];
      float* dst = dvol + c.off[s] + sub * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(fmaf(coef, uu.x, w * g.x)),
                   "f"(fmaf(coef, uu.y, w * g.y)), "f"(fmaf(coef, uu.z, w * g.z)), "f"(fmaf(coef, uu.w, w * g.w))
                   : "memory");
    }
  }
}

inline bool vol_ok(int Z, int Y, int X, int C) { return Z > 0 && Y > 0 && X > 0 && C > 0 && (C % 4) == 0; }

}  // namespace

extern "C" {

int pv2_field_sample_fwd(const float* vol, const float* pts, int64_t P, int Z, int Y, int X, int C, int c_use, int ca,
                         float* out_a, int64_t a_row, int64_t a_lo, float* out_b, int64_t b_row, void* stream_) {
  PV2_CHECK_ARG(P >= 0 && vol_ok(Z, Y, X, C) && (c_use == 64 || c_use == 128) && c_use <= C && ca >= 0 && ca <= c_use &&
                (ca % 4) == 0);
  if (P == 0) return 0;
  PV2_CHECK_ARG(vol && pts && (ca == 0 || out_a) && (ca == c_use || out_b));
  cudaStream_t stream = (cudaStream_t)stream_;
  const int lanes = c_use / 4;
  const int grid = pv2_grid_for(P * lanes, 256);
  if (lanes == 16)
    field_sample_fwd_kernel<16><<<grid, 256, 0, stream>>>(vol, pts, P, Z, Y, X, C, ca, out_a, a_row, a_lo, out_b, b_row);
  else
    field_sample_fwd_kernel<32><<<grid, 256, 0, stream>>>(vol, pts, P, Z, Y, X, C, ca, out_a, a_row, a_lo, out_b, b_row);
  PV2_DONE(1);
}

int pv2_field_post_fwd(const float* vol, const float* pts, const float* dirs, int samples_per_ray, const float* u,
                       const float* f_r, const float* out_geo, int64_t geo_row, const float* Mr, const float* cr, int64_t P,
                       int Z, int Y, int X, int C, float* grad, float* rgb, void* stream_) {
  PV2_CHECK_ARG(P >= 0 && vol_ok(Z, Y, X, C) && C >= 64 && samples_per_ray > 0);
  if (P == 0) return 0;
  PV2_CHECK_ARG(vol && pts && dirs && u && f_r && out_geo && Mr && grad && rgb);
  field_post_fwd_kernel<<<pv2_grid_for(P * 16, 256), 256, 0, (cudaStream_t)stream_>>>(
      vol, pts, dirs, samples_per_ray, u, f_r, out_geo, geo_row, Mr, cr, P, Z, Y, X, C, grad, rgb);
  PV2_DONE(1);
}

int pv2_field_post_bwd(const float* vol, const float* pts, const float* dirs, int samples_per_ray, const float* f_r,
                       const float* out_geo, int64_t geo_row, const float* grad, const float* rgb, const float* Mr,
                       const float* g_rgb, const float* g_grad, const float* g_sdf, int64_t P, int Z, int Y, int X, int C,
                       float* gbar, float* dF, int64_t dF_row, float* doutbar, float* ubar, float* dMr, float* dcr,
                       float* ubar_sum, float* dout_sum, void* stream_) {
  PV2_CHECK_ARG(P >= 0 && vol_ok(Z, Y, X, C) && C >= 64 && samples_per_ray > 0);
  if (P == 0) return 0;
  PV2_CHECK_ARG(vol && pts && dirs && f_r && out_geo && grad && rgb && Mr && g_rgb && g_grad && g_sdf && gbar && dF &&
                doutbar && ubar && dMr && dcr);
  // bounded grid: every block ends with ~400 atomics for the colour-head weight gradient
  field_post_bwd_kernel<<<pv2_grid_for(P * 16, 256, 2), 256, 0, (cudaStream_t)stream_>>>(
      vol, pts, dirs, samples_per_ray, f_r, out_geo, geo_row, grad, rgb, Mr, g_rgb, g_grad, g_sdf, P, Z, Y, X, C, gbar,
      dF, dF_row, doutbar, ubar, dMr, dcr, ubar_sum, dout_sum);
  PV2_DONE(1);
}

int pv2_field_sample_bwd(const float* pts, const float* dF, int64_t dF_row, const float* u, const float* gbar, int64_t P,
                         int Z, int Y, int X, int C, int cs, float* dvol, void* stream_) {
  PV2_CHECK_ARG(P >= 0 && vol_ok(Z, Y, X, C) && C <= 128 && (256 % (C / 4)) == 0 && cs >= 0 && cs <= C);
  if (P == 0) return 0;
  PV2_CHECK_ARG(pts && dF && dvol && (u == nullptr || gbar != nullptr));
  field_sample_bwd_kernel<<<pv2_grid_for(P * (C / 4), 256), 256, 0, (cudaStream_t)stream_>>>(pts, dF, dF_row, u, gbar, P,
                                                                                            Z, Y, X, C, cs, dvol);
  PV2_DONE(1);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of a dense per-row linear layer:  dw[co, ci] (+)= sum_j dy[j, co] * x[j, ci]   (fp32 FMA, exact)
// x / dy may be given in split-precision form (value = hi + lo, lo at +lo_off); lo_off = 0 means plain.
namespace {

__global__ void __launch_bounds__(256) dense_wgrad_kernel(const float* __restrict__ x, int64_t x_row, int64_t x_lo,
                                                          const float* __restrict__ dy, int64_t dy_row, int64_t dy_lo,
                                                          int64_t rows, int cin, int cout, int64_t rows_per_chunk,
                                                          int ci_tiles, float* __restrict__ dw) {
  constexpr int TN = 64, RC = 16, PAD = 4;
  __shared__ float Ds[RC][TN + PAD];
  __shared__ float Xs[RC][TN + PAD];
  const int tid = threadIdx.x;
  const int co0 = (blockIdx.y / ci_tiles) * TN;
  const int ci0 = (blockIdx.y % ci_tiles) * TN;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_chunk;
  int64_t r_end = r_begin + rows_per_chunk;
  if (r_end > rows) r_end = rows;
  const int tr = (tid / 16) * 4, tc = (tid % 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int lr = tid / 16, lc = (tid % 16) * 4;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += RC) {
    const int64_t j = r0 + lr;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int co = co0 + lc + e, ci = ci0 + lc + e;
      float dv = 0.f, xv = 0.f;
      if (j < r_end) {
        if (co < cout) { dv = dy[j * dy_row + co]; if (dy_lo) dv += dy[j * dy_row + dy_lo + co]; }
        if (ci < cin) { xv = x[j * x_row + ci]; if (x_lo) xv += x[j * x_row + x_lo + ci]; }
      }
      Ds[lr][lc + e] = dv;
      Xs[lr][lc + e] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) {
      const float4 a = *reinterpret_cast<const float4*>(&Ds[rr][tr]);
      const float4 b = *reinterpret_cast<const float4*>(&Xs[rr][tc]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jx = 0; jx < 4; ++jx) acc[i][jx] = fmaf(av[i], bv[jx], acc[i][jx]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + tr + i;
    if (co >= cout) continue;
#pragma unroll
    for (int jx = 0; jx < 4; ++jx) {
      const int ci = ci0 + tc + jx;
      if (ci < cin) atomicAdd(&dw[(int64_t)co * cin + ci], acc[i][jx]);
    }
  }
}

}  // namespace

extern "C" int pv2_wgrad_umma(const float*, int64_t, int64_t, const float*, int64_t, int64_t, const int32_t*, const int32_t*,
                              const uint8_t*, float*, int64_t, int64_t, int, int, int, void*, size_t, void*);

extern "C" int pv2_dense_wgrad(const float* x, int64_t x_row, int64_t x_lo_off, const float* dy, int64_t dy_row,
                               int64_t dy_lo_off, int64_t rows, int cin, int cout, float* dw, void* workspace,
                               size_t workspace_bytes, void* stream_) {
  {
    const char* e = getenv("PV2_FORCE_SIMT");
    if (!(e && e[0] == '1') && cin >= 96) {
      int rc = pv2_wgrad_umma(x, x_row, x_lo_off, dy, dy_row, dy_lo_off, nullptr, nullptr, nullptr, dw, rows, rows, cin, cout, 1, workspace,
                              workspace_bytes, stream_);
      if (rc != PV2_EUNSUPPORTED && rc != PV2_EWORKSPACE) return rc;
    }
  }
  PV2_CHECK_ARG(rows >= 0 && cin > 0 && cout > 0);
  if (rows == 0) return 0;
  PV2_CHECK_ARG(x && dy && dw);
  const int co_tiles = (cout + 63) / 64, ci_tiles = (cin + 63) / 64;
  int64_t chunks = (4LL * PV2_SM_COUNT * 4 + co_tiles * ci_tiles - 1) / (co_tiles * ci_tiles);
  const int64_t max_chunks = (rows + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int64_t rpc = ((rows + chunks - 1) / chunks + 15) / 16 * 16;
  chunks = (rows + rpc - 1) / rpc;
  dim3 grid((unsigned)chunks, (unsigned)(co_tiles * ci_tiles));
  dense_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, x_row, x_lo_off, dy, dy_row, dy_lo_off, rows, cin, cout,
                                                             rpc, ci_tiles, dw);
  PV2_DONE(1);
}
