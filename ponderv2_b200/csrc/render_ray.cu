// Per-ray kernels of the NeuS renderer: everything between "rays in" and "sample points for the field", and between
// "field values per sample" and "loss terms" (SURVEY §8a rows a8-a13 and a18-a21).  The reference runs these as ~150
// small torch kernels forward and ~300 backward (scene_colliders.py:38-99, ray_samplers.py:55-107,227-463,
// rays.py:83-153, sdf_field.py:122-146, renderers.py:5-75, base_surface_model.py:102-211); here they are four kernels
// forward and two backward, one warp per ray, scans done with warp shuffles:
//
//   ray_setup      collider (slab test) + stratified spacing bins + coarse sample points             (no gradient)
//   ray_resample   fixed-inv_s alphas -> weights -> PDF inverse-CDF resampling -> sorted merge ->
//                  starts / deltas / normalised fine sample points, global min/max of the starts    (no gradient; the
//                  reference detaches the samples, ray_samplers.py:321)
//   ray_composite  NeuS alpha (SingleVariance inv_s) -> transmittance -> weights -> rgb / depth / normal   fwd + bwd
//   ray_loss       depth-L1, rgb-L1 (+mse for psnr), free-space, sdf and eikonal partial sums                fwd + bwd
//
// All of it is latency/HBM-bound elementwise + scan work on [R, S] arrays (24 B per ray in, ~40 B per sample out).
#include "pv2_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxS0 = 256;      // coarse samples per ray (C4: 192)
constexpr int kMaxNb = 128;      // importance bins per ray (Si + 1; C4: 65)
constexpr int kMaxRounds = 8;    // S <= 256 fine samples per ray

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// torch.linspace(start, end, steps) element `i` in fp32 (ATen RangeFactories.cu: symmetric around the midpoint)
__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
  const float step = (end - start) / (float)(steps - 1);
  return (i < steps / 2) ? start + step * (float)i : end - step * (float)(steps - i - 1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// inclusive scans across the 32 lanes
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v *= t; }
  return v;
}
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
  return v;
}
// inclusive suffix sum (lane l gets v_l + v_{l+1} + ... + v_31)
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v += t; }
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// AABBBoxCollider (scene_colliders.py:38-99) for one ray
__device__ __forceinline__ void collide(const float* __restrict__ o, const float* __restrict__ d, const float* bbox,
                                        float near_plane, float& near, float& far) {
  float tn = -INFINITY, tf = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float inv = 1.f / (d[a] + 1e-6f);
    const float ta = (bbox[a] - o[a]) * inv, tb = (bbox[3 + a] - o[a]) * inv;
    tn = fmaxf(tn, fminf(ta, tb));
    tf = fminf(tf, fmaxf(ta, tb));
  }
  tn = fmaxf(tn, near_plane);
  const bool hit = tn < tf;
  near = hit ? tn : 0.f;
  far = hit ? tf : 0.f;
}

struct Bbox6 { float v[6]; };

// UniformSampler (ray_samplers.py:55-107): bins[r, i], i in [0, S0], and the coarse points o + d * euclid(bins[i]), i < S0
__global__ void ray_setup_kernel(const float* __restrict__ origins, const float* __restrict__ dirs,
                                 const float* __restrict__ noise, int noise_cols, int64_t R, int S0, Bbox6 bbox,
                                 float near_plane, float* __restrict__ nears, float* __restrict__ fars,
                                 float* __restrict__ bins, float* __restrict__ pts) {
  const int steps = S0 + 1;
  const int64_t total = R * steps;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / steps;
    const int i = (int)(idx - r * steps);
    const float o[3] = {origins[r * 3], origins[r * 3 + 1], origins[r * 3 + 2]};
    const float d[3] = {dirs[r * 3], dirs[r * 3 + 1], dirs[r * 3 + 2]};
    float near, far;
    collide(o, d, bbox.v, near_plane, near, far);
    float b = linspace_at(0.f, 1.f, steps, i);
    if (noise != nullptr) {
      const float lower = (i == 0) ? b : (b + linspace_at(0.f, 1.f, steps, i - 1)) / 2.f;
      const float upper = (i == S0) ? b : (linspace_at(0.f, 1.f, steps, i + 1) + b) / 2.f;
      const float t = noise[r * noise_cols + (noise_cols == 1 ? 0 : i)];
      b = lower + (upper - lower) * t;
    }
    bins[idx] = b;
    if (i == 0) { nears[r] = near; fars[r] = far; }
    if (i < S0) {
      const float e = b * far + (1.f - b) * near;
      float* p = pts + (r * S0 + i) * 3;
      p[0] = o[0] + d[0] * e; p[1] = o[1] + d[1] * e; p[2] = o[2] + d[2] * e;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct ResampleParams {
  const float* origins; const float* dirs; const float* nears; const float* fars;
  const float* bins;        // [R, S0+1] spacing bins
  const float* sdf;         // [R, S0] coarse sdf
  const float* noise;       // [R, noise_cols] or nullptr (eval: bin centres)
  int noise_cols;           // nb or 1
  int64_t R;
  int S0, Si;
  float inv_s;              // base_variance * 2^iter (ray_samplers.py:392)
  float u_end;              // float(1 - 1/nb)
  float u_center;           // float(1/(2 nb))
  int norm_pts; float norm_scale;   // 1 + norm_padding + 1e-3 (sdf_field.py:58-74)
  float* starts; float* deltas;     // [R, S]
  float* pts_norm;                  // [R, S, 3]
  float* init_weights;              // [R, S0]
  float* new_bins;                  // [R, Si] spacing positions of the importance samples
  int* minmax;                      // [2] float bits: min / max of all starts (non-negative floats order like ints)
};

// NeuSSampler.generate_ray_samples with num_upsample_steps = 1 (ray_samplers.py:355-463), PDFSampler (:227-322) and
// merge_ray_samples (rays.py:118-153).  One warp per ray.
__global__ void __launch_bounds__(kWarpsPerBlock * 32) ray_resample_kernel(const ResampleParams p) {
  __shared__ float s_bins[kWarpsPerBlock][kMaxS0 + 1];
  __shared__ float s_sdf[kWarpsPerBlock][kMaxS0];
  __shared__ float s_w[kWarpsPerBlock][kMaxS0];
  __shared__ float s_cdf[kWarpsPerBlock][kMaxS0 + 1];
  __shared__ float s_new[kWarpsPerBlock][kMaxNb];
  __shared__ float s_merged[kWarpsPerBlock][kMaxS0 + kMaxNb];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S0 = p.S0, Si = p.Si, nb = p.Si + 1, S = p.S0 + p.Si;
  float* bs = s_bins[wib]; float* sd = s_sdf[wib]; float* w = s_w[wib]; float* cdf = s_cdf[wib];
  float* nw = s_new[wib]; float* mg = s_merged[wib];
  float vmin = INFINITY, vmax = -INFINITY;
  for (int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + wib; r < p.R; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    const float near = p.nears[r], far = p.fars[r];
    for (int i = lane; i <= S0; i += 32) bs[i] = p.bins[r * (S0 + 1) + i];
    for (int i = lane; i < S0; i += 32) sd[i] = p.sdf[r * S0 + i];
    __syncwarp();
    auto euclid = [&](float b) { return b * far + (1.f - b) * near; };
    // raw cos of every section, stored in w[] for the moment
    for (int i = lane; i < S0 - 1; i += 32) {
      const float dlt = euclid(bs[i + 1]) - euclid(bs[i]);
      w[i] = (sd[i + 1] - sd[i]) / (dlt + 1e-5f);
    }
    __syncwarp();
    // alphas -> weights, sequential rounds of 32 with a running transmittance
    float carry = 1.f;
    const int n_alpha = S0 - 1;
    for (int base = 0; base < S0; base += 32) {
      const int i = base + lane;
      float alpha = 0.f;
      if (i < n_alpha) {
        const float dlt = euclid(bs[i + 1]) - euclid(bs[i]);
        const float mid = (sd[i] + sd[i + 1]) * 0.5f;
        const float prev_cos = (i == 0) ? 0.f : w[i - 1];
        float c = fminf(prev_cos, w[i]);
        c = fminf(fmaxf(c, -1e3f), 0.f);
        const float pc = sigmoidf_((mid - c * dlt * 0.5f) * p.inv_s);
        const float nc = sigmoidf_((mid + c * dlt * 0.5f) * p.inv_s);
        alpha = (pc - nc + 1e-5f) / (pc + 1e-5f);
      }
      const float f = (i < n_alpha) ? (1.f - alpha + 1e-7f) : 1.f;
      const float incl = warp_scan_mul(f, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.f;
      const float wt = (i < n_alpha) ? alpha * carry * excl : 0.f;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      __syncwarp();            // every lane has consumed w[i-1], w[i] (the raw cos) of this round
      if (i < S0) mg[i] = wt;  // staged; w[] still holds cos values needed by the next round
    }
    __syncwarp();
    float wsum = 0.f;
    for (int i = lane; i < S0; i += 32) { w[i] = mg[i]; wsum += mg[i]; }
    wsum = warp_sum(wsum);
    __syncwarp();
    if (p.init_weights != nullptr)
      for (int i = lane; i < S0; i += 32) p.init_weights[r * S0 + i] = w[i];
    // pdf / cdf (ray_samplers.py:241-255)
    const float pad = fmaxf(1e-5f - wsum, 0.f);
    const float denom = wsum + pad;
    float run = 0.f;
    if (lane == 0) cdf[0] = 0.f;
    for (int base = 0; base < S0; base += 32) {
      const int i = base + lane;
      const float pdf = (i < S0) ? (w[i] + pad / (float)S0) / denom : 0.f;
      const float incl = warp_scan_add(pdf, lane);
      if (i < S0) cdf[i + 1] = fminf(1.f, run + incl);
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    // inverse-CDF samples
    for (int j = lane; j < nb; j += 32) {
      float u = linspace_at(0.f, p.u_end, nb, j);
      if (p.noise != nullptr) u = u + p.noise[r * p.noise_cols + (p.noise_cols == 1 ? 0 : j)] / (float)nb;
      else u = u + p.u_center;
      int lo = 0, hi = S0 + 1;   // searchsorted(right=True): number of cdf entries <= u
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
      const int below = min(max(lo - 1, 0), S0), above = min(max(lo, 0), S0);
      const float c0 = cdf[below], c1 = cdf[above], b0 = bs[below], b1 = bs[above];
      float den = c1 - c0;
      if (den < 1e-5f) den = 1.f;
      const float t = fminf(fmaxf((u - c0) / den, 0.f), 1.f);
      nw[j] = b0 + t * (b1 - b0);
    }
    __syncwarp();
    if (p.new_bins != nullptr)
      for (int j = lane; j < Si; j += 32) p.new_bins[r * Si + j] = nw[j];
    const float end_sp = fmaxf(bs[S0], nw[nb - 1]);
    // merge the two ascending lists (existing first on ties)
    for (int i = lane; i < S0; i += 32) {
      const float v = bs[i];
      int lo = 0, hi = Si;       // number of new samples < v
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (nw[mid] < v) lo = mid + 1; else hi = mid; }
      mg[i + lo] = v;
    }
    for (int j = lane; j < Si; j += 32) {
      const float v = nw[j];
      int lo = 0, hi = S0;       // number of existing samples <= v
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (bs[mid] <= v) lo = mid + 1; else hi = mid; }
      mg[j + lo] = v;
    }
    __syncwarp();
    const float o0 = p.origins[r * 3], o1 = p.origins[r * 3 + 1], o2 = p.origins[r * 3 + 2];
    const float d0 = p.dirs[r * 3], d1 = p.dirs[r * 3 + 1], d2 = p.dirs[r * 3 + 2];
    for (int s = lane; s < S; s += 32) {
      const float st = euclid(mg[s]);
      const float en = euclid(s + 1 < S ? mg[s + 1] : end_sp);
      p.starts[r * S + s] = st;
      p.deltas[r * S + s] = en - st;
      float q[3] = {o0 + d0 * st, o1 + d1 * st, o2 + d2 * st};
      if (p.norm_pts) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          float v = q[a] / p.norm_scale + 0.5f;
          if (v >= 1.f) v = 1.f - 10e-4f;
          if (v < 0.f) v = 0.f;
          q[a] = v;
        }
      }
      float* dst = p.pts_norm + (r * S + s) * 3;
      dst[0] = q[0]; dst[1] = q[1]; dst[2] = q[2];
      vmin = fminf(vmin, st); vmax = fmaxf(vmax, st);
    }
    __syncwarp();
  }
  vmin = warp_min(vmin); vmax = warp_max(vmax);
  if (lane == 0 && vmax >= vmin) {
    atomicMin(&p.minmax[0], __float_as_int(vmin));
    atomicMax(&p.minmax[1], __float_as_int(vmax));
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct CompositeParams {
  const float* sdf;      // [R, S]
  const float* grad;     // [R, S, 3]
  const float* rgbs;     // [R, S, 3] or nullptr
  const float* starts; const float* deltas;   // [R, S]
  const float* dirs;     // [R, 3]
  const float* variance; // [1]: inv_s = clip(exp(10 v), 1e-6, 1e6)   (sdf_field.py:45-55)
  const int* minmax;     // float bits of the global min / max start (renderers.py:49-50)
  float cos_anneal;      // SDFField._cos_anneal_ratio (1: never updated by the reference)
  int64_t R; int S;
  int clamp_rgb;         // eval: rgb.clamp(0, 1)
  // forward outputs
  float* weights;        // [R, S]
  float* rgb;            // [R, 3]
  float* depth;          // [R]
  float* normal;         // [R, 3]
  // backward inputs (may be nullptr) / outputs
  const float* g_rgb; const float* g_depth; const float* g_normal; const float* g_weights;
  float* g_sdf; float* g_grad; float* g_rgbs; float* g_variance;
};

struct AlphaTerms { float alpha, raw, pc, nc, prv, nxt, c; };

__device__ __forceinline__ AlphaTerms neus_alpha(float sdf, float gx, float gy, float gz, float d0, float d1, float d2,
                                                 float delta, float inv_s, float r) {
  AlphaTerms t;
  t.c = d0 * gx + d1 * gy + d2 * gz;
  const float ic = -(fmaxf(-t.c * 0.5f + 0.5f, 0.f) * (1.f - r) + fmaxf(-t.c, 0.f) * r);
  t.nxt = sdf + ic * delta * 0.5f;
  t.prv = sdf - ic * delta * 0.5f;
  t.pc = sigmoidf_(t.prv * inv_s);
  t.nc = sigmoidf_(t.nxt * inv_s);
  t.raw = (t.pc - t.nc + 1e-5f) / (t.pc + 1e-5f);
  t.alpha = fminf(fmaxf(t.raw, 0.f), 1.f);
  return t;
}

// get_alpha (sdf_field.py:122-146) -> weights (rays.py:83-105) -> RGB/Depth/Normal renderers (renderers.py:5-75).
// kBackward recomputes the forward per ray and propagates (g_rgb, g_depth, g_normal, g_weights) to the samples.
template <bool kBackward>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) ray_composite_kernel(const CompositeParams p) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  const int rounds = (S + 31) / 32;
  const float v = __ldg(p.variance);
  const float inv_s_raw = expf(v * 10.f);
  const float inv_s = fminf(fmaxf(inv_s_raw, 1e-6f), 1e6f);
  const float tmin = __int_as_float(p.minmax[0]), tmax = __int_as_float(p.minmax[1]);
  float g_invs = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + wib; r < p.R; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    const float d0 = p.dirs[r * 3], d1 = p.dirs[r * 3 + 1], d2 = p.dirs[r * 3 + 2];
    float al[kMaxRounds], T[kMaxRounds], wt[kMaxRounds];
    float carry = 1.f, W = 0.f, A = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxRounds; ++q) {
      al[q] = 0.f; T[q] = 0.f; wt[q] = 0.f;
      if (q < rounds) {
        const int s = q * 32 + lane;
        const bool ok = s < S;
        const int64_t e = r * S + (ok ? s : 0);
        float alpha = 0.f;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (ok) {
          gx = p.grad[e * 3]; gy = p.grad[e * 3 + 1]; gz = p.grad[e * 3 + 2];
          alpha = neus_alpha(p.sdf[e], gx, gy, gz, d0, d1, d2, p.deltas[e], inv_s, p.cos_anneal).alpha;
        }
        const float f = ok ? (1.f - alpha + 1e-7f) : 1.f;
        const float incl = warp_scan_mul(f, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.f;
        al[q] = alpha;
        T[q] = carry * excl;
        wt[q] = ok ? alpha * T[q] : 0.f;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        if (ok) {
          W += wt[q];
          A += wt[q] * p.starts[e];
          if (p.rgbs != nullptr) { c0 += wt[q] * p.rgbs[e * 3]; c1 += wt[q] * p.rgbs[e * 3 + 1]; c2 += wt[q] * p.rgbs[e * 3 + 2]; }
          const float nrm = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);
          n0 += wt[q] * gx / nrm; n1 += wt[q] * gy / nrm; n2 += wt[q] * gz / nrm;
          if (!kBackward) p.weights[e] = wt[q];
        }
      }
    }
    W = warp_sum(W); A = warp_sum(A);
    const float D_raw = A / (W + 1e-10f);
    if (!kBackward) {
      c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
      n0 = warp_sum(n0); n1 = warp_sum(n1); n2 = warp_sum(n2);
      if (lane == 0) {
        p.depth[r] = fmaxf(fminf(D_raw, tmax), tmin);
        p.normal[r * 3] = n0; p.normal[r * 3 + 1] = n1; p.normal[r * 3 + 2] = n2;
        if (p.rgb != nullptr) {
          // background colour is (0, 0, 0): rgb + bg * (1 - W)
          if (p.clamp_rgb) { c0 = fminf(fmaxf(c0, 0.f), 1.f); c1 = fminf(fmaxf(c1, 0.f), 1.f); c2 = fminf(fmaxf(c2, 0.f), 1.f); }
          p.rgb[r * 3] = c0; p.rgb[r * 3 + 1] = c1; p.rgb[r * 3 + 2] = c2;
        }
      }
    } else {
      const float gr0 = p.g_rgb ? p.g_rgb[r * 3] : 0.f, gr1 = p.g_rgb ? p.g_rgb[r * 3 + 1] : 0.f,
                  gr2 = p.g_rgb ? p.g_rgb[r * 3 + 2] : 0.f;
      float gd = p.g_depth ? p.g_depth[r] : 0.f;
      if (!(D_raw < tmax) || !(D_raw > tmin)) gd = 0.f;     // clipped to the global start range: no gradient
      const float gn0 = p.g_normal ? p.g_normal[r * 3] : 0.f, gn1 = p.g_normal ? p.g_normal[r * 3 + 1] : 0.f,
                  gn2 = p.g_normal ? p.g_normal[r * 3 + 2] : 0.f;
      const float gdw = gd / (W + 1e-10f);
      float suffix = 0.f;   // sum over later rounds of w_j * gw_j
#pragma unroll
      for (int q = kMaxRounds - 1; q >= 0; --q) {
        if (q < rounds) {
          const int s = q * 32 + lane;
          const bool ok = s < S;
          const int64_t e = r * S + (ok ? s : 0);
          float gw = 0.f, gx = 0.f, gy = 0.f, gz = 0.f, nrm = 1.f;
          float rg0 = 0.f, rg1 = 0.f, rg2 = 0.f;
          if (ok) {
            gx = p.grad[e * 3]; gy = p.grad[e * 3 + 1]; gz = p.grad[e * 3 + 2];
            nrm = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);
            if (p.rgbs != nullptr) { rg0 = p.rgbs[e * 3]; rg1 = p.rgbs[e * 3 + 1]; rg2 = p.rgbs[e * 3 + 2]; }
            gw = gr0 * rg0 + gr1 * rg1 + gr2 * rg2 + gdw * (p.starts[e] - D_raw) +
                 (gn0 * gx + gn1 * gy + gn2 * gz) / nrm + (p.g_weights ? p.g_weights[e] : 0.f);
          }
          const float wg = wt[q] * gw;
          const float incl = warp_rscan_add(wg, lane);           // this lane and later lanes of the round
          const float later = incl - wg + suffix;                 // strictly later samples
          suffix += __shfl_sync(0xffffffffu, incl, 0);
          if (ok) {
            // d L / d alpha_i = T_i gw_i - (sum_{j>i} w_j gw_j) / (1 - alpha_i + 1e-7)
            float ga = T[q] * gw - later / (1.f - al[q] + 1e-7f);
            const float dlt = p.deltas[e];
            const AlphaTerms t = neus_alpha(p.sdf[e], gx, gy, gz, d0, d1, d2, dlt, inv_s, p.cos_anneal);
            if (t.raw < 0.f || t.raw > 1.f) ga = 0.f;
            const float den = t.pc + 1e-5f;
            const float da_dpc = t.nc / (den * den), da_dnc = -1.f / den;
            const float spc = t.pc * (1.f - t.pc), snc = t.nc * (1.f - t.nc);
            const float g_prv = ga * da_dpc * inv_s * spc;
            const float g_nxt = ga * da_dnc * inv_s * snc;
            g_invs += ga * (da_dpc * t.prv * spc + da_dnc * t.nxt * snc);
            const float g_ic = (g_nxt - g_prv) * dlt * 0.5f;
            const float dic_dc = 0.5f * (1.f - p.cos_anneal) * ((-t.c * 0.5f + 0.5f) > 0.f ? 1.f : 0.f) +
                                 p.cos_anneal * (t.c < 0.f ? 1.f : 0.f);
            const float g_c = g_ic * dic_dc;
            // normal = grad / max(|grad|, eps): d/d grad
            const float wn = wt[q];
            const float hx = gx / nrm, hy = gy / nrm, hz = gz / nrm;
            const float dotn = hx * gn0 + hy * gn1 + hz * gn2;
            p.g_sdf[e] = g_prv + g_nxt;
            p.g_grad[e * 3] = g_c * d0 + wn * (gn0 - hx * dotn) / nrm;
            p.g_grad[e * 3 + 1] = g_c * d1 + wn * (gn1 - hy * dotn) / nrm;
            p.g_grad[e * 3 + 2] = g_c * d2 + wn * (gn2 - hz * dotn) / nrm;
            if (p.g_rgbs != nullptr) { p.g_rgbs[e * 3] = wn * gr0; p.g_rgbs[e * 3 + 1] = wn * gr1; p.g_rgbs[e * 3 + 2] = wn * gr2; }
          }
        }
      }
    }
  }
  if (kBackward) {
    g_invs = warp_sum(g_invs);
    // inv_s = clip(exp(10 v)): d inv_s / d v = 10 inv_s inside the clip range
    if (lane == 0 && g_invs != 0.f && inv_s_raw >= 1e-6f && inv_s_raw <= 1e6f) atomicAdd(p.g_variance, g_invs * 10.f * inv_s);
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct LossParams {
  const float* depth_pred;   // [R]
  const float* rgb_pred;     // [R, 3] or nullptr
  const float* depth_gt;     // [R]
  const float* rgb_gt;       // [R, 3] or nullptr
  const float* sdf;          // [R, S]
  const float* z;            // [R, S]
  const float* grad;         // [R, S, 3]
  int64_t R; int S;
  float trunc;
  float* sums;               // [11]: numerators [0..4] = sum valid|gt-pred|, sum |rgb diff|, sum relu(trunc-sdf) front,
                             //       sum |z+sdf-gt| sdf_mask, sum (|grad|-1)^2;  counts [5..9] = n valid, 0, n front,
                             //       n sdf_mask, 0;  [10] = sum rgb diff^2 (psnr)
  // backward: coef[5] = upstream * weight / normaliser for depth, rgb, free-space, sdf, eikonal
  const float* coef;
  float* g_depth; float* g_rgb; float* g_sdf; float* g_grad;
};

// SurfaceModel.get_loss (base_surface_model.py:102-211): masks and per-term sums; kBackward writes the gradients of
// sum_k coef[k] * term_k with respect to depth, rgb, sdf and gradients.
template <bool kBackward>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) ray_loss_kernel(const LossParams p) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  float cf[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (kBackward) {
#pragma unroll
    for (int i = 0; i < 5; ++i) cf[i] = p.coef[i];
  }
  for (int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + wib; r < p.R; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    const float gt = p.depth_gt[r];
    const bool valid = gt > 0.f;
    if (lane == 0) {
      const float diff = gt - p.depth_pred[r];
      if (!kBackward) {
        acc[0] += valid ? fabsf(diff) : 0.f;
        acc[1] += valid ? 1.f : 0.f;
      } else {
        // d |gt - pred| / d pred = -sign(gt - pred)
        p.g_depth[r] = valid ? -cf[0] * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) : 0.f;
      }
    }
    if (p.rgb_pred != nullptr && lane < 3) {
      const float diff = p.rgb_pred[r * 3 + lane] - p.rgb_gt[r * 3 + lane];
      if (!kBackward) { acc[2] += fabsf(diff); acc[3] += diff * diff; }
      else p.g_rgb[r * 3 + lane] = cf[1] * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
    }
    for (int s = lane; s < p.S; s += 32) {
      const int64_t e = r * p.S + s;
      const float sdf = p.sdf[e], z = p.z[e];
      const bool front = valid && (z < gt - p.trunc);
      const bool back = valid && (z > gt + p.trunc);
      const bool sm = valid && !(front || back);   // written this way on purpose: nvcc 12.9 miscompiles `valid && !front && !back` (drops the `back` test)
      const float gx = p.grad[e * 3], gy = p.grad[e * 3 + 1], gz = p.grad[e * 3 + 2];
      const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
      const float fs = p.trunc - sdf;
      const float sd = z + sdf - gt;
      if (!kBackward) {
        acc[4] += front ? fmaxf(fs, 0.f) : 0.f;
        acc[5] += front ? 1.f : 0.f;
        acc[6] += sm ? fabsf(sd) : 0.f;
        acc[7] += sm ? 1.f : 0.f;
        acc[8] += (nrm - 1.f) * (nrm - 1.f);
      } else {
        float g = 0.f;
        if (front && fs > 0.f) g -= cf[2];
        if (sm) g += cf[3] * (sd > 0.f ? 1.f : (sd < 0.f ? -1.f : 0.f));
        p.g_sdf[e] = g;
        const float k = (nrm > 0.f) ? cf[4] * 2.f * (nrm - 1.f) / nrm : 0.f;
        p.g_grad[e * 3] = k * gx; p.g_grad[e * 3 + 1] = k * gy; p.g_grad[e * 3 + 2] = k * gz;
      }
    }
  }
  if (!kBackward) {
    __shared__ float s_acc[kWarpsPerBlock][9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float v = warp_sum(acc[i]);
      if (lane == 0) s_acc[wib][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
      // acc order: depth num, n valid, rgb abs, rgb sq, free num, n front, sdf num, n sdf_mask, eikonal num
      const int slot[9] = {0, 5, 1, 10, 2, 7, 3, 8, 4};
      float v = 0.f;
      for (int w2 = 0; w2 < kWarpsPerBlock; ++w2) v += s_acc[w2][threadIdx.x];
      if (v != 0.f) atomicAdd(&p.sums[slot[threadIdx.x]], v);
    }
  }
}

inline int ray_grid(int64_t R) {
  int64_t blocks = (R + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = (int64_t)PV2_SM_COUNT * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" {

int pv2_ray_setup(const float* origins, const float* dirs, const float* noise, int noise_cols, int64_t R, int S0,
                  const float* bbox_host, float near_plane, float* nears, float* fars, float* bins, float* pts,
                  void* stream_) {
  PV2_CHECK_ARG(R >= 0 && S0 >= 2 && bbox_host != nullptr && (noise == nullptr || noise_cols == 1 || noise_cols == S0 + 1));
  if (R == 0) return 0;
  PV2_CHECK_ARG(origins && dirs && nears && fars && bins && pts);
  Bbox6 bb;
  for (int i = 0; i < 6; ++i) bb.v[i] = bbox_host[i];
  ray_setup_kernel<<<pv2_grid_for(R * (S0 + 1), 256), 256, 0, (cudaStream_t)stream_>>>(origins, dirs, noise, noise_cols, R, S0, bb,
                                                                                  near_plane, nears, fars, bins, pts);
  PV2_DONE(1);
}

int pv2_ray_resample(const float* origins, const float* dirs, const float* nears, const float* fars, const float* bins,
                     const float* sdf, const float* noise, int noise_cols, int64_t R, int S0, int Si, float inv_s,
                     int norm_pts, float norm_padding, float* starts, float* deltas, float* pts_norm, float* init_weights,
                     float* new_bins, int32_t* minmax, void* stream_) {
  PV2_CHECK_ARG(R >= 0 && S0 >= 2 && S0 <= kMaxS0 && Si >= 1 && Si + 1 <= kMaxNb);
  PV2_CHECK_ARG(noise == nullptr || noise_cols == 1 || noise_cols == Si + 1);
  if (R == 0) return 0;
  PV2_CHECK_ARG(origins && dirs && nears && fars && bins && sdf && starts && deltas && pts_norm && minmax);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nb = Si + 1;
  ResampleParams p{};
  p.origins = origins; p.dirs = dirs; p.nears = nears; p.fars = fars; p.bins = bins; p.sdf = sdf; p.noise = noise;
  p.noise_cols = noise_cols; p.R = R; p.S0 = S0; p.Si = Si; p.inv_s = inv_s;
  p.u_end = (float)(1.0 - 1.0 / (double)nb);
  p.u_center = (float)(1.0 / (2.0 * (double)nb));
  p.norm_pts = norm_pts; p.norm_scale = (float)(1.0 + (double)norm_padding + 10e-4);
  p.starts = starts; p.deltas = deltas; p.pts_norm = pts_norm; p.init_weights = init_weights; p.new_bins = new_bins;
  p.minmax = minmax;
  const int32_t init[2] = {0x7f800000, 0};   // +inf, 0.0f as float bits (starts are non-negative)
  cudaMemcpyAsync(minmax, init, sizeof(init), cudaMemcpyHostToDevice, stream);
  ray_resample_kernel<<<ray_grid(R), kWarpsPerBlock * 32, 0, stream>>>(p);
  PV2_DONE(2);
}

int pv2_ray_composite_fwd(const float* sdf, const float* grad, const float* rgbs, const float* starts,
                          const float* deltas, const float* dirs, const float* variance, const int32_t* minmax,
                          float cos_anneal, int64_t R, int S, int clamp_rgb, float* weights, float* rgb, float* depth,
                          float* normal, void* stream_) {
  PV2_CHECK_ARG(R >= 0 && S >= 1 && S <= 32 * kMaxRounds);
  if (R == 0) return 0;
  PV2_CHECK_ARG(sdf && grad && starts && deltas && dirs && variance && minmax && weights && depth && normal);
  PV2_CHECK_ARG((rgbs == nullptr) == (rgb == nullptr));
  CompositeParams p{};
  p.sdf = sdf; p.grad = grad; p.rgbs = rgbs; p.starts = starts; p.deltas = deltas; p.dirs = dirs; p.variance = variance;
  p.minmax = minmax; p.cos_anneal = cos_anneal; p.R = R; p.S = S; p.clamp_rgb = clamp_rgb;
  p.weights = weights; p.rgb = rgb; p.depth = depth; p.normal = normal;
  ray_composite_kernel<false><<<ray_grid(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream_>>>(p);
  PV2_DONE(1);
}

int pv2_ray_composite_bwd(const float* sdf, const float* grad, const float* rgbs, const float* starts,
                          const float* deltas, const float* dirs, const float* variance, const int32_t* minmax,
                          float cos_anneal, int64_t R, int S, const float* g_rgb, const float* g_depth,
                          const float* g_normal, const float* g_weights, float* g_sdf, float* g_grad, float* g_rgbs,
                          float* g_variance, void* stream_) {
  PV2_CHECK_ARG(R >= 0 && S >= 1 && S <= 32 * kMaxRounds);
  if (R == 0) return 0;
  PV2_CHECK_ARG(sdf && grad && starts && deltas && dirs && variance && minmax && g_sdf && g_grad && g_variance);
  PV2_CHECK_ARG(rgbs != nullptr || g_rgbs == nullptr);
  CompositeParams p{};
  p.sdf = sdf; p.grad = grad; p.rgbs = rgbs; p.starts = starts; p.deltas = deltas; p.dirs = dirs; p.variance = variance;
  p.minmax = minmax; p.cos_anneal = cos_anneal; p.R = R; p.S = S;
  p.g_rgb = g_rgb; p.g_depth = g_depth; p.g_normal = g_normal; p.g_weights = g_weights;
  p.g_sdf = g_sdf; p.g_grad = g_grad; p.g_rgbs = g_rgbs; p.g_variance = g_variance;
  ray_composite_kernel<true><<<ray_grid(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream_>>>(p);
  PV2_DONE(1);
}

int pv2_ray_loss_fwd(const float* depth_pred, const float* rgb_pred, const float* depth_gt, const float* rgb_gt,
                     const float* sdf, const float* z, const float* grad, int64_t R, int S, float trunc, float* sums,
                     void* stream_) {
  PV2_CHECK_ARG(R >= 0 && S >= 1);
  PV2_CHECK_ARG(sums != nullptr);
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaMemsetAsync(sums, 0, 11 * sizeof(float), stream);
  if (R == 0) return 0;
  PV2_CHECK_ARG(depth_pred && depth_gt && sdf && z && grad && ((rgb_pred == nullptr) == (rgb_gt == nullptr)));
  LossParams p{};
  p.depth_pred = depth_pred; p.rgb_pred = rgb_pred; p.depth_gt = depth_gt; p.rgb_gt = rgb_gt; p.sdf = sdf; p.z = z;
  p.grad = grad; p.R = R; p.S = S; p.trunc = trunc; p.sums = sums;
  ray_loss_kernel<false><<<ray_grid(R), kWarpsPerBlock * 32, 0, stream>>>(p);
  PV2_DONE(2);
}

int pv2_ray_loss_bwd(const float* depth_pred, const float* rgb_pred, const float* depth_gt, const float* rgb_gt,
                     const float* sdf, const float* z, const float* grad, int64_t R, int S, float trunc,
                     const float* coef, float* g_depth, float* g_rgb, float* g_sdf, float* g_grad, void* stream_) {
  PV2_CHECK_ARG(R >= 0 && S >= 1);
  if (R == 0) return 0;
  PV2_CHECK_ARG(depth_pred && depth_gt && sdf && z && grad && coef && g_depth && g_sdf && g_grad);
  PV2_CHECK_ARG((rgb_pred == nullptr) == (rgb_gt == nullptr) && (rgb_pred == nullptr) == (g_rgb == nullptr));
  LossParams p{};
  p.depth_pred = depth_pred; p.rgb_pred = rgb_pred; p.depth_gt = depth_gt; p.rgb_gt = rgb_gt; p.sdf = sdf; p.z = z;
  p.grad = grad; p.R = R; p.S = S; p.trunc = trunc; p.coef = coef;
  p.g_depth = g_depth; p.g_rgb = g_rgb; p.g_sdf = g_sdf; p.g_grad = g_grad;
  ray_loss_kernel<true><<<ray_grid(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream_>>>(p);
  PV2_DONE(1);
}

}  // extern "C"
