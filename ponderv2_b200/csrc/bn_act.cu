// BatchNorm1d (batch statistics) + optional residual add + optional ReLU over [N, C] voxel features, forward and
// backward: row a6 of the hot path (ponder/models/sparse_unet/spconv_unet_v1m1_base.py:70-83 `BasicBlock.forward`:
// conv -> bn -> relu -> conv -> bn -> (+ residual) -> relu, and the conv -> bn -> relu stems/down/up blocks :111-180).
//
// torch runs this as BN-stats, BN-apply, add, ReLU forward (5 passes over [N, C]) and ReLU-bwd, BN-reduce, BN-apply
// backward (8 passes).  Here: forward = one statistics pass + one fused apply pass (normalise, affine, + residual, ReLU);
// backward = one reduction pass (d gamma, d beta with the ReLU mask applied on the fly) + one fused apply pass that
// writes dx and the residual gradient.  Pure HBM-bound elementwise/reduction work: float4 accesses, per-block partial
// sums combined in double by a one-block finalize kernel (no atomics, deterministic).
#include <cooperative_groups.h>

#include "pv2_common.cuh"

namespace {

// four consecutive channels of an [N, C] feature matrix stored as fp32 (16 B) or bf16 (8 B); arithmetic is fp32 either way
template <bool kBf16>
__device__ __forceinline__ float4 ld4(const void* base, int64_t i) {
  if constexpr (kBf16) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(base) + i);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return __ldg(reinterpret_cast<const float4*>(base) + i);
  }
}
template <bool kBf16>
__device__ __forceinline__ void st4(void* base, int64_t i, const float4& v) {
  if constexpr (kBf16) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    reinterpret_cast<uint2*>(base)[i] = u;
  } else {
    reinterpret_cast<float4*>(base)[i] = v;
  }
}

constexpr int kRowsPerBlock = 64;    // many small blocks: the reduction passes are latency-bound at 100 k rows
constexpr int kBnThreads = 256;

// Per-block column sums of up to two quantities.  Thread t owns channel group (t % C4) and rows (t / C4) + i * RP.
// MODE 0: (x, x*x).  MODE 1: (dz, dz * xhat) with dz = dy * (y > 0 if relu), xhat = (x - mean) * invstd.
template <int MODE, bool kBf16>
__global__ void __launch_bounds__(kBnThreads) bn_partial_kernel(const void* __restrict__ x, const void* __restrict__ dy,
                                                                const void* __restrict__ y, const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, int64_t N, int C4, int relu,
                                                                float4* __restrict__ partial /*[nblk][2][C4]*/) {
  __shared__ float4 sa[kBnThreads], sb[kBnThreads];
  const int tid = threadIdx.x;
  const int RP = kBnThreads / C4;          // rows per pass
  const int cg = tid % C4, rt = tid / C4;
  const bool active = rt < RP;
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerBlock;
  int64_t rows = N - row0;
  if (rows > kRowsPerBlock) rows = kRowsPerBlock;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  float4 mu = a, is = a;
  if (MODE == 1 && active) {
    mu = *reinterpret_cast<const float4*>(mean + cg * 4);
    is = *reinterpret_cast<const float4*>(invstd + cg * 4);
  }
  if (active) {
#pragma unroll 4
    for (int64_t r = rt; r < rows; r += RP) {
      const int64_t e = (row0 + r) * C4 + cg;
      const float4 v = ld4<kBf16>(x, e);
      if (MODE == 0) {
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        b.x = fmaf(v.x, v.x, b.x); b.y = fmaf(v.y, v.y, b.y); b.z = fmaf(v.z, v.z, b.z); b.w = fmaf(v.w, v.w, b.w);
      } else {
        float4 g = ld4<kBf16>(dy, e);
        if (relu) {
          const float4 o = ld4<kBf16>(y, e);
          g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
        }
        a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
        b.x = fmaf(g.x, (v.x - mu.x) * is.x, b.x); b.y = fmaf(g.y, (v.y - mu.y) * is.y, b.y);
        b.z = fmaf(g.z, (v.z - mu.z) * is.z, b.z); b.w = fmaf(g.w, (v.w - mu.w) * is.w, b.w);
      }
    }
  }
  sa[tid] = a; sb[tid] = b;
  __syncthreads();
  if (tid < C4) {
    float4 ta = sa[tid], tb = sb[tid];
    for (int r = 1; r < RP; ++r) {
      const float4 u = sa[r * C4 + tid], w = sb[r * C4 + tid];
      ta.x += u.x; ta.y += u.y; ta.z += u.z; ta.w += u.w;
      tb.x += w.x; tb.y += w.y; tb.z += w.z; tb.w += w.w;
    }
    partial[((int64_t)blockIdx.x * 2) * C4 + tid] = ta;
    partial[((int64_t)blockIdx.x * 2 + 1) * C4 + tid] = tb;
  }
}

// Combine the per-block partial sums in double, deterministically: block = 32 channels x 32 warps, warp w sums partial
// blocks w, w + 32, ... (coalesced 128-byte reads), the 32 warp totals are added in a fixed order.
// MODE 0: mean / invstd (+ running statistics, torch semantics: momentum, unbiased running variance)
// MODE 1: dgamma = sum dz * xhat, dbeta = sum dz
template <int MODE>
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int nblk, int C, int64_t N,
                                                          float eps, float momentum, float* __restrict__ out_a,
                                                          float* __restrict__ out_b, float* __restrict__ running_mean,
                                                          float* __restrict__ running_var, float* __restrict__ acc_a,
                                                          float* __restrict__ acc_b) {
  __shared__ double s_a[32][32], s_b[32][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double sa = 0.0, sb = 0.0;
  if (c < C) {
#pragma unroll 4
    for (int b = w; b < nblk; b += 32) {
      sa += (double)__ldg(&partial[((int64_t)b * 2) * C + c]);
      sb += (double)__ldg(&partial[((int64_t)b * 2 + 1) * C + c]);
    }
  }
  s_a[w][lane] = sa; s_b[w][lane] = sb;
  __syncthreads();
  if (w != 0 || c >= C) return;
  sa = 0.0; sb = 0.0;
#pragma unroll
  for (int i = 0; i < 32; ++i) { sa += s_a[i][lane]; sb += s_b[i][lane]; }
  if (MODE == 0) {
    const double m = sa / (double)N;
    double var = sb / (double)N - m * m;
    if (var < 0.0) var = 0.0;
    out_a[c] = (float)m;
    out_b[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {
      const double unb = N > 1 ? var * (double)N / (double)(N - 1) : var;
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
    }
  } else {
    out_a[c] = (float)sb;   // dgamma of this call (bn_apply_bwd reads it)
    out_b[c] = (float)sa;   // dbeta
    if (acc_a != nullptr) { acc_a[c] += (float)sb; acc_b[c] += (float)sa; }   // ... and accumulated into the caller's buffer
  }
}

// y = [relu]((x - mean) * invstd * gamma + beta [+ res])
template <bool kBf16>
__global__ void bn_apply_fwd_kernel(const void* __restrict__ x, const void* __restrict__ res, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, int64_t total4, int C4, int relu, void* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C4);
    const float4 mu = *reinterpret_cast<const float4*>(mean + cg * 4);
    const float4 is = *reinterpret_cast<const float4*>(invstd + cg * 4);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + cg * 4);
    const float4 be = *reinterpret_cast<const float4*>(beta + cg * 4);
    const float4 v = ld4<kBf16>(x, i);
    float4 o;
    o.x = (v.x - mu.x) * is.x * ga.x + be.x; o.y = (v.y - mu.y) * is.y * ga.y + be.y;
    o.z = (v.z - mu.z) * is.z * ga.z + be.z; o.w = (v.w - mu.w) * is.w * ga.w + be.w;
    if (res != nullptr) { const float4 r = ld4<kBf16>(res, i); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    st4<kBf16>(y, i, o);
  }
}

// dz = dy * (y > 0 if relu);  dx = gamma * invstd * (dz - dbeta / N - xhat * dgamma / N);  dres = dz
template <bool kBf16>
__global__ void bn_apply_bwd_kernel(const void* __restrict__ x, const void* __restrict__ dy, const void* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, int64_t total4, int C4, int relu, float inv_n,
                                    void* __restrict__ dx, void* __restrict__ dres) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C4);
    const float4 mu = *reinterpret_cast<const float4*>(mean + cg * 4);
    const float4 is = *reinterpret_cast<const float4*>(invstd + cg * 4);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + cg * 4);
    const float4 dg = *reinterpret_cast<const float4*>(dgamma + cg * 4);
    const float4 db = *reinterpret_cast<const float4*>(dbeta + cg * 4);
    const float4 v = ld4<kBf16>(x, i);
    float4 g = ld4<kBf16>(dy, i);
    if (relu) {
      const float4 o = ld4<kBf16>(y, i);
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    if (dres != nullptr) st4<kBf16>(dres, i, g);
    float4 d;
    d.x = ga.x * is.x * (g.x - db.x * inv_n - (v.x - mu.x) * is.x * dg.x * inv_n);
    d.y = ga.y * is.y * (g.y - db.y * inv_n - (v.y - mu.y) * is.y * dg.y * inv_n);
    d.z = ga.z * is.z * (g.z - db.z * inv_n - (v.z - mu.z) * is.z * dg.z * inv_n);
    d.w = ga.w * is.w * (g.w - db.w * inv_n - (v.w - mu.w) * is.w * dg.w * inv_n);
    st4<kBf16>(dx, i, d);
  }
}

// ---- small batches (N <= kSmallN rows: the deep U-Net levels, 381 / 1567 voxels at C2) ---------------------------------
// One launch instead of three.  A thread-block CLUSTER of kSmallSplit CTAs owns 32 channels (8 float4 groups): each CTA
// sums its slice of the rows (8 channel groups x 32 row lanes, double accumulation), the CTAs exchange their totals
// through distributed shared memory in a fixed order (deterministic), every CTA derives the statistics and applies them
// to its own rows, which are still in L1/L2.  (The first version ran one CTA per 32 channels: 8 CTAs for C = 256, 25 -
// 35 us per launch, 1.7 ms per step in profiles/r2s_launch_summary.txt - slower than the three launches it replaced.)
constexpr int kSmallN = 2048;
constexpr int kSmallSplit = 8;

__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return __hiloint2double(hi, lo);
}

// Block + cluster reduction of the per-thread sums a[4], b[4] (thread = channel group tid & 7, row lane tid >> 3).
// Returns, in threads tid < 32 (channel group tid & 7, component tid >> 3), the cluster totals of a and b.
__device__ __forceinline__ void small_reduce(double a[4], double b[4], double (*s_w)[2][8][4], double (*s_t)[8][4],
                                             double& tot_a, double& tot_b) {
  namespace cgp = cooperative_groups;
  cgp::cluster_group cluster = cgp::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cgl = tid & 7;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] += shfl_xor_f64(a[i], 8);  b[i] += shfl_xor_f64(b[i], 8);
    a[i] += shfl_xor_f64(a[i], 16); b[i] += shfl_xor_f64(b[i], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { s_w[warp][0][cgl][i] = a[i]; s_w[warp][1][cgl][i] = b[i]; }
  }
  __syncthreads();
  if (tid < 32) {
    const int c = tid & 7, i = tid >> 3;
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { sa += s_w[w][0][c][i]; sb += s_w[w][1][c][i]; }
    s_t[0][c][i] = sa; s_t[1][c][i] = sb;
  }
  cluster.sync();
  tot_a = 0.0; tot_b = 0.0;
  if (tid < 32) {
    const int c = tid & 7, i = tid >> 3;
#pragma unroll
    for (int q = 0; q < kSmallSplit; ++q) {
      const double* remote = reinterpret_cast<const double*>(cluster.map_shared_rank(&s_t[0][0][0], q));
      tot_a += remote[(0 * 8 + c) * 4 + i];
      tot_b += remote[(1 * 8 + c) * 4 + i];
    }
  }
}

template <bool kBf16>
__global__ void __cluster_dims__(1, kSmallSplit, 1) __launch_bounds__(256)
bn_small_fwd_kernel(const void* __restrict__ x, const void* __restrict__ res, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                    float momentum, float eps, int relu, int64_t N, int C4, void* __restrict__ y, float* __restrict__ mean,
                    float* __restrict__ invstd) {
  namespace cgp = cooperative_groups;
  __shared__ double s_w[8][2][8][4], s_t[2][8][4];
  __shared__ float s_mu[8][4], s_is[8][4];
  const int cgl = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int cg = blockIdx.x * 8 + cgl;
  const bool on = cg < C4;
  const int64_t chunk = (N + kSmallSplit - 1) / kSmallSplit;
  const int64_t r0 = (int64_t)blockIdx.y * chunk;
  const int64_t r1 = (r0 + chunk < N) ? r0 + chunk : N;
  double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  if (on) {
#pragma unroll 4
    for (int64_t r = r0 + rl; r < r1; r += 32) {
      const float4 v = ld4<kBf16>(x, r * C4 + cg);
      a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
      b[0] += (double)v.x * v.x; b[1] += (double)v.y * v.y; b[2] += (double)v.z * v.z; b[3] += (double)v.w * v.w;
    }
  }
  double sa, sb;
  small_reduce(a, b, s_w, s_t, sa, sb);
  if (threadIdx.x < 32) {
    const int c = threadIdx.x & 7, i = threadIdx.x >> 3;
    const int ch = (blockIdx.x * 8 + c) * 4 + i;
    const double m = sa / (double)N;
    double var = sb / (double)N - m * m;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    s_mu[c][i] = (float)m; s_is[c][i] = is;
    if (blockIdx.y == 0 && blockIdx.x * 8 + c < C4) {
      mean[ch] = (float)m; invstd[ch] = is;
      if (running_mean != nullptr) {
        const double unb = N > 1 ? var * (double)N / (double)(N - 1) : var;
        running_mean[ch] = (float)((1.0 - momentum) * running_mean[ch] + momentum * m);
        running_var[ch] = (float)((1.0 - momentum) * running_var[ch] + momentum * unb);
      }
    }
  }
  cgp::this_cluster().sync();   // statistics visible to the block; no CTA leaves while its totals may still be read
  if (!on) return;
  const float4 mu = make_float4(s_mu[cgl][0], s_mu[cgl][1], s_mu[cgl][2], s_mu[cgl][3]);
  const float4 is = make_float4(s_is[cgl][0], s_is[cgl][1], s_is[cgl][2], s_is[cgl][3]);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + cg * 4);
  const float4 be = *reinterpret_cast<const float4*>(beta + cg * 4);
#pragma unroll 2
  for (int64_t r = r0 + rl; r < r1; r += 32) {
    const int64_t e = r * C4 + cg;
    const float4 v = ld4<kBf16>(x, e);
    float4 o;
    o.x = (v.x - mu.x) * is.x * ga.x + be.x; o.y = (v.y - mu.y) * is.y * ga.y + be.y;
    o.z = (v.z - mu.z) * is.z * ga.z + be.z; o.w = (v.w - mu.w) * is.w * ga.w + be.w;
    if (res != nullptr) { const float4 q = ld4<kBf16>(res, e); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    st4<kBf16>(y, e, o);
  }
}

template <bool kBf16>
__global__ void __cluster_dims__(1, kSmallSplit, 1) __launch_bounds__(256)
bn_small_bwd_kernel(const void* __restrict__ x, const void* __restrict__ dy, const void* __restrict__ y,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    int relu, int64_t N, int C4, void* __restrict__ dx, void* __restrict__ dres, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, int accumulate) {
  namespace cgp = cooperative_groups;
  __shared__ double s_w[8][2][8][4], s_t[2][8][4];
  __shared__ float s_dg[8][4], s_db[8][4];
  const int cgl = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int cg = blockIdx.x * 8 + cgl;
  const bool on = cg < C4;
  const int64_t chunk = (N + kSmallSplit - 1) / kSmallSplit;
  const int64_t r0 = (int64_t)blockIdx.y * chunk;
  const int64_t r1 = (r0 + chunk < N) ? r0 + chunk : N;
  float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu, ga = mu;
  if (on) {
    mu = *reinterpret_cast<const float4*>(mean + cg * 4);
    is = *reinterpret_cast<const float4*>(invstd + cg * 4);
    ga = *reinterpret_cast<const float4*>(gamma + cg * 4);
  }
  auto dz_of = [&](int64_t e) {
    float4 g = ld4<kBf16>(dy, e);
    if (relu) {
      const float4 o = ld4<kBf16>(y, e);
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    return g;
  };
  double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  if (on) {
#pragma unroll 2
    for (int64_t r = r0 + rl; r < r1; r += 32) {
      const int64_t e = r * C4 + cg;
      const float4 v = ld4<kBf16>(x, e);
      const float4 g = dz_of(e);
      a[0] += g.x; a[1] += g.y; a[2] += g.z; a[3] += g.w;
      b[0] += (double)g.x * ((v.x - mu.x) * is.x); b[1] += (double)g.y * ((v.y - mu.y) * is.y);
      b[2] += (double)g.z * ((v.z - mu.z) * is.z); b[3] += (double)g.w * ((v.w - mu.w) * is.w);
    }
  }
  double sa, sb;
  small_reduce(a, b, s_w, s_t, sa, sb);
  if (threadIdx.x < 32) {
    const int c = threadIdx.x & 7, i = threadIdx.x >> 3;
    const int ch = (blockIdx.x * 8 + c) * 4 + i;
    s_db[c][i] = (float)sa; s_dg[c][i] = (float)sb;
    if (blockIdx.y == 0 && blockIdx.x * 8 + c < C4) {
      if (accumulate) { dbeta[ch] += (float)sa; dgamma[ch] += (float)sb; }
      else { dbeta[ch] = (float)sa; dgamma[ch] = (float)sb; }
    }
  }
  cgp::this_cluster().sync();
  if (!on) return;
  const float inv_n = 1.0f / (float)N;
  const float4 dg = make_float4(s_dg[cgl][0], s_dg[cgl][1], s_dg[cgl][2], s_dg[cgl][3]);
  const float4 db = make_float4(s_db[cgl][0], s_db[cgl][1], s_db[cgl][2], s_db[cgl][3]);
#pragma unroll 2
  for (int64_t r = r0 + rl; r < r1; r += 32) {
    const int64_t e = r * C4 + cg;
    const float4 v = ld4<kBf16>(x, e);
    const float4 g = dz_of(e);
    if (dres != nullptr) st4<kBf16>(dres, e, g);
    float4 d;
    d.x = ga.x * is.x * (g.x - db.x * inv_n - (v.x - mu.x) * is.x * dg.x * inv_n);
    d.y = ga.y * is.y * (g.y - db.y * inv_n - (v.y - mu.y) * is.y * dg.y * inv_n);
    d.z = ga.z * is.z * (g.z - db.z * inv_n - (v.z - mu.z) * is.z * dg.z * inv_n);
    d.w = ga.w * is.w * (g.w - db.w * inv_n - (v.w - mu.w) * is.w * dg.w * inv_n);
    st4<kBf16>(dx, e, d);
  }
}

inline int nblocks_for(int64_t n) { return (int)((n + kRowsPerBlock - 1) / kRowsPerBlock); }
inline bool shape_ok(int64_t n, int c) { return n >= 0 && c >= 4 && c <= 1024 && (c % 4) == 0 && (kBnThreads / (c / 4)) >= 1; }

}  // namespace

extern "C" {

size_t pv2_bn_workspace_bytes(int64_t n, int c) {
  if (!shape_ok(n, c)) return 0;
  return (((size_t)nblocks_for(n) * 2 * c + 2 * (size_t)c) * sizeof(float) + 255) / 256 * 256;   // partial sums + one [2][c] scratch
}

}  // extern "C"

template <bool kBf16>
static int bn_fwd_t(const void* x, const void* res, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, int relu, int64_t n, int c, void* y, float* mean,
                    float* invstd, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PV2_CHECK_ARG(shape_ok(n, c));
  if (n == 0) return 0;
  PV2_CHECK_ARG(x && gamma && beta && y && mean && invstd && workspace);
  const uintptr_t io_mask = kBf16 ? 7 : 15;
  PV2_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)res) & io_mask) == 0);
  PV2_CHECK_ARG((((uintptr_t)mean | (uintptr_t)invstd | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0);
  if (workspace_bytes < pv2_bn_workspace_bytes(n, c)) return PV2_EWORKSPACE;
  const int nblk = nblocks_for(n), C4 = c / 4;
  if (n <= kSmallN) {
    bn_small_fwd_kernel<kBf16><<<dim3((C4 + 7) / 8, kSmallSplit), 256, 0, stream>>>(x, res, gamma, beta, running_mean, running_var, momentum, eps,
                                                                relu, n, C4, y, mean, invstd);
    PV2_DONE(1);
  }
  bn_partial_kernel<0, kBf16><<<nblk, kBnThreads, 0, stream>>>(x, nullptr, nullptr, nullptr, nullptr, n, C4, 0,
                                                                (float4*)workspace);
  bn_finalize_kernel<0><<<(c + 31) / 32, 1024, 0, stream>>>((const float*)workspace, nblk, c, n, eps, momentum, mean, invstd,
                                                             running_mean, running_var, nullptr, nullptr);
  const int64_t total4 = n * C4;
  bn_apply_fwd_kernel<kBf16><<<pv2_grid_for(total4, 256), 256, 0, stream>>>(x, res, mean, invstd, gamma, beta, total4, C4,
                                                                           relu, y);
  PV2_DONE(3);
}

template <bool kBf16>
static int bn_bwd_t(const void* x, const void* dy, const void* y, const float* gamma, const float* mean,
                    const float* invstd, int relu, int64_t n, int c, void* dx, void* dres, float* dgamma, float* dbeta,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream, int accumulate = 0) {
  PV2_CHECK_ARG(shape_ok(n, c));
  if (n == 0) return 0;
  PV2_CHECK_ARG(x && dy && gamma && mean && invstd && dx && dgamma && dbeta && workspace && (!relu || y));
  const uintptr_t io_mask = kBf16 ? 7 : 15;
  PV2_CHECK_ARG((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)y | (uintptr_t)dx | (uintptr_t)dres) & io_mask) == 0);
  PV2_CHECK_ARG((((uintptr_t)mean | (uintptr_t)invstd | (uintptr_t)gamma | (uintptr_t)dgamma | (uintptr_t)dbeta) & 15) == 0);
  if (workspace_bytes < pv2_bn_workspace_bytes(n, c)) return PV2_EWORKSPACE;
  const int nblk = nblocks_for(n), C4 = c / 4;
  if (n <= kSmallN) {
    bn_small_bwd_kernel<kBf16><<<dim3((C4 + 7) / 8, kSmallSplit), 256, 0, stream>>>(x, dy, y, mean, invstd, gamma, relu, n, C4, dx, dres, dgamma,
                                                                dbeta, accumulate);
    PV2_DONE(1);
  }
  bn_partial_kernel<1, kBf16><<<nblk, kBnThreads, 0, stream>>>(x, dy, y, mean, invstd, n, C4, relu, (float4*)workspace);
  // accumulate mode: this call's sums go to the scratch behind the partials (read by the apply pass) and are added into
  // dgamma / dbeta, which then are slices of the caller's gradient buffer
  float* own_g = accumulate ? (float*)workspace + (size_t)nblk * 2 * c : dgamma;
  float* own_b = accumulate ? own_g + c : dbeta;
  bn_finalize_kernel<1><<<(c + 31) / 32, 1024, 0, stream>>>((const float*)workspace, nblk, c, n, 0.f, 0.f, own_g, own_b, nullptr,
                                                             nullptr, accumulate ? dgamma : nullptr, accumulate ? dbeta : nullptr);
  const int64_t total4 = n * C4;
  bn_apply_bwd_kernel<kBf16><<<pv2_grid_for(total4, 256), 256, 0, stream>>>(x, dy, y, mean, invstd, gamma, own_g, own_b,
                                                                           total4, C4, relu, 1.0f / (float)n, dx, dres);
  PV2_DONE(3);
}

extern "C" {

int pv2_bn_act_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float momentum, float eps, int relu, int64_t n, int c, float* y, float* mean,
                   float* invstd, void* workspace, size_t workspace_bytes, void* stream_) {
  return bn_fwd_t<false>(x, res, gamma, beta, running_mean, running_var, momentum, eps, relu, n, c, y, mean, invstd,
                         workspace, workspace_bytes, (cudaStream_t)stream_);
}

int pv2_bn_act_bwd(const float* x, const float* dy, const float* y, const float* gamma, const float* mean,
                   const float* invstd, int relu, int64_t n, int c, float* dx, float* dres, float* dgamma, float* dbeta,
                   void* workspace, size_t workspace_bytes, void* stream_) {
  return bn_bwd_t<false>(x, dy, y, gamma, mean, invstd, relu, n, c, dx, dres, dgamma, dbeta, workspace, workspace_bytes,
                         (cudaStream_t)stream_);
}

/* Same with the [n, c] feature matrices (x, res, y / dy, dx, dres) in `dtype` (PV2_F32 or PV2_BF16): the bf16 backbone
 * of BASELINE configs[2] (the reference runs fp16 autocast: BatchNorm takes half inputs, keeps fp32 statistics and
 * parameters, returns half).  Statistics, gamma / beta and their gradients stay fp32. */
int pv2_bn_act_fwd_t(const void* x, const void* res, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int relu, int64_t n, int c, void* y, float* mean,
                     float* invstd, int dtype, void* workspace, size_t workspace_bytes, void* stream_) {
  if (dtype == PV2_F32)
    return bn_fwd_t<false>(x, res, gamma, beta, running_mean, running_var, momentum, eps, relu, n, c, y, mean, invstd,
                           workspace, workspace_bytes, (cudaStream_t)stream_);
  if (dtype == PV2_BF16)
    return bn_fwd_t<true>(x, res, gamma, beta, running_mean, running_var, momentum, eps, relu, n, c, y, mean, invstd,
                          workspace, workspace_bytes, (cudaStream_t)stream_);
  return PV2_EUNSUPPORTED;
}

int pv2_bn_act_bwd_t(const void* x, const void* dy, const void* y, const float* gamma, const float* mean,
                     const float* invstd, int relu, int64_t n, int c, void* dx, void* dres, float* dgamma, float* dbeta,
                     int accumulate, int dtype, void* workspace, size_t workspace_bytes, void* stream_) {
  if (dtype == PV2_F32)
    return bn_bwd_t<false>(x, dy, y, gamma, mean, invstd, relu, n, c, dx, dres, dgamma, dbeta, workspace, workspace_bytes,
                           (cudaStream_t)stream_, accumulate);
  if (dtype == PV2_BF16)
    return bn_bwd_t<true>(x, dy, y, gamma, mean, invstd, relu, n, c, dx, dres, dgamma, dbeta, workspace, workspace_bytes,
                          (cudaStream_t)stream_, accumulate);
  return PV2_EUNSUPPORTED;
}

}  // extern "C"
