// Sparse convolution arithmetic, exact-fp32 SIMT path (any channel count, f32 or bf16 storage).
//
//   y[j,:] = bias + sum_k W_k x[nbr[k][j],:]         (output-stationary: every output row is
//                                                      written exactly once, no atomics)
//   dw[:,k,:] += dy^T x[nbr[k][:],:]                  (split over row chunks, fp32 atomics)
//
// Replaces spconv's SubMConv3d / SparseConv3d / SparseInverseConv3d fwd, dgrad, wgrad
// (call sites spconv_unet_v1m1_base.py:47-66,111-119,135-142,171-177).  The tensor-core
// (tcgen05) variant for 16-aligned channel counts lives in spconv_umma.cu; this file is the
// path for the ragged stem (Cin=6/4) and for bit-faithful fp32 accumulation.
#include "pv2_common.cuh"

namespace {

constexpr int TM = 64;   // output rows per CTA
constexpr int TN = 64;   // output channels per CTA
constexpr int KC = 16;   // reduction chunk (input channels)
constexpr int PAD = 4;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <typename T>
__global__ void __launch_bounds__(256) gather_gemm_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                          int64_t w_sco, int64_t w_sk, const float* __restrict__ bias,
                                                          const int32_t* __restrict__ nbr,
                                                          const int32_t* __restrict__ order, T* __restrict__ y,
                                                          int64_t n_out, int cin, int cout, int kvol) {
  __shared__ float Xs[KC][TM + PAD];
  __shared__ float Ws[KC][TN + PAD];
  __shared__ int32_t rows_s[TM];

  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  const int col0 = blockIdx.y * TN;
  const int tr = (tid / 16) * 4;  // 4 rows per thread
  const int tc = (tid % 16) * 4;  // 4 cols per thread

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const int lr = tid / KC;  // 0..15 (row within a 16-row pass)
  const int lc = tid % KC;  // channel within the chunk

  for (int k = 0; k < kvol; ++k) {
    int any = 0;
    if (tid < TM) {
      int64_t j = row0 + tid;
      int32_t r = (j < n_out) ? __ldg(&nbr[(int64_t)k * n_out + j]) : -1;
      rows_s[tid] = r;
      any = (r >= 0);
    }
    // block-uniform skip of empty (tile, offset) pairs; also publishes rows_s
    if (!__syncthreads_or(any)) continue;

    for (int c0 = 0; c0 < cin; c0 += KC) {
      // gathered X tile -> Xs[ci][row]
#pragma unroll
      for (int p = 0; p < TM / 16; ++p) {
        int r = p * 16 + lr;
        int32_t src = rows_s[r];
        int ci = c0 + lc;
        float v = 0.f;
        if (src >= 0 && ci < cin) v = to_f32<T>(x[(int64_t)src * cin + ci]);
        Xs[lc][r] = v;
      }
      // W_k tile -> Ws[ci][co]
#pragma unroll
      for (int p = 0; p < TN / 16; ++p) {
        int co = p * 16 + lr;
        int ci = c0 + lc;
        float v = 0.f;
        if (col0 + co < cout && ci < cin) v = to_f32<T>(w[(int64_t)(col0 + co) * w_sco + (int64_t)k * w_sk + ci]);
        Ws[lc][co] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&Xs[kk][tr]);
        float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tc]);
        float av[4] = {a.x, a.y, a.z, a.w};
        float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jx = 0; jx < 4; ++jx) acc[i][jx] = fmaf(av[i], bv[jx], acc[i][jx]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t j = row0 + tr + i;
    if (j >= n_out) continue;
    if (order != nullptr) j = __ldg(&order[j]);   // nbr is in tile order: position -> output row
#pragma unroll
    for (int jx = 0; jx < 4; ++jx) {
      int co = col0 + tc + jx;
      if (co < cout) {
        float v = acc[i][jx] + (bias ? __ldg(&bias[co]) : 0.f);
        y[j * cout + co] = from_f32<T>(v);
      }
    }
  }
}

// dw[co,k,ci] += sum_{j in chunk} dy[j,co] * x[nbr[k][j],ci]
// grid: x = row chunk, y = k, z = (co tile, ci tile)
template <typename T>
__global__ void __launch_bounds__(256) wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                    const int32_t* __restrict__ nbr,
                                                    const int32_t* __restrict__ order, float* __restrict__ dw,
                                                    int64_t n_out, int cin, int cout, int kvol, int64_t rows_per_chunk,
                                                    int ci_tiles) {
  constexpr int RC = 16;  // rows per smem pass
  __shared__ float Ds[RC][TN + PAD];  // dy tile   [row][co]
  __shared__ float Xs[RC][TN + PAD];  // x tile    [row][ci]
  __shared__ int32_t rows_s[RC];

  const int tid = threadIdx.x;
  const int k = blockIdx.y;
  const int co0 = (blockIdx.z / ci_tiles) * TN;
  const int ci0 = (blockIdx.z % ci_tiles) * TN;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_chunk;
  int64_t r_end = r_begin + rows_per_chunk;
  if (r_end > n_out) r_end = n_out;
  const int tr = (tid / 16) * 4;  // co
  const int tc = (tid % 16) * 4;  // ci

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const int lr = tid / 16;       // 0..15 row in pass
  const int lc = (tid % 16) * 4; // 4 consecutive channels

  for (int64_t r0 = r_begin; r0 < r_end; r0 += RC) {
    int any = 0;
    if (tid < RC) {
      int64_t j = r0 + tid;
      int32_t r = (j < r_end) ? __ldg(&nbr[(int64_t)k * n_out + j]) : -1;
      rows_s[tid] = r;
      any = (r >= 0);
    }
    if (!__syncthreads_or(any)) continue;
    {
      int64_t j = r0 + lr;
      int32_t src = rows_s[lr];
      if (order != nullptr && src >= 0) j = __ldg(&order[j]);   // nbr is in tile order: position -> dy row
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int co = co0 + lc + e, ci = ci0 + lc + e;
        float dv = 0.f, xv = 0.f;
        if (src >= 0) {
          if (co < cout) dv = to_f32<T>(dy[j * cout + co]);
          if (ci < cin) xv = to_f32<T>(x[(int64_t)src * cin + ci]);
        }
        Ds[lr][lc + e] = dv;
        Xs[lr][lc + e] = xv;
      }
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) {
      float4 a = *reinterpret_cast<const float4*>(&Ds[rr][tr]);
      float4 b = *reinterpret_cast<const float4*>(&Xs[rr][tc]);
      float av[4] = {a.x, a.y, a.z, a.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jx = 0; jx < 4; ++jx) acc[i][jx] = fmaf(av[i], bv[jx], acc[i][jx]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + tr + i;
    if (co >= cout) continue;
#pragma unroll
    for (int jx = 0; jx < 4; ++jx) {
      int ci = ci0 + tc + jx;
      if (ci < cin && acc[i][jx] != 0.f) atomicAdd(&dw[((int64_t)co * kvol + k) * cin + ci], acc[i][jx]);
    }
  }
}

}  // namespace

extern "C" {

int pv2_spconv_gather_gemm_simt(const void* x, const void* w, int64_t w_sco, int64_t w_sk, const float* bias,
                                const int32_t* nbr, const int32_t* order, void* y, int64_t n_in, int64_t n_out, int cin,
                                int cout, int kvol, int dtype, void* stream_) {
  PV2_CHECK_ARG(n_in >= 0 && n_out >= 0 && cin > 0 && cout > 0 && kvol > 0);
  if (n_out == 0) return 0;
  PV2_CHECK_ARG(x && w && nbr && y);
  cudaStream_t stream = (cudaStream_t)stream_;
  dim3 grid((unsigned)((n_out + TM - 1) / TM), (unsigned)((cout + TN - 1) / TN));
  if (dtype == PV2_F32)
    gather_gemm_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (const float*)w, w_sco, w_sk, bias, nbr, order, (float*)y, n_out, cin, cout, kvol);
  else if (dtype == PV2_BF16)
    gather_gemm_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, w_sco, w_sk, bias, nbr, order, (__nv_bfloat16*)y, n_out, cin, cout, kvol);
  else
    return PV2_EUNSUPPORTED;
  PV2_DONE(1);
}

int pv2_spconv_wgrad_simt(const void* x, const void* dy, const int32_t* nbr, const int32_t* order, float* dw, int64_t n_in,
                          int64_t n_out, int cin, int cout, int kvol, int dtype, void* stream_) {
  PV2_CHECK_ARG(n_in >= 0 && n_out >= 0 && cin > 0 && cout > 0 && kvol > 0);
  if (n_out == 0) return 0;
  PV2_CHECK_ARG(x && dy && nbr && dw);
  cudaStream_t stream = (cudaStream_t)stream_;
  int co_tiles = (cout + TN - 1) / TN, ci_tiles = (cin + TN - 1) / TN;
  int64_t per_chunk_blocks = (int64_t)kvol * co_tiles * ci_tiles;
  // aim for ~4 waves of 148 SMs x 4 resident CTAs, but keep chunks >= 256 rows
  int64_t chunks = (4LL * PV2_SM_COUNT * 4 + per_chunk_blocks - 1) / per_chunk_blocks;
  int64_t max_chunks = (n_out + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t rows_per_chunk = ((n_out + chunks - 1) / chunks + 15) / 16 * 16;
  chunks = (n_out + rows_per_chunk - 1) / rows_per_chunk;
  dim3 grid((unsigned)chunks, (unsigned)kvol, (unsigned)(co_tiles * ci_tiles));
  if (dtype == PV2_F32)
    wgrad_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (const float*)dy, nbr, order, dw, n_out, cin, cout, kvol, rows_per_chunk, ci_tiles);
  else if (dtype == PV2_BF16)
    wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, nbr, order, dw, n_out, cin, cout, kvol, rows_per_chunk, ci_tiles);
  else
    return PV2_EUNSUPPORTED;
  PV2_DONE(1);
}

}  // extern "C"
